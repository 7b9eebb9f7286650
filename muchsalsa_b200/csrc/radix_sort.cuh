// radix_sort.cuh — stable LSD radix sort of (key, value) pairs in HBM, hand-written.
//
// Per 8-bit digit: (1) per-tile digit histogram, (2) exclusive scan of the digit-major histogram matrix,
// (3) stable scatter.  The scatter ranks a chunk of 256 keys with __match_any_sync (peers with the same digit in the
// warp) + a per-warp count table + a per-digit running base, so equal keys keep their input order.
// Only the digit positions the caller declares significant are visited (ids are dense, so a key of two 17-bit ids
// needs 6 digit passes, not 8).
#pragma once

#include "common.cuh"
#include "scan.cuh"

namespace msgpu {

constexpr int kSortBlock = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile  = kSortBlock * kSortItems;

template <class K> __device__ __forceinline__ unsigned digit_of(K key, int shift, unsigned mask) {
  return unsigned(key >> shift) & mask;
}

template <class K>
__global__ void __launch_bounds__(kSortBlock)
    radix_hist_kernel(const K *__restrict__ keys, uint64_t n, int shift, unsigned mask, uint32_t n_tiles,
                      uint32_t *__restrict__ hist /* [256][n_tiles] */) {
  pdl_sync();
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  uint64_t base = uint64_t(blockIdx.x) * kSortTile;
#pragma unroll 4
  for (int j = 0; j < kSortItems; ++j) {
    uint64_t i = base + uint64_t(j) * kSortBlock + threadIdx.x;
    if (i < n) atomicAdd(&h[digit_of(keys[i], shift, mask)], 1u);
  }
  __syncthreads();
  hist[uint64_t(threadIdx.x) * n_tiles + blockIdx.x] = h[threadIdx.x];
}

template <class K, class V>
__global__ void __launch_bounds__(kSortBlock)
    radix_scatter_kernel(const K *__restrict__ keys_in, const V *__restrict__ vals_in, K *__restrict__ keys_out,
                         V *__restrict__ vals_out, uint64_t n, int shift, unsigned mask, uint32_t n_tiles,
                         const uint32_t *__restrict__ hist_scanned) {
  pdl_sync();
  constexpr int       kWarps = kSortBlock / 32;
  __shared__ uint32_t warp_cnt[kWarps][256];
  __shared__ uint32_t bin_base[256];
  __shared__ uint32_t chunk_base[256];
  const unsigned      t = threadIdx.x, w = warp_id(), l = lane_id();
  bin_base[t] = hist_scanned[uint64_t(t) * n_tiles + blockIdx.x];
  uint64_t base = uint64_t(blockIdx.x) * kSortTile;
  for (int j = 0; j < kSortItems; ++j) {
#pragma unroll
    for (int ww = 0; ww < kWarps; ++ww) warp_cnt[ww][t] = 0;
    __syncthreads();
    uint64_t i     = base + uint64_t(j) * kSortBlock + t;
    bool     valid = i < n;
    K        key   = valid ? keys_in[i] : K(0);
    V        val   = valid ? vals_in[i] : V(0);
    unsigned d     = valid ? digit_of(key, shift, mask) : 256u; // invalid lanes form their own peer group
    unsigned peers = __match_any_sync(0xffffffffu, d);
    unsigned rank  = __popc(peers & ((1u << l) - 1u));
    if (valid && rank == 0) warp_cnt[w][d] = __popc(peers);
    __syncthreads();
    {
      uint32_t sum = 0;
#pragma unroll
      for (int ww = 0; ww < kWarps; ++ww) {
        uint32_t c      = warp_cnt[ww][t];
        warp_cnt[ww][t] = sum;
        sum += c;
      }
      chunk_base[t] = bin_base[t];
      bin_base[t] += sum;
    }
    __syncthreads();
    if (valid) {
      uint32_t dst  = chunk_base[d] + warp_cnt[w][d] + rank;
      keys_out[dst] = key;
      vals_out[dst] = val;
    }
    __syncthreads();
  }
}

// Sorts n pairs by the key bits listed in `bit_ranges` = {lo0, hi0, lo1, hi1, ...} (half-open, least significant range
// first).  Result ends up in (keys, vals); (keys_tmp, vals_tmp) is scratch of the same size.  n < 2^32.
template <class K, class V>
void radix_sort_pairs(Arena &arena, cudaStream_t st, K *&keys, V *&vals, K *&keys_tmp, V *&vals_tmp, uint64_t n,
                      const std::vector<int> &bit_ranges) {
  if (n <= 1) return;
  uint32_t  n_tiles = uint32_t((n + kSortTile - 1) / kSortTile);
  uint32_t *hist    = arena.alloc<uint32_t>(uint64_t(256) * n_tiles);
  for (size_t r = 0; r + 1 < bit_ranges.size(); r += 2) {
    for (int lo = bit_ranges[r]; lo < bit_ranges[r + 1]; lo += 8) {
      int      bits = bit_ranges[r + 1] - lo < 8 ? bit_ranges[r + 1] - lo : 8;
      unsigned mask = (1u << bits) - 1u;
      MSGPU_LAUNCH("radix_hist", st, radix_hist_kernel<K><<<n_tiles, kSortBlock, 0, st>>>(keys, n, lo, mask, n_tiles, hist));
      exclusive_scan<uint32_t>(arena, st, LoadAs<uint32_t, uint32_t>{hist}, uint64_t(256) * n_tiles, hist,
                               (uint32_t *)nullptr);
      MSGPU_LAUNCH("radix_scatter", st, radix_scatter_kernel<K, V><<<n_tiles, kSortBlock, 0, st>>>(keys, vals, keys_tmp, vals_tmp, n, lo, mask, n_tiles, hist));
      std::swap(keys, keys_tmp);
      std::swap(vals, vals_tmp);
    }
  }
}

inline int bits_for(uint64_t n_values) { // bits needed to represent ids 0 .. n_values-1
  int b = 1;
  while (b < 64 && (uint64_t(1) << b) < n_values) ++b;
  return b;
}

} // namespace msgpu
