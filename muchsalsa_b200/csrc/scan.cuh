// scan.cuh — device-wide exclusive prefix sum (reduce / scan-partials / down-sweep), hand-written.
//
// The arrays scanned on this path are per-line, per-match or per-unitig: one to two orders of magnitude smaller than
// the edge-match stream, so the reduce-then-scan form (2 reads + 1 write of the input, two launches) is used instead of a
// single-pass look-back scan.  Input values come from a functor so flags packed in bytes / derived counts need no staging array.
#pragma once

#include "common.cuh"

namespace msgpu {

constexpr int kScanBlock = 256;
constexpr int kScanItems = 8; // contiguous items per thread
constexpr int kScanTile  = kScanBlock * kScanItems;

// Per-tile sums; the block that finishes last turns them into exclusive tile offsets in place and writes the grand
// total (the classic "last block done" pattern: one launch instead of a reduce and a scan-of-partials launch).
// `done` is a zero-initialised counter that the last block resets, so it can be reused by the next scan on the stream.
template <class T, class InFn>
__global__ void __launch_bounds__(kScanBlock)
    scan_reduce_kernel(InFn in, uint64_t n, T *__restrict__ partials, T *__restrict__ total_out, unsigned *done) {
  __shared__ T    sm[33];
  __shared__ bool is_last;
  uint64_t        base = uint64_t(blockIdx.x) * kScanTile + uint64_t(threadIdx.x) * kScanItems;
  T               s    = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    uint64_t i = base + j;
    if (i < n) s += in(i);
  }
  T total;
  block_exclusive_scan<T>(s, sm, &total);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = total;
    __threadfence();
    is_last = atomicAdd(done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const uint64_t n_tiles = gridDim.x;
  T              carry   = 0;
  for (uint64_t start = 0; start < n_tiles; start += kScanBlock) {
    uint64_t i = start + threadIdx.x;
    T        v = i < n_tiles ? __ldcg(partials + i) : T(0);
    T        tot;
    T        ex = block_exclusive_scan<T>(v, sm, &tot);
    if (i < n_tiles) partials[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) {
    if (total_out) *total_out = carry;
    *done = 0;
  }
}

template <class T, class InFn>
__global__ void __launch_bounds__(kScanBlock)
    scan_downsweep_kernel(InFn in, uint64_t n, const T *__restrict__ partials, T *out /* may alias the input */) {
  __shared__ T sm[33];
  uint64_t     base = uint64_t(blockIdx.x) * kScanTile + uint64_t(threadIdx.x) * kScanItems;
  T            v[kScanItems];
  T            s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    uint64_t i = base + j;
    v[j]       = i < n ? in(i) : T(0);
    s += v[j];
  }
  T total;
  T ex = block_exclusive_scan<T>(s, sm, &total) + partials[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    uint64_t i = base + j;
    if (i < n) out[i] = ex;
    ex += v[j];
  }
}

// out[i] = sum_{j<i} in(j) for i < n; *total_out (device) = sum of all.  `out` may be null (total only).
// Scratch comes from the arena.  n == 0 is allowed (total = 0).
template <class T, class InFn>
void exclusive_scan(Arena &arena, cudaStream_t st, InFn in, uint64_t n, T *out, T *total_out) {
  if (n == 0) {
    if (total_out) MSGPU_CUDA(cudaMemsetAsync(total_out, 0, sizeof(T), st));
    return;
  }
  uint64_t n_tiles  = (n + kScanTile - 1) / kScanTile;
  T       *partials = arena.alloc<T>(n_tiles);
  MSGPU_LAUNCH("scan_reduce", st, scan_reduce_kernel<T, InFn><<<(unsigned)n_tiles, kScanBlock, 0, st>>>(in, n, partials, total_out, arena.counter()));
  if (out) {
    MSGPU_LAUNCH("scan_downsweep", st, scan_downsweep_kernel<T, InFn><<<(unsigned)n_tiles, kScanBlock, 0, st>>>(in, n, partials, out));
  }
}

// common input functors
template <class S, class T> struct LoadAs {
  const S *p;
  __device__ T operator()(uint64_t i) const { return T(p[i]); }
};
// 1 iff (flags[i] & mask) == mask
template <class T> struct FlagBit {
  const uint8_t *p;
  uint8_t        mask;
  __device__ T operator()(uint64_t i) const { return T((p[i] & mask) == mask); }
};

} // namespace msgpu
