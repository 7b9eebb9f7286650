// scan.cuh — device-wide exclusive prefix sum in ONE launch (single pass, decoupled look-back), hand-written.
//
// The arrays scanned on this path are per-line, per-match, per-read or per-unitig: tens to hundreds of tiles, so what a
// scan costs is its launches, not its bytes.  One kernel: a block takes a ticket (tiles are numbered in the order blocks
// start, so a tile only ever waits for tiles that are already running), reduces its 2048 items, publishes the tile
// aggregate, looks back over its predecessors' aggregates / inclusive prefixes with one warp, publishes its own
// inclusive prefix and writes its outputs.  Every input value is evaluated once.  Input values come from a functor so
// flags packed in bytes / derived counts need no staging array.
//
// The per-tile status words live in a region of the arena that outlives reset() and is used by scans only.  A status
// is ONE 64-bit word — value << 18 | epoch << 2 | state (1: tile aggregate, 2: inclusive prefix) — so a single load
// observes value and state together and no fence is needed.  The epoch is the serial number of the scan that wrote
// the word: stale words read as "not ready" and no launch is spent on clearing them.  Sums must stay below 2^46
// (7e13; the largest quantity scanned here is the pair count, ~1e11 at human scale); a larger one traps.
// The look-back window is the whole block (256 predecessors per round trip), so a scan over a few hundred tiles costs
// one or two memory round trips after its own reduce.
#pragma once

#include "common.cuh"

namespace msgpu {

constexpr int      kScanBlock     = 256;
constexpr int      kScanItems     = 8; // contiguous items per thread
constexpr int      kScanTile      = kScanBlock * kScanItems;
constexpr int      kScanValShift  = 18;
constexpr unsigned kScanEpochMask = 0xffffu;

template <class T, class InFn>
__global__ void __launch_bounds__(kScanBlock)
    scan_onepass_kernel(InFn in, uint64_t n, T *out /* may alias the input; may be null */, T *__restrict__ total_out,
                        unsigned long long *state, unsigned epoch, unsigned *counters /* [0] ticket, [1] done */) {
  __shared__ T        sm[33];
  __shared__ unsigned s_tile;
  __shared__ T        w_sum[kScanBlock / 32];
  __shared__ unsigned w_inc[kScanBlock / 32];
  if (threadIdx.x == 0) s_tile = atomicAdd(&counters[0], 1u);
  __syncthreads();
  const unsigned tile = s_tile;
  const uint64_t base = uint64_t(tile) * kScanTile + uint64_t(threadIdx.x) * kScanItems;
  T              v[kScanItems];
  T              s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    const uint64_t i = base + j;
    v[j]             = i < n ? in(i) : T(0);
    s += v[j];
  }
  T total;
  T ex = block_exclusive_scan<T>(s, sm, &total);
  volatile unsigned long long *vs  = state;
  const unsigned long long     tag = (unsigned long long)(epoch & kScanEpochMask) << 2;
  if ((unsigned long long)total >> (64 - kScanValShift)) __trap();
  if (threadIdx.x == 0) vs[tile] = ((unsigned long long)total << kScanValShift) | tag | (tile == 0 ? 2ull : 1ull);
  T prefix = 0;
  if (tile > 0) {
    const unsigned l = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long      hi = (long long)tile - 1; // window = tiles hi, hi-1, ..., hi-255
    while (true) {
      const long long    p    = hi - (long long)threadIdx.x;
      unsigned long long word = 0;
      if (p >= 0) {
        do { word = vs[p]; } while ((word & (kScanEpochMask << 2)) != tag || (word & 3ull) == 0ull);
      }
      const bool     is_incl = p >= 0 && (word & 3ull) == 2ull;
      const unsigned im      = __ballot_sync(0xffffffffu, is_incl);
      const unsigned stop    = im ? unsigned(__ffs(im) - 1) : 32u; // nearest predecessor with an inclusive prefix
      T              x       = (p >= 0 && l <= stop) ? T(word >> kScanValShift) : T(0);
      x                      = warp_reduce_sum(x);
      if (l == 0) {
        w_sum[w] = x;
        w_inc[w] = im != 0u;
      }
      __syncthreads();
      bool done = false;
#pragma unroll
      for (int ww = 0; ww < kScanBlock / 32; ++ww) {
        if (!done) prefix += w_sum[ww];
        done = done || w_inc[ww] != 0u;
      }
      __syncthreads();
      if (done || hi - kScanBlock < 0) break;
      hi -= kScanBlock;
    }
    if ((unsigned long long)(prefix + total) >> (64 - kScanValShift)) __trap();
    if (threadIdx.x == 0) vs[tile] = ((unsigned long long)(prefix + total) << kScanValShift) | tag | 2ull;
  }
  if (threadIdx.x == 0 && tile == gridDim.x - 1 && total_out) *total_out = prefix + total;
  if (out) {
    ex += prefix;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
      const uint64_t i = base + j;
      if (i < n) out[i] = ex;
      ex += v[j];
    }
  }
  if (threadIdx.x == 0) { // the last block to finish leaves the counters at zero for the next scan on the stream
    __threadfence();
    if (atomicAdd(&counters[1], 1u) == gridDim.x - 1) {
      counters[0] = 0;
      counters[1] = 0;
    }
  }
}

// out[i] = sum_{j<i} in(j) for i < n; *total_out (device) = sum of all.  `out` may be null (total only).
// n == 0 is allowed (total = 0).
template <class T, class InFn>
void exclusive_scan(Arena &arena, cudaStream_t st, InFn in, uint64_t n, T *out, T *total_out) {
  if (n == 0) {
    if (total_out) MSGPU_CUDA(cudaMemsetAsync(total_out, 0, sizeof(T), st));
    return;
  }
  const uint64_t      n_tiles = (n + kScanTile - 1) / kScanTile;
  unsigned            epoch   = 0;
  unsigned long long *state   = static_cast<unsigned long long *>(arena.scan_state(n_tiles * sizeof(unsigned long long), st, &epoch));
  MSGPU_LAUNCH("scan", st, scan_onepass_kernel<T, InFn><<<(unsigned)n_tiles, kScanBlock, 0, st>>>(in, n, out, total_out, state, epoch, arena.counter() + 2));
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
