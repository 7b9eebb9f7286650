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
// The look-back window is the whole block (512 predecessors per round trip) and a tile holds 8192 items, so the scans of
// this path (at most a few hundred tiles) resolve in one round trip after their own reduce; pollers back off with
// nanosleep so that they do not crowd out the tiles that are still reducing.
// `live` (optional): a count in device memory; items at index >= live contribute nothing and their outputs are not
// written, tiles entirely behind it retire at once.  Lets a scan be launched over a capacity bound when the actual
// count is only known on the device.
#pragma once

#include "common.cuh"

namespace msgpu {

constexpr int      kScanBlock     = 512;
constexpr int      kScanItems     = 16; // contiguous items per thread
constexpr int      kScanTile      = kScanBlock * kScanItems;
constexpr int      kScanValShift  = 18;
constexpr unsigned kScanEpochMask = 0xffffu;

struct ScanLive { // live = min(n, *p32 or *p64 (whichever is set) + extra); both null = n
  const unsigned           *p32   = nullptr;
  const unsigned long long *p64   = nullptr;
  uint64_t                  extra = 0;
};

template <class T, class InFn>
__global__ void __launch_bounds__(kScanBlock)
    scan_onepass_kernel(InFn in, uint64_t n_cap, ScanLive live, T *out /* may alias the input; may be null */,
                        T *__restrict__ total_out, unsigned long long *state, unsigned epoch,
                        unsigned *counters /* ring of ticket counters, one per epoch */) {
  constexpr int       kWarps = kScanBlock / 32;
  __shared__ unsigned s_tile;
  __shared__ uint64_t s_n;
  __shared__ T        w_tot[kWarps]; // warp totals of the tile, then their exclusive prefix
  __shared__ T        w_sum[kWarps];
  __shared__ unsigned w_inc[kWarps];
  const unsigned l = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    // the ticket counter of this scan; the slot half a ring ahead is cleared for a later scan (scans on one stream
    // are ordered, so nobody is using it)
    s_tile = atomicAdd(&counters[epoch & 31u], 1u);
    if (blockIdx.x == 0) counters[(epoch + 16u) & 31u] = 0u;
    uint64_t n = n_cap;
    if (live.p32 || live.p64) {
      const uint64_t c = (live.p32 ? uint64_t(*live.p32) : uint64_t(*live.p64)) + live.extra;
      n                = c < n_cap ? c : n_cap;
    }
    s_n = n;
  }
  __syncthreads();
  const unsigned tile      = s_tile;
  const uint64_t n         = s_n;
  const unsigned last_tile = n ? unsigned((n - 1) / kScanTile) : 0u; // the tile that reports the total
  if (tile > last_tile) return;
  // a warp owns 512 consecutive items; lane l takes items l, l + 32, ... of them, so every load and store of the warp
  // is one contiguous run
  const uint64_t wbase = uint64_t(tile) * kScanTile + uint64_t(w) * (32 * kScanItems);
  T              v[kScanItems];
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    const uint64_t i = wbase + uint64_t(j) * 32 + l;
    v[j]             = i < n ? in(i) : T(0);
  }
  T carry = 0; // exclusive prefix inside the warp's run
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    const T incl = warp_inclusive_scan(v[j]);
    const T row  = __shfl_sync(0xffffffffu, incl, 31);
    v[j]         = carry + incl - v[j];
    carry += row;
  }
  if (l == 0) w_tot[w] = carry;
  __syncthreads();
  T total = 0, wex = 0;
#pragma unroll
  for (int ww = 0; ww < kWarps; ++ww) {
    const T c = w_tot[ww];
    wex += ww < int(w) ? c : T(0);
    total += c;
  }
  volatile unsigned long long *vs  = state;
  const unsigned long long     tag = (unsigned long long)(epoch & kScanEpochMask) << 2;
  if ((unsigned long long)total >> (64 - kScanValShift)) __trap();
  if (threadIdx.x == 0) vs[tile] = ((unsigned long long)total << kScanValShift) | tag | (tile == 0 ? 2ull : 1ull);
  T prefix = 0;
  if (tile > 0) {
    long long hi = (long long)tile - 1; // window = tiles hi, hi-1, ..., hi-(kScanBlock-1)
    while (true) {
      const long long    p    = hi - (long long)threadIdx.x;
      unsigned long long word = 0;
      if (p >= 0) {
        word = vs[p];
        while ((word & (kScanEpochMask << 2)) != tag || (word & 3ull) == 0ull) {
          __nanosleep(32);
          word = vs[p];
        }
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
      for (int ww = 0; ww < kWarps; ++ww) {
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
  if (threadIdx.x == 0 && tile == last_tile && total_out) *total_out = prefix + total;
  if (out) {
    const T off = prefix + wex;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
      const uint64_t i = wbase + uint64_t(j) * 32 + l;
      if (i < n) out[i] = off + v[j];
    }
  }
}

// out[i] = sum_{j<i} in(j) for i < n; *total_out (device) = sum of all.  `out` may be null (total only).
// n == 0 is allowed (total = 0).  With `live`, see above: outputs exist for i < live only.
template <class T, class InFn>
void exclusive_scan(Arena &arena, cudaStream_t st, InFn in, uint64_t n, T *out, T *total_out, ScanLive live = ScanLive{}) {
  if (n == 0) {
    if (total_out) MSGPU_CUDA(cudaMemsetAsync(total_out, 0, sizeof(T), st));
    return;
  }
  const uint64_t      n_tiles = (n + kScanTile - 1) / kScanTile;
  unsigned            epoch   = 0;
  unsigned long long *state   = static_cast<unsigned long long *>(arena.scan_state(n_tiles * sizeof(unsigned long long), st, &epoch));
  MSGPU_LAUNCH("scan", st, scan_onepass_kernel<T, InFn><<<(unsigned)n_tiles, kScanBlock, 0, st>>>(in, n, live, out, total_out, state, epoch, arena.counter() + 8));
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
