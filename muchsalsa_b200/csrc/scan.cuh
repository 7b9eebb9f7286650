// scan.cuh — device-wide exclusive prefix sum in ONE launch (single pass, decoupled look-back), hand-written.
//
// The arrays scanned on this path are per-line, per-match, per-read or per-unitig: one to a few hundred tiles, so what
// a scan costs is its latency, not its bytes.  One kernel: a block takes a ticket (tiles are numbered in the order
// blocks start, so a tile only ever waits for tiles that are already running), reduces its items (512 threads x 2, 8
// or 16 items by the length of the scan; a warp owns consecutive ones, lane-striped, so every load and store is one
// contiguous run), publishes the tile aggregate,
// looks back over its predecessors' aggregates / inclusive prefixes, publishes its own inclusive prefix and writes
// its outputs.  Every input value is evaluated once.  Input values come from a functor so flags packed in bytes /
// derived counts need no staging array.  Up to four independent scans share one launch (blockIdx.y picks the scan).
//
// The per-tile status words live in a region of the arena that outlives reset() and is used by scans only.  A status
// is ONE 64-bit word — value << 18 | epoch << 2 | state (1: tile aggregate, 2: inclusive prefix) — so a single load
// observes value and state together and no fence is needed.  The epoch is the serial number of the scan that wrote
// the word: stale words read as "not ready" and no launch is spent on clearing them.  Sums must stay below 2^46
// (7e13; the largest quantity scanned here is the pair count, ~1e11 at human scale); a larger one traps.
// The look-back window is the whole block (512 predecessors per round trip), so the scans of this path (at most a few
// hundred tiles) resolve in one or two round trips after their own reduce.
// `live` (optional): a count in device memory; items at index >= live contribute nothing and their outputs are not
// written, tiles entirely behind it retire at once.  Lets a scan be launched over a capacity bound when the actual
// count is only known on the device.
#pragma once

#include "common.cuh"

namespace msgpu {

constexpr int      kScanBlock     = 512;
constexpr int      kScanItems     = 16; // items per thread (lane-striped inside the warp's run) of a LONG scan
constexpr int      kScanTile      = kScanBlock * kScanItems;
// A thread evaluates its items one after the other, and the input functors of this path are chains of dependent loads
// (a row's weight, the prefix before it, ...): at 16 items per thread a scan over 7e4 reads is nine blocks that each
// wait for 16 round trips to L2 (ncu: three quarters of the stall samples of such a scan sit on the functor's loads).
// Short scans therefore take 2 (or 8) items per thread — as many tiles as fit the GPU in one wave (4 blocks per SM).
constexpr uint64_t kScanWaveBlocks = 148 * 4;
inline int scan_items_for(uint64_t n) {
  if (n <= kScanWaveBlocks * kScanBlock * 2) return 2;
  if (n <= kScanWaveBlocks * kScanBlock * 8) return 8;
  return kScanItems;
}
constexpr int      kScanValShift  = 18;
constexpr unsigned kScanEpochMask = 0xffffu;

struct ScanLive { // live = min(n, *p32 or *p64 (whichever is set) + extra); both null = n
  const unsigned           *p32   = nullptr;
  const unsigned long long *p64   = nullptr;
  uint64_t                  extra = 0;
};

// One scan of a launch: what to scan and where its status words and ticket counter live
template <class T, class InFn> struct ScanJob {
  using value_type = T;
  InFn                in;
  uint64_t            n_cap;
  ScanLive            live;
  T                  *out;       // may alias the input; may be null
  T                  *total_out; // device; may be null
  unsigned long long *state;
  unsigned            epoch;
  unsigned           *ticket;      // the slot's ticket counter (grows forever) ...
  unsigned            ticket_base; // ... and its value when this scan starts (host bookkeeping, Arena::scan_state)
  int                 items;       // items per thread: 2, 8 or kScanItems (scan_items_for)
  unsigned            n_tiles;     // tiles of n_cap at that size
};

// the work of one block of one scan; blocks past the scan's tiles (a launch is sized for its longest scan) leave
template <int kItems, class Job> __device__ __forceinline__ void scan_block_n(const Job &job) {
  using T = typename Job::value_type;
  constexpr int       kWarps = kScanBlock / 32;
  constexpr int       kTile  = kScanBlock * kItems;
  __shared__ unsigned s_tile;
  __shared__ uint64_t s_n;
  __shared__ T        w_tot[kWarps]; // warp totals of the tile
  __shared__ T        w_sum[kWarps];
  __shared__ unsigned w_inc[kWarps];
  const uint64_t n_tiles_cap = job.n_tiles;
  if (blockIdx.x >= n_tiles_cap) return;
  const unsigned epoch = job.epoch;
  const unsigned l = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    // tiles are numbered in the order blocks start; every block of a multi-tile scan takes exactly one ticket, which
    // is what the host adds to the slot's base for the next scan.  A one-tile scan needs no ticket.
    s_tile = n_tiles_cap > 1 ? atomicAdd(job.ticket, 1u) - job.ticket_base : 0u;
    uint64_t n = job.n_cap;
    if (job.live.p32 || job.live.p64) {
      const uint64_t c = (job.live.p32 ? uint64_t(*job.live.p32) : uint64_t(*job.live.p64)) + job.live.extra;
      n                = c < job.n_cap ? c : job.n_cap;
    }
    s_n = n;
  }
  __syncthreads();
  const unsigned tile      = s_tile;
  const uint64_t n         = s_n;
  const unsigned last_tile = n ? unsigned((n - 1) / kTile) : 0u; // the tile that reports the total
  if (tile > last_tile) return;
  // a warp owns 512 consecutive items; lane l takes items l, l + 32, ... of them, so every load and store of the warp
  // is one contiguous run
  const uint64_t wbase = uint64_t(tile) * kTile + uint64_t(w) * (32 * kItems);
  T              v[kItems];
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    const uint64_t i = wbase + uint64_t(j) * 32 + l;
    v[j]             = i < n ? job.in(i) : T(0);
  }
  T carry = 0; // exclusive prefix inside the warp's run
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
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
  T prefix = 0;
  if (last_tile > 0) { // several live tiles: publish, look back
    volatile unsigned long long *vs  = job.state;
    const unsigned long long     tag = (unsigned long long)(epoch & kScanEpochMask) << 2;
    if ((unsigned long long)total >> (64 - kScanValShift)) __trap();
    if (threadIdx.x == 0) vs[tile] = ((unsigned long long)total << kScanValShift) | tag | (tile == 0 ? 2ull : 1ull);
    if (tile > 0) {
      long long hi = (long long)tile - 1; // window = tiles hi, hi-1, ..., hi-(kScanBlock-1)
      while (true) {
        const long long    p    = hi - (long long)threadIdx.x;
        unsigned long long word = 0;
        if (p >= 0) {
          word = vs[p];
          while ((word & (kScanEpochMask << 2)) != tag || (word & 3ull) == 0ull) {
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
  }
  if (threadIdx.x == 0 && tile == last_tile && job.total_out) *job.total_out = prefix + total;
  if (job.out) {
    const T off = prefix + wex;
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
      const uint64_t i = wbase + uint64_t(j) * 32 + l;
      if (i < n) job.out[i] = off + v[j];
    }
  }
}

template <class Job> __device__ __forceinline__ void scan_block(const Job &job) {
  if (job.items == 2) scan_block_n<2>(job);
  else if (job.items == 8) scan_block_n<8>(job);
  else scan_block_n<kScanItems>(job);
}

// up to four independent scans in ONE launch (blockIdx.y picks the scan): they cost the latency of one
template <class J0> __global__ void __launch_bounds__(kScanBlock) scan_kernel(J0 j0) {
  pdl_sync();
  scan_block(j0);
}
template <class J0, class J1> __global__ void __launch_bounds__(kScanBlock) scan_kernel(J0 j0, J1 j1) {
  pdl_sync();
  if (blockIdx.y == 0) scan_block(j0); else scan_block(j1);
}
template <class J0, class J1, class J2, class J3>
__global__ void __launch_bounds__(kScanBlock) scan_kernel(J0 j0, J1 j1, J2 j2, J3 j3) {
  pdl_sync();
  if (blockIdx.y == 0) scan_block(j0);
  else if (blockIdx.y == 1) scan_block(j1);
  else if (blockIdx.y == 2) scan_block(j2);
  else scan_block(j3);
}

// Prepares a job: out[i] = sum_{j<i} in(j) for i < n; *total_out (device) = sum of all.  `out` may be null (total
// only).  With `live`, see above: outputs exist for i < live only.  `slot` = which of the (up to four) concurrent
// scans of the coming launch this is; each gets its own status words and epoch.
template <class T, class InFn>
ScanJob<T, InFn> scan_job(Arena &arena, cudaStream_t st, int slot, InFn in, uint64_t n, T *out, T *total_out,
                          ScanLive live = ScanLive{}) {
  const int      items   = scan_items_for(n);
  const uint64_t n_tiles = (n + uint64_t(kScanBlock) * items - 1) / (uint64_t(kScanBlock) * items);
  unsigned       epoch = 0, *ticket = nullptr, ticket_base = 0;
  // the launch covers this scan's tiles whenever it happens at all (run_scans sizes the grid for the longest job and
  // skips the launch only when every job is empty), so a multi-tile scan takes exactly n_tiles tickets
  auto *state = static_cast<unsigned long long *>(arena.scan_state(
      slot, n_tiles * sizeof(unsigned long long), st, &epoch, n_tiles > 1 ? unsigned(n_tiles) : 0u, &ticket, &ticket_base));
  return ScanJob<T, InFn>{in, n, live, out, total_out, state, epoch, ticket, ticket_base, items, unsigned(n_tiles)};
}

template <class J0> void run_scans(Arena &arena, cudaStream_t st, const J0 &j0) {
  if (!j0.n_cap) return;
  MSGPU_LAUNCH("scan", st, launch_k(scan_kernel<J0>, j0.n_tiles, kScanBlock, 0, st, j0));
}
template <class J0, class J1> void run_scans(Arena &arena, cudaStream_t st, const J0 &j0, const J1 &j1) {
  const unsigned n = j0.n_tiles > j1.n_tiles ? j0.n_tiles : j1.n_tiles; // (an empty scan has no tiles)
  if (!n) return;
  MSGPU_LAUNCH("scan", st, (launch_k(scan_kernel<J0, J1>, dim3(n, 2), kScanBlock, 0, st, j0, j1)));
}
template <class J0, class J1, class J2, class J3>
void run_scans(Arena &arena, cudaStream_t st, const J0 &j0, const J1 &j1, const J2 &j2, const J3 &j3) {
  unsigned n = j0.n_tiles;
  n          = j1.n_tiles > n ? j1.n_tiles : n;
  n          = j2.n_tiles > n ? j2.n_tiles : n;
  n          = j3.n_tiles > n ? j3.n_tiles : n;
  if (!n) return;
  MSGPU_LAUNCH("scan", st, (launch_k(scan_kernel<J0, J1, J2, J3>, dim3(n, 4), kScanBlock, 0, st, j0, j1, j2, j3)));
}

// a single scan.  n == 0 is allowed (total = 0).
template <class T, class InFn>
void exclusive_scan(Arena &arena, cudaStream_t st, InFn in, uint64_t n, T *out, T *total_out, ScanLive live = ScanLive{}) {
  if (n == 0) {
    if (total_out) MSGPU_CUDA(cudaMemsetAsync(total_out, 0, sizeof(T), st));
    return;
  }
  run_scans(arena, st, scan_job<T, InFn>(arena, st, 0, in, n, out, total_out, live));
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
