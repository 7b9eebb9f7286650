// coverage.cu — the per-unitig coverage profile of the pipeline's unitig filter.
//
// Replaces the two passes of pipeline/unitig_filter.py over the PAF (the script decides which unitigs — and which
// stretches of over-covered unitigs — reach the assembler):
//   pass 1  unitig_filter.py:34-75    per BLOCK of consecutive lines naming one unitig: the number of distinct reads,
//                                     the unitig length, and the maximum of a per-position coverage profile in which a
//                                     read counts once (its first line inside the block)
//   pass 2  unitig_filter.py:102-176  for the blocks of unitigs whose maximum is above the cut-off: the profile of ALL
//                                     lines of the block, cut into the maximal runs of positions with coverage <= q3
//                                     that are at least 500 long (cut_peaks)
// The percentiles between the passes (:83-97) and the FASTA formatting stay on the host (muchsalsa_b200/prefilter.py).
//
// The script walks python lists position by position.  Here a block owns a run of (length + 1) slots in two flat
// counter arrays — interval starts and interval ends — filled by atomics; two prefix sums in ONE launch turn them into
// running totals whose difference is the coverage (both are monotone, so the scan's unsigned packing holds; a signed
// difference array would not), and because every interval ends inside its block the difference returns to zero at
// every block boundary: no segmented scan is needed.  One warp per block then takes the maximum / marks the runs.
#include <cstring>

#include "../../include/msgpu.h"
#include "internal.h"
#include "scan.cuh"

namespace msgpu {

// per (block, read) the record with the smallest line: the read's first line inside the block (unitig_filter.py:58-62)
struct BlockReadTable {
  unsigned long long *keys;
  uint32_t           *minline;
  uint64_t            mask;
};

__global__ void __launch_bounds__(256)
    cov_first_hit_insert_kernel(MatchSoA m, const uint32_t *__restrict__ block, BlockReadTable t,
                                uint32_t *__restrict__ slot) {
  pdl_sync();
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < m.n; i += uint64_t(gridDim.x) * blockDim.x) {
    const unsigned long long key = ((unsigned long long)block[i] << 32 | m.read[i]) + 1ull;
    uint64_t                 q   = mix64(key) & t.mask;
    for (;;) {
      const unsigned long long prev = atomicCAS(&t.keys[q], 0ull, key);
      if (prev == 0ull || prev == key) break;
      q = (q + 1) & t.mask;
    }
    atomicMin(&t.minline[q], m.line[i]);
    slot[i] = uint32_t(q);
  }
}

// block_len = the unitig length of the block's first line (the script sizes the profile from it, unitig_filter.py:67)
__global__ void __launch_bounds__(256)
    cov_block_len_kernel(MatchSoA m, const uint32_t *__restrict__ block, uint32_t *__restrict__ block_len) {
  pdl_sync();
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < m.n; i += uint64_t(gridDim.x) * blockDim.x)
    if (i == 0 || block[i] != block[i - 1]) block_len[block[i]] = uint32_t(m.score[i]);
}

struct LenPlusOne {
  const uint32_t *len;
  __device__ uint64_t operator()(uint64_t b) const { return uint64_t(len[b]) + 1ull; }
};

// interval events.  kPass1: only a read's first line inside its block counts, and the block's distinct reads are
// counted; pass 2: every line of the blocks flagged in `outlier`.  A line that reaches beyond ITS BLOCK's profile
// (another line of the block declared a longer unitig) is an IndexError in the script: flagged.
template <bool kPass1>
__global__ void __launch_bounds__(256)
    cov_events_kernel(MatchSoA m, const uint32_t *__restrict__ block, const uint64_t *__restrict__ block_off,
                      const uint32_t *__restrict__ block_len, BlockReadTable t, const uint32_t *__restrict__ slot,
                      const uint8_t *__restrict__ outlier, uint32_t *__restrict__ starts, uint32_t *__restrict__ ends,
                      uint32_t *__restrict__ block_n_reads, unsigned int *__restrict__ bad) {
  pdl_sync();
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < m.n; i += uint64_t(gridDim.x) * blockDim.x) {
    const uint32_t b = block[i];
    if (kPass1) {
      if (t.minline[slot[i]] != m.line[i]) continue;
      atomicAdd(&block_n_reads[b], 1u);
    } else if (!outlier[b]) {
      continue;
    }
    const int a = m.ir0[i], e = m.ir1[i];
    if (e >= int(block_len[b])) {
      *bad = 1;
      continue;
    }
    if (e < a) continue; // range(a, e + 1) is empty
    atomicAdd(&starts[block_off[b] + uint64_t(a)], 1u);
    atomicAdd(&ends[block_off[b] + uint64_t(e) + 1ull], 1u);
  }
}

// After the two prefix sums (exclusive, in place, over one slot more than there are positions):
//   coverage(block b, position p) = starts[off + p + 1] - ends[off + p + 1]
// One warp per block.  kCut = false: the maximum.  kCut = true (outlier blocks only): run starts are flagged.
template <bool kCut>
__global__ void __launch_bounds__(256)
    cov_block_scan_kernel(uint64_t n_blocks, const uint64_t *__restrict__ block_off, const uint32_t *__restrict__ block_len,
                          const uint32_t *__restrict__ starts, const uint32_t *__restrict__ ends,
                          uint32_t *__restrict__ block_max, const uint8_t *__restrict__ outlier, double q3,
                          uint8_t *__restrict__ run_start_flag) {
  pdl_sync();
  const unsigned l  = threadIdx.x & 31;
  const uint64_t nw = (uint64_t(gridDim.x) * blockDim.x) >> 5;
  for (uint64_t b = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; b < n_blocks; b += nw) {
    const uint64_t off = block_off[b];
    const uint32_t len = block_len[b];
    if (!kCut) {
      uint32_t mx = 0;
      for (uint32_t p = l; p < len; p += 32) {
        const uint32_t c = starts[off + p + 1] - ends[off + p + 1];
        mx               = c > mx ? c : mx;
      }
      for (int d = 16; d > 0; d >>= 1) {
        const uint32_t o = __shfl_xor_sync(0xffffffffu, mx, d);
        mx               = o > mx ? o : mx;
      }
      if (l == 0) block_max[b] = mx;
    } else {
      if (!outlier[b]) continue;
      for (uint32_t p = l; p < len; p += 32) { // a run of positions with coverage <= q3 starts here (cut_peaks)
        const bool below = double(starts[off + p + 1] - ends[off + p + 1]) <= q3;
        const bool prev  = p > 0 && double(starts[off + p] - ends[off + p]) <= q3;
        run_start_flag[off + p] = (below && !prev) ? 1 : 0;
      }
    }
  }
}

// every run writes its block and start at its index; the position where it ends writes the end (one warp per block)
__global__ void __launch_bounds__(256)
    cov_runs_kernel(uint64_t n_blocks, const uint64_t *__restrict__ block_off, const uint32_t *__restrict__ block_len,
                    const uint32_t *__restrict__ starts, const uint32_t *__restrict__ ends,
                    const uint8_t *__restrict__ outlier, double q3, const uint8_t *__restrict__ run_start_flag,
                    const uint32_t *__restrict__ run_index, uint32_t *__restrict__ run_block,
                    int32_t *__restrict__ run_start, int32_t *__restrict__ run_end) {
  pdl_sync();
  const unsigned l  = threadIdx.x & 31;
  const uint64_t nw = (uint64_t(gridDim.x) * blockDim.x) >> 5;
  for (uint64_t b = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; b < n_blocks; b += nw) {
    if (!outlier[b]) continue;
    const uint64_t off = block_off[b];
    const uint32_t len = block_len[b];
    for (uint32_t p = l; p < len; p += 32) {
      const bool below = double(starts[off + p + 1] - ends[off + p + 1]) <= q3;
      if (!below) continue;
      const uint32_t k = run_index[off + p] + (run_start_flag[off + p] ? 1u : 0u) - 1u; // the run this position is in
      if (run_start_flag[off + p]) {
        run_block[k] = uint32_t(b);
        run_start[k] = int32_t(p);
      }
      const bool next = p + 1 < len && double(starts[off + p + 2] - ends[off + p + 2]) <= q3;
      if (!next) run_end[k] = int32_t(p);
    }
  }
}

struct LongRun { // cut_peaks keeps a run iff it is at least 500 positions long (unitig_filter.py:113,122)
  const int32_t *start, *end;
  __device__ uint32_t operator()(uint64_t k) const { return end[k] - start[k] + 1 >= 500 ? 1u : 0u; }
};

__global__ void __launch_bounds__(256)
    cov_keep_runs_kernel(uint64_t n_runs, LongRun keep, const uint32_t *__restrict__ pos, const uint32_t *__restrict__ run_block,
                         uint32_t *__restrict__ frag_block, int32_t *__restrict__ frag_start, int32_t *__restrict__ frag_end) {
  pdl_sync();
  for (uint64_t k = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < n_runs; k += uint64_t(gridDim.x) * blockDim.x) {
    if (!keep(k)) continue;
    const uint32_t o = pos[k];
    frag_block[o]    = run_block[k];
    frag_start[o]    = keep.start[k];
    frag_end[o]      = keep.end[k];
  }
}

// the LAST block of every unitig decides (mapping_freq is a dict keyed by the unitig: later blocks overwrite)
__global__ void __launch_bounds__(256)
    cov_last_block_kernel(uint64_t n_blocks, const uint32_t *__restrict__ block_unitig, uint32_t *__restrict__ last_block) {
  pdl_sync();
  for (uint64_t b = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; b < n_blocks; b += uint64_t(gridDim.x) * blockDim.x)
    atomicMax(&last_block[block_unitig[b]], uint32_t(b));
}
__global__ void __launch_bounds__(256)
    cov_outlier_kernel(uint64_t n_blocks, const uint32_t *__restrict__ block_unitig, const uint32_t *__restrict__ last_block,
                       const uint32_t *__restrict__ block_max, double cutoff, uint8_t *__restrict__ outlier) {
  pdl_sync();
  for (uint64_t b = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; b < n_blocks; b += uint64_t(gridDim.x) * blockDim.x)
    outlier[b] = double(block_max[last_block[block_unitig[b]]]) > cutoff ? 1 : 0;
}

template <class T> static T cov_read_back(Arena &arena, cudaStream_t st, const T *d) {
  void *h = arena.host_scratch();
  MSGPU_CUDA(cudaMemcpyAsync(h, d, sizeof(T), cudaMemcpyDeviceToHost, st));
  MSGPU_CUDA(cudaStreamSynchronize(st));
  T out;
  memcpy(&out, h, sizeof(T));
  return out;
}

// the two running totals of one pass, in place (exclusive, over n_positions + 1 slots), in ONE launch
static void cov_prefix_sums(Arena &arena, cudaStream_t st, CoverageTables &c) {
  run_scans(arena, st,
            scan_job<uint32_t>(arena, st, 0, LoadAs<uint32_t, uint32_t>{c.starts}, c.n_positions + 1, c.starts, (uint32_t *)nullptr),
            scan_job<uint32_t>(arena, st, 1, LoadAs<uint32_t, uint32_t>{c.ends}, c.n_positions + 1, c.ends, (uint32_t *)nullptr));
}

void run_unitig_coverage(Arena &arena, cudaStream_t st, const MatchSoA &m, uint64_t n_unitigs, CoverageTables &c) {
  c = CoverageTables{};
  const uint64_t M = m.n;
  c.n_unitigs      = n_unitigs;
  c.n_blocks       = run_scrubber_blocks(arena, st, m, &c.rec_block, &c.block_unitig); // (same block rule as the scrubber)
  const uint64_t B = c.n_blocks;
  c.block_len     = arena.alloc<uint32_t>(B + 1);
  c.block_n_reads = arena.alloc<uint32_t>(B + 1);
  c.block_max     = arena.alloc<uint32_t>(B + 1);
  c.block_off     = arena.alloc<uint64_t>(B + 1);
  unsigned long long *d_total = arena.alloc<unsigned long long>(1);
  unsigned int       *d_bad   = arena.alloc<unsigned int>(1);
  MSGPU_CUDA(cudaMemsetAsync(c.block_n_reads, 0, (B + 1) * sizeof(uint32_t), st));
  MSGPU_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(unsigned int), st));
  const int gm = grid_for(M, 256);
  if (M) MSGPU_LAUNCH("cov_block_len", st, cov_block_len_kernel<<<gm, 256, 0, st>>>(m, c.rec_block, c.block_len));
  exclusive_scan<uint64_t>(arena, st, LenPlusOne{c.block_len}, B, c.block_off, reinterpret_cast<uint64_t *>(d_total));
  c.n_positions = cov_read_back(arena, st, d_total); // sizes the two counter arrays
  if (c.n_positions >= 0xfffffff0ull) throw BuildError{MSGPU_ERR_LIMIT, "unitig coverage: more than 2^32 profile positions per device"};
  MSGPU_CUDA(cudaMemcpyAsync(c.block_off + B, d_total, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
  c.starts = arena.alloc<uint32_t>(c.n_positions + 2);
  c.ends   = arena.alloc<uint32_t>(c.n_positions + 2);
  MSGPU_CUDA(cudaMemsetAsync(c.starts, 0, (c.n_positions + 2) * sizeof(uint32_t), st));
  MSGPU_CUDA(cudaMemsetAsync(c.ends, 0, (c.n_positions + 2) * sizeof(uint32_t), st));
  // a read counts once per block: its first line there
  uint64_t cap = 1024;
  while (cap < 2 * M + 2) cap <<= 1;
  c.first_hit.keys    = arena.alloc<unsigned long long>(cap);
  c.first_hit.minline = arena.alloc<uint32_t>(cap);
  c.first_hit.mask    = cap - 1;
  c.slot              = arena.alloc<uint32_t>(M + 1);
  MSGPU_CUDA(cudaMemsetAsync(c.first_hit.keys, 0, cap * sizeof(unsigned long long), st));
  MSGPU_CUDA(cudaMemsetAsync(c.first_hit.minline, 0xff, cap * sizeof(uint32_t), st));
  BlockReadTable t{c.first_hit.keys, c.first_hit.minline, c.first_hit.mask};
  if (M) {
    MSGPU_LAUNCH("cov_first_hit", st, cov_first_hit_insert_kernel<<<gm, 256, 0, st>>>(m, c.rec_block, t, c.slot));
    MSGPU_LAUNCH("cov_events", st, cov_events_kernel<true><<<gm, 256, 0, st>>>(m, c.rec_block, c.block_off, c.block_len, t, c.slot, nullptr, c.starts, c.ends, c.block_n_reads, d_bad));
  }
  cov_prefix_sums(arena, st, c);
  if (B) MSGPU_LAUNCH("cov_block_max", st, cov_block_scan_kernel<false><<<grid_for(B * 32, 256, 8), 256, 0, st>>>(B, c.block_off, c.block_len, c.starts, c.ends, c.block_max, nullptr, 0.0, nullptr));
  if (cov_read_back(arena, st, d_bad))
    throw BuildError{MSGPU_ERR_INVALID_PAF, "a line reaches beyond the unitig length its block started with (IndexError in unitig_filter.py)"};
  c.m     = m;
  c.valid = true;
}

void run_unitig_cut(Arena &arena, cudaStream_t st, CoverageTables &c, double cutoff, double q3) {
  if (!c.valid) throw BuildError{MSGPU_ERR_NO_RESULT, "no unitig coverage to cut"};
  const uint64_t B = c.n_blocks, M = c.m.n, T = c.n_positions;
  uint32_t *last_block = arena.alloc<uint32_t>(c.n_unitigs + 1);
  c.block_outlier      = arena.alloc<uint8_t>(B + 1);
  uint8_t  *run_flag   = arena.alloc<uint8_t>(T + 2);
  uint32_t *run_index  = arena.alloc<uint32_t>(T + 2);
  uint32_t *d_n        = arena.alloc<uint32_t>(2);
  unsigned int *d_bad  = arena.alloc<unsigned int>(1);
  MSGPU_CUDA(cudaMemsetAsync(last_block, 0, (c.n_unitigs + 1) * sizeof(uint32_t), st));
  MSGPU_CUDA(cudaMemsetAsync(run_flag, 0, T + 2, st));
  MSGPU_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(unsigned int), st));
  MSGPU_CUDA(cudaMemsetAsync(c.starts, 0, (T + 2) * sizeof(uint32_t), st));
  MSGPU_CUDA(cudaMemsetAsync(c.ends, 0, (T + 2) * sizeof(uint32_t), st));
  const int gb = grid_for(B, 256), gm = grid_for(M, 256), gw = grid_for(B * 32, 256, 8);
  BlockReadTable t{c.first_hit.keys, c.first_hit.minline, c.first_hit.mask};
  if (B) {
    MSGPU_LAUNCH("cov_last_block", st, cov_last_block_kernel<<<gb, 256, 0, st>>>(B, c.block_unitig, last_block));
    MSGPU_LAUNCH("cov_outlier", st, cov_outlier_kernel<<<gb, 256, 0, st>>>(B, c.block_unitig, last_block, c.block_max, cutoff, c.block_outlier));
  }
  if (M) MSGPU_LAUNCH("cov_events", st, cov_events_kernel<false><<<gm, 256, 0, st>>>(c.m, c.rec_block, c.block_off, c.block_len, t, c.slot, c.block_outlier, c.starts, c.ends, nullptr, d_bad));
  cov_prefix_sums(arena, st, c);
  if (B) MSGPU_LAUNCH("cov_run_starts", st, cov_block_scan_kernel<true><<<gw, 256, 0, st>>>(B, c.block_off, c.block_len, c.starts, c.ends, nullptr, c.block_outlier, q3, run_flag));
  exclusive_scan<uint32_t>(arena, st, LoadAs<uint8_t, uint32_t>{run_flag}, T, run_index, d_n);
  const uint64_t n_runs = T ? cov_read_back(arena, st, d_n) : 0;
  uint32_t *run_block = arena.alloc<uint32_t>(n_runs + 1);
  int32_t  *run_start = arena.alloc<int32_t>(n_runs + 1), *run_end = arena.alloc<int32_t>(n_runs + 1);
  if (B && n_runs) MSGPU_LAUNCH("cov_runs", st, cov_runs_kernel<<<gw, 256, 0, st>>>(B, c.block_off, c.block_len, c.starts, c.ends, c.block_outlier, q3, run_flag, run_index, run_block, run_start, run_end));
  uint32_t *pos = arena.alloc<uint32_t>(n_runs + 1);
  LongRun   keep{run_start, run_end};
  exclusive_scan<uint32_t>(arena, st, keep, n_runs, pos, d_n + 1);
  c.n_frags    = n_runs ? cov_read_back(arena, st, d_n + 1) : 0;
  c.frag_block = arena.alloc<uint32_t>(c.n_frags + 1);
  c.frag_start = arena.alloc<int32_t>(c.n_frags + 1);
  c.frag_end   = arena.alloc<int32_t>(c.n_frags + 1);
  if (n_runs) MSGPU_LAUNCH("cov_keep_runs", st, cov_keep_runs_kernel<<<grid_for(n_runs, 256), 256, 0, st>>>(n_runs, keep, pos, run_block, c.frag_block, c.frag_start, c.frag_end));
  MSGPU_CUDA(cudaStreamSynchronize(st));
  c.cut_valid = true;
}

} // namespace msgpu
