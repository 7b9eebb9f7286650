// msgpu_api.cu — the C ABI declared in include/msgpu.h: context, construction entry points, result marshalling.
//
// The host side of this path is C++ like the reference's; no exception crosses the boundary (the reference reports
// errors as std::runtime_error — "Can't open blast file." BlastFileAccessor.cpp:44, "Invalid BLAST file."
// BlastFileReader.cpp:98 — which become status codes + msgpu_last_error()).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <new>
#include <string>

#include "../../include/msgpu.h"
#include "internal.h"
#include "seq.h"

using namespace msgpu;

namespace {

struct PinnedBuf {
  void  *p   = nullptr;
  size_t cap = 0;
  void  *get(size_t bytes) {
    if (bytes > cap) {
      if (p) cudaFreeHost(p);
      p = nullptr;
      size_t want = bytes + bytes / 4 + 4096;
      MSGPU_CUDA(cudaMallocHost(&p, want));
      cap = want;
    }
    return p;
  }
  ~PinnedBuf() {
    if (p) cudaFreeHost(p);
  }
};

} // namespace

struct msgpu_ctx {
  int          device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  Arena        arena;       // per-build device memory
  Arena        input_arena; // staged PAF text
  std::string  error;

  bool          have_result = false;
  uint32_t      n_parts     = 1; // > 1: the edge tables of the last build are one slice of a multi-GPU build
  uint64_t      n_lines     = 0;
  ReadTable     reads;
  UnitigTable   unitigs;
  MatchSoA      matches;
  EdgeTables    edges;
  msgpu_timings timings{};
  cudaEvent_t   ev[6]{};

  PinnedBuf    host_result; // one pinned slab holding every host table
  PinnedBuf    staging;     // pinned staging for pageable PAF input
  bool         host_valid = false;
  msgpu_tables host_tables{};

  OrderTables  orders; // msgpu_chain_overlaps
  float        chain_ms = 0.f;
  PinnedBuf    host_orders;
  bool         orders_host_valid = false;
  msgpu_orders host_orders_view{};

  FilterView               filter_view; // msgpu_filter_view_describe
  FilterTables             filters; // msgpu_edge_filters
  float                    filter_ms = 0.f;
  PinnedBuf                host_filters;
  bool                     filters_host_valid = false;
  msgpu_edge_filter_tables host_filters_view{};

  // chunked emission (msgpu_set_chunk_tiles / msgpu_chunk_next)
  struct Chunks {
    bool        active = false; // the last build stopped after planning; chunks are being delivered
    PairPlan    plan;
    PairScratch sc[2];
    EdgeTables  out[2];
    uint32_t    next_tile = 0;         // first tile whose "tiles" half has not been launched
    int         cur = -1;              // buffer set whose tiles half is in flight and rows half is not (-1: none left)
    uint32_t    cur_lo = 0, cur_hi = 0;
    uint64_t    index = 0, em_base = 0, edge_base = 0;
    bool        taken = false;         // a chunk has been delivered to the caller
    EdgeTables  last;                  // the chunk last delivered (for msgpu_chunk_digest)
    bool        have_last = false;
    cudaEvent_t rows_done = nullptr;
  } chunks;
  uint32_t     chunk_tiles = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t  copy_done[2]{};
  bool         copy_pending[2] = {false, false}; // a download of that buffer set is in flight on copy_stream

  CoverageTables coverage; // msgpu_unitig_coverage_*
  PinnedBuf      host_coverage, host_fragments;

  SeqStore *seq = nullptr;
  Profiler  profiler;
};

namespace {

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

template <class Fn> int guarded(msgpu_ctx *ctx, Fn &&fn) {
  if (!ctx) return MSGPU_ERR_INVALID_ARG;
  struct ProfScope {
    explicit ProfScope(Profiler *p) { current_profiler() = p; }
    ~ProfScope() { current_profiler() = nullptr; }
  };
  try {
    DeviceGuard g(ctx->device);
    ProfScope   ps(&ctx->profiler);
    fn();
    ctx->error.clear();
    return MSGPU_OK;
  } catch (BuildError const &e) {
    ctx->error = e.message;
    return e.status;
  } catch (CudaError const &e) {
    ctx->error = e.what();
    return MSGPU_ERR_CUDA;
  } catch (std::bad_alloc const &) {
    ctx->error = "out of host memory";
    return MSGPU_ERR_NOMEM;
  } catch (std::exception const &e) {
    ctx->error = e.what();
    return MSGPU_ERR_INTERNAL;
  }
}

float elapsed(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

// pairing of a build: whole (resident) or, with chunking on, planned only — the chunks follow through
// msgpu_chunk_next / msgpu_result_host
void pair_or_plan(msgpu_ctx *ctx, uint32_t part, uint32_t n_parts) {
  cudaStream_t st = ctx->stream;
  {
    cudaEvent_t keep  = ctx->chunks.rows_done; // the event outlives the builds
    ctx->chunks       = msgpu_ctx::Chunks{};
    ctx->chunks.rows_done = keep;
  }
  if (!ctx->chunk_tiles) {
    run_pairing(ctx->arena, st, ctx->matches, ctx->unitigs, ctx->reads.n, ctx->edges, ctx->ev[3], part, n_parts);
    return;
  }
  auto &c = ctx->chunks;
  pairing_plan(ctx->arena, st, ctx->matches, ctx->unitigs, ctx->reads.n, part, n_parts, c.plan);
  ctx->edges           = EdgeTables{};
  ctx->edges.n_pairs   = c.plan.n_pairs;
  ctx->edges.pair_rows = c.plan.rows;
  ctx->edges.p_read    = c.plan.p_read;
  const uint32_t my_tiles = c.plan.tile_hi - c.plan.tile_lo;
  const uint32_t per      = ctx->chunk_tiles < my_tiles ? ctx->chunk_tiles : (my_tiles ? my_tiles : 1u);
  const uint64_t cap      = pairing_rows_bound(c.plan, per);
  for (int b = 0; b < 2; ++b) {
    pairing_scratch_alloc(ctx->arena, per, c.sc[b]);
    pairing_outputs_alloc(ctx->arena, cap, c.out[b]);
  }
  MSGPU_CUDA(cudaEventRecord(ctx->ev[3], st));
  c.active    = true;
  c.next_tile = c.plan.tile_lo;
  // the tiles half of the first chunk starts right away
  c.cur_lo = c.next_tile;
  c.cur_hi = c.cur_lo + per < c.plan.tile_hi ? c.cur_lo + per : c.plan.tile_hi;
  pairing_tiles(ctx->arena, st, c.plan, c.cur_lo, c.cur_hi, c.sc[0]);
  c.next_tile = c.cur_hi;
  c.cur       = 0;
}

// delivers the next chunk into `e` (device tables of the chunk); false when there is none left
bool next_chunk(msgpu_ctx *ctx, EdgeTables &e, uint64_t *index, bool *last) {
  auto &c = ctx->chunks;
  if (!c.active || c.cur < 0) return false;
  cudaStream_t st = ctx->stream;
  if (!ctx->chunks.rows_done) MSGPU_CUDA(cudaEventCreateWithFlags(&ctx->chunks.rows_done, cudaEventDisableTiming));
  const int      b  = c.cur;
  const uint32_t lo = c.cur_lo, hi = c.cur_hi;
  if (ctx->copy_pending[b]) { // msgpu_result_host is still downloading the chunk that used this buffer set
    MSGPU_CUDA(cudaStreamWaitEvent(st, ctx->copy_done[b], 0));
    ctx->copy_pending[b] = false;
  }
  pairing_rows(ctx->arena, st, c.plan, lo, hi, c.sc[b], c.out[b], c.em_base);
  // the counts of this chunk travel to pinned memory behind its rows; the tiles half of the NEXT chunk is queued
  // before the host waits, so the GPU keeps sorting while the caller works on this chunk
  char *h = static_cast<char *>(ctx->arena.host_scratch());
  MSGPU_CUDA(cudaMemcpyAsync(h, c.plan.flags, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  MSGPU_CUDA(cudaMemcpyAsync(h + 64, c.sc[b].totals, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  MSGPU_CUDA(cudaEventRecord(c.rows_done, st));
  if (c.next_tile < c.plan.tile_hi) {
    const uint32_t per = c.sc[1 - b].max_tiles;
    c.cur_lo           = c.next_tile;
    c.cur_hi           = c.cur_lo + per < c.plan.tile_hi ? c.cur_lo + per : c.plan.tile_hi;
    pairing_tiles(ctx->arena, st, c.plan, c.cur_lo, c.cur_hi, c.sc[1 - b]);
    c.next_tile = c.cur_hi;
    c.cur       = 1 - b;
  } else {
    c.cur = -1;
  }
  MSGPU_CUDA(cudaEventSynchronize(c.rows_done));
  const uint32_t           *h_flags = reinterpret_cast<const uint32_t *>(h);
  const unsigned long long *h_tot   = reinterpret_cast<const unsigned long long *>(h + 64);
  if (h_flags[2]) throw BuildError{MSGPU_ERR_LIMIT, "a read shares 4096 or more unitigs with other reads"};
  if (h_flags[3]) throw BuildError{MSGPU_ERR_INTERNAL, "pairing: tile invariant violated (code " + std::to_string(h_flags[3]) + ")"};
  e         = c.out[b];
  e.n_em    = h_tot[0];
  e.n_edges = h_tot[1];
  e.n_pairs = c.plan.n_pairs;
  *index    = c.index++;
  *last     = c.cur < 0;
  c.em_base += e.n_em;
  c.edge_base += e.n_edges;
  ctx->edges.n_em    = c.em_base; // rows delivered so far (msgpu_result_counts)
  ctx->edges.n_edges = c.edge_base;
  c.last      = e;
  c.have_last = true;
  return true;
}

// device part of a build: PAF text in HBM -> all tables in HBM.  pair = false: ingest only (one shard of a
// multi-GPU build); the edge tables stay empty.
void build_device(msgpu_ctx *ctx, const char *d_paf, size_t n_bytes, bool pair = true, bool drop_last_line = true) {
  cudaStream_t st  = ctx->stream;
  ctx->have_result = false;
  ctx->host_valid  = false;
  ctx->n_parts     = 1;
  ctx->orders      = OrderTables{};
  ctx->filters     = FilterTables{};
  ctx->filter_view = FilterView{};
  ctx->arena.reset();
  ctx->profiler.clear();
  MSGPU_CUDA(cudaEventRecord(ctx->ev[1], st));
  run_ingest(ctx->arena, st, d_paf, n_bytes, &ctx->n_lines, ctx->reads, ctx->unitigs, ctx->matches, drop_last_line);
  MSGPU_CUDA(cudaEventRecord(ctx->ev[2], st));
  if (pair) {
    pair_or_plan(ctx, 0, 1);
  } else {
    ctx->chunks.active  = false;
    ctx->edges          = EdgeTables{};
    ctx->edges.em_begin = ctx->arena.alloc<uint64_t>(1);
    MSGPU_CUDA(cudaMemsetAsync(ctx->edges.em_begin, 0, sizeof(uint64_t), st));
    MSGPU_CUDA(cudaEventRecord(ctx->ev[3], st));
  }
  MSGPU_CUDA(cudaEventRecord(ctx->ev[4], st));
  MSGPU_CUDA(cudaStreamSynchronize(st));
  ctx->timings.ingest_ms = elapsed(ctx->ev[1], ctx->ev[2]);
  ctx->timings.pair_ms   = elapsed(ctx->ev[2], ctx->ev[3]);
  ctx->timings.merge_ms  = elapsed(ctx->ev[3], ctx->ev[4]);
  ctx->timings.total_ms  = elapsed(ctx->ev[1], ctx->ev[4]);
  ctx->timings.h2d_ms = ctx->timings.d2h_ms = 0.f;
  ctx->have_result = true;
}

void build_host(msgpu_ctx *ctx, const char *host_paf, size_t n_bytes, bool pair = true, bool drop_last_line = true) {
  cudaStream_t st = ctx->stream;
  ctx->input_arena.reset();
  char *d_paf = ctx->input_arena.alloc<char>(n_bytes + 16);
  MSGPU_CUDA(cudaEventRecord(ctx->ev[0], st));
  if (n_bytes) MSGPU_CUDA(cudaMemcpyAsync(d_paf, host_paf, n_bytes, cudaMemcpyHostToDevice, st));
  build_device(ctx, d_paf, n_bytes, pair, drop_last_line);
  ctx->timings.h2d_ms = elapsed(ctx->ev[0], ctx->ev[1]);
  ctx->timings.total_ms += ctx->timings.h2d_ms;
}

template <class T> size_t slab_reserve(size_t &off, size_t n) {
  size_t at = (off + 63) & ~size_t(63);
  off       = at + n * sizeof(T);
  return at;
}

} // namespace

extern "C" {

int msgpu_abi_version(void) { return MSGPU_ABI_VERSION; }

int msgpu_create(int cuda_device, msgpu_ctx **out_ctx) {
  if (!out_ctx) return MSGPU_ERR_INVALID_ARG;
  *out_ctx = nullptr;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || cuda_device < 0 || cuda_device >= n_dev) {
    // No CPU fallback exists on purpose: without a CUDA device this library cannot do anything.
    return MSGPU_ERR_CUDA;
  }
  msgpu_ctx *ctx = new (std::nothrow) msgpu_ctx;
  if (!ctx) return MSGPU_ERR_NOMEM;
  ctx->device = cuda_device;
  int rc      = guarded(ctx, [&] {
    MSGPU_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    for (auto &e : ctx->ev) MSGPU_CUDA(cudaEventCreate(&e));
  });
  if (rc != MSGPU_OK) {
    delete ctx;
    return rc;
  }
  *out_ctx = ctx;
  return MSGPU_OK;
}

void msgpu_destroy(msgpu_ctx *ctx) {
  if (!ctx) return;
  DeviceGuard g(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  seq_destroy(ctx->seq);
  if (ctx->chunks.rows_done) cudaEventDestroy(ctx->chunks.rows_done);
  for (auto &e : ctx->copy_done)
    if (e) cudaEventDestroy(e);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  for (auto &e : ctx->ev)
    if (e) cudaEventDestroy(e);
  ctx->arena.release();
  ctx->input_arena.release();
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

const char *msgpu_last_error(const msgpu_ctx *ctx) { return ctx ? ctx->error.c_str() : "null context"; }

int msgpu_set_stream(msgpu_ctx *ctx, void *cuda_stream) {
  if (!ctx) return MSGPU_ERR_INVALID_ARG;
  ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
  return MSGPU_OK;
}

int msgpu_build_from_paf_device(msgpu_ctx *ctx, const char *dev_paf, size_t n_bytes) {
  return guarded(ctx, [&] {
    if (!dev_paf && n_bytes) throw BuildError{MSGPU_ERR_INVALID_ARG, "null PAF buffer"};
    build_device(ctx, dev_paf, n_bytes);
  });
}

int msgpu_build_from_paf_buffer(msgpu_ctx *ctx, const char *host_paf, size_t n_bytes) {
  return guarded(ctx, [&] {
    if (!host_paf && n_bytes) throw BuildError{MSGPU_ERR_INVALID_ARG, "null PAF buffer"};
    build_host(ctx, host_paf, n_bytes);
  });
}

int msgpu_build_from_paf(msgpu_ctx *ctx, const char *paf_path) {
  return guarded(ctx, [&] {
    if (!paf_path) throw BuildError{MSGPU_ERR_INVALID_ARG, "null path"};
    int fd = ::open(paf_path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) throw BuildError{MSGPU_ERR_IO, "Can't open blast file."}; // BlastFileAccessor.cpp:44
    struct stat sb {};
    if (fstat(fd, &sb) != 0) {
      ::close(fd);
      throw BuildError{MSGPU_ERR_IO, "Can't open blast file."};
    }
    size_t n   = size_t(sb.st_size);
    void  *map = n ? mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
    ::close(fd);
    if (n && map == MAP_FAILED) throw BuildError{MSGPU_ERR_IO, "Can't open blast file."};
    try {
      build_host(ctx, static_cast<const char *>(map), n);
    } catch (...) {
      if (n) munmap(map, n);
      throw;
    }
    if (n) munmap(map, n);
  });
}

namespace {
__global__ void __launch_bounds__(256)
    block_to_anchor_kernel(uint32_t *__restrict__ em_unitig, uint64_t n, const uint32_t *__restrict__ block_unitig) {
  pdl_sync();
  for (uint64_t k = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += uint64_t(gridDim.x) * blockDim.x)
    em_unitig[k] = block_unitig[em_unitig[k]];
}

// the scrubber's anchor graph (pipeline/scrubber_bfs.py:57-114) from PAF text in HBM
void anchor_graph_device(msgpu_ctx *ctx, const char *d_paf, size_t n_bytes) {
  cudaStream_t st  = ctx->stream;
  ctx->have_result = false;
  ctx->host_valid  = false;
  ctx->n_parts     = 1;
  ctx->orders      = OrderTables{};
  ctx->filters     = FilterTables{};
  ctx->filter_view = FilterView{};
  {
    cudaEvent_t keep      = ctx->chunks.rows_done;
    ctx->chunks           = msgpu_ctx::Chunks{};
    ctx->chunks.rows_done = keep;
  }
  ctx->arena.reset();
  ctx->profiler.clear();
  MSGPU_CUDA(cudaEventRecord(ctx->ev[1], st));
  run_ingest(ctx->arena, st, d_paf, n_bytes, &ctx->n_lines, ctx->reads, ctx->unitigs, ctx->matches, false, 1);
  MSGPU_CUDA(cudaEventRecord(ctx->ev[2], st));
  // groups = blocks of consecutive processed lines naming the same anchor; every pair of a block is an edge
  uint32_t *block = nullptr, *block_unitig = nullptr;
  const uint64_t n_blocks = run_scrubber_blocks(ctx->arena, st, ctx->matches, &block, &block_unitig);
  MatchSoA by_block = ctx->matches;
  by_block.unitig   = block;
  UnitigTable blocks;
  blocks.n = n_blocks;
  build_group_offsets(ctx->arena, st, by_block, blocks);
  run_pairing_fused(ctx->arena, st, by_block, blocks, ctx->reads.n, ctx->edges, ctx->ev[3], 0, 1, nullptr, true);
  if (ctx->edges.n_em)
    MSGPU_LAUNCH("block_to_anchor", st, block_to_anchor_kernel<<<grid_for(ctx->edges.n_em, 256), 256, 0, st>>>(
                                            ctx->edges.em_unitig, ctx->edges.n_em, block_unitig));
  ctx->edges.pair_rows = nullptr; // (the stages after construction do not apply to this graph)
  MSGPU_CUDA(cudaEventRecord(ctx->ev[4], st));
  MSGPU_CUDA(cudaStreamSynchronize(st));
  ctx->timings           = msgpu_timings{};
  ctx->timings.ingest_ms = elapsed(ctx->ev[1], ctx->ev[2]);
  ctx->timings.pair_ms   = elapsed(ctx->ev[2], ctx->ev[3]);
  ctx->timings.merge_ms  = elapsed(ctx->ev[3], ctx->ev[4]);
  ctx->timings.total_ms  = elapsed(ctx->ev[1], ctx->ev[4]);
  ctx->have_result       = true;
}
} // namespace

int msgpu_anchor_graph_from_paf_device(msgpu_ctx *ctx, const char *dev_paf, size_t n_bytes) {
  return guarded(ctx, [&] {
    if (!dev_paf && n_bytes) throw BuildError{MSGPU_ERR_INVALID_ARG, "null PAF buffer"};
    anchor_graph_device(ctx, dev_paf, n_bytes);
  });
}

int msgpu_anchor_graph_from_paf_buffer(msgpu_ctx *ctx, const char *host_paf, size_t n_bytes) {
  return guarded(ctx, [&] {
    if (!host_paf && n_bytes) throw BuildError{MSGPU_ERR_INVALID_ARG, "null PAF buffer"};
    cudaStream_t st = ctx->stream;
    ctx->input_arena.reset();
    char *d_paf = ctx->input_arena.alloc<char>(n_bytes + 16);
    MSGPU_CUDA(cudaEventRecord(ctx->ev[0], st));
    if (n_bytes) MSGPU_CUDA(cudaMemcpyAsync(d_paf, host_paf, n_bytes, cudaMemcpyHostToDevice, st));
    anchor_graph_device(ctx, d_paf, n_bytes);
    ctx->timings.h2d_ms = elapsed(ctx->ev[0], ctx->ev[1]);
    ctx->timings.total_ms += ctx->timings.h2d_ms;
  });
}

namespace {
void unitig_coverage_device(msgpu_ctx *ctx, const char *d_paf, size_t n_bytes, msgpu_coverage *out) {
  cudaStream_t st  = ctx->stream;
  ctx->have_result = false;
  ctx->host_valid  = false;
  ctx->n_parts     = 1;
  ctx->orders      = OrderTables{};
  ctx->filters     = FilterTables{};
  ctx->filter_view = FilterView{};
  ctx->coverage    = CoverageTables{};
  {
    cudaEvent_t keep      = ctx->chunks.rows_done;
    ctx->chunks           = msgpu_ctx::Chunks{};
    ctx->chunks.rows_done = keep;
  }
  ctx->arena.reset();
  ctx->profiler.clear();
  MSGPU_CUDA(cudaEventRecord(ctx->ev[1], st));
  run_ingest(ctx->arena, st, d_paf, n_bytes, &ctx->n_lines, ctx->reads, ctx->unitigs, ctx->matches, false, 2);
  MSGPU_CUDA(cudaEventRecord(ctx->ev[2], st));
  run_unitig_coverage(ctx->arena, st, ctx->matches, ctx->unitigs.n, ctx->coverage);
  ctx->edges          = EdgeTables{}; // no graph: the read / unitig / record tables are what msgpu_result_host returns
  ctx->edges.em_begin = ctx->arena.alloc<uint64_t>(1);
  MSGPU_CUDA(cudaMemsetAsync(ctx->edges.em_begin, 0, sizeof(uint64_t), st));
  MSGPU_CUDA(cudaEventRecord(ctx->ev[3], st));
  const CoverageTables &c = ctx->coverage;
  const uint64_t        B = c.n_blocks;
  char *h = static_cast<char *>(ctx->host_coverage.get(4 * (B + 1) * sizeof(uint32_t) + 64));
  uint32_t *hu = reinterpret_cast<uint32_t *>(h);
  const uint32_t *src[4] = {c.block_unitig, c.block_n_reads, c.block_len, c.block_max};
  for (int k = 0; k < 4; ++k)
    if (B) MSGPU_CUDA(cudaMemcpyAsync(hu + k * (B + 1), src[k], B * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  MSGPU_CUDA(cudaEventRecord(ctx->ev[4], st));
  MSGPU_CUDA(cudaStreamSynchronize(st));
  ctx->timings           = msgpu_timings{};
  ctx->timings.ingest_ms = elapsed(ctx->ev[1], ctx->ev[2]);
  ctx->timings.pair_ms   = elapsed(ctx->ev[2], ctx->ev[3]);
  ctx->timings.total_ms  = elapsed(ctx->ev[1], ctx->ev[4]);
  ctx->have_result       = true;
  out->n_blocks          = B;
  out->block_unitig      = hu;
  out->block_n_reads     = hu + (B + 1);
  out->block_len         = hu + 2 * (B + 1);
  out->block_max_cov     = hu + 3 * (B + 1);
}
} // namespace

int msgpu_unitig_coverage_from_paf_device(msgpu_ctx *ctx, const char *dev_paf, size_t n_bytes, msgpu_coverage *out) {
  if (!out) return MSGPU_ERR_INVALID_ARG;
  return guarded(ctx, [&] {
    if (!dev_paf && n_bytes) throw BuildError{MSGPU_ERR_INVALID_ARG, "null PAF buffer"};
    unitig_coverage_device(ctx, dev_paf, n_bytes, out);
  });
}

int msgpu_unitig_coverage_from_paf_buffer(msgpu_ctx *ctx, const char *host_paf, size_t n_bytes, msgpu_coverage *out) {
  if (!out) return MSGPU_ERR_INVALID_ARG;
  return guarded(ctx, [&] {
    if (!host_paf && n_bytes) throw BuildError{MSGPU_ERR_INVALID_ARG, "null PAF buffer"};
    cudaStream_t st = ctx->stream;
    ctx->input_arena.reset();
    char *d_paf = ctx->input_arena.alloc<char>(n_bytes + 16);
    if (n_bytes) MSGPU_CUDA(cudaMemcpyAsync(d_paf, host_paf, n_bytes, cudaMemcpyHostToDevice, st));
    unitig_coverage_device(ctx, d_paf, n_bytes, out);
  });
}

int msgpu_unitig_coverage_cut(msgpu_ctx *ctx, double cutoff, double q3, msgpu_fragments *out) {
  if (!out) return MSGPU_ERR_INVALID_ARG;
  return guarded(ctx, [&] {
    if (!ctx->have_result || !ctx->coverage.valid)
      throw BuildError{MSGPU_ERR_NO_RESULT, "msgpu_unitig_coverage_cut needs msgpu_unitig_coverage_from_paf_* first"};
    cudaStream_t st = ctx->stream;
    ctx->profiler.clear();
    run_unitig_cut(ctx->arena, st, ctx->coverage, cutoff, q3);
    const CoverageTables &c = ctx->coverage;
    const uint64_t        B = c.n_blocks, F = c.n_frags;
    const size_t o_block = (B + 64) & ~size_t(63), o_start = o_block + (F + 1) * 4, o_end = o_start + (F + 1) * 4;
    char *h = static_cast<char *>(ctx->host_fragments.get(o_end + (F + 1) * 4 + 64));
    if (B) MSGPU_CUDA(cudaMemcpyAsync(h, c.block_outlier, B, cudaMemcpyDeviceToHost, st));
    if (F) {
      MSGPU_CUDA(cudaMemcpyAsync(h + o_block, c.frag_block, F * 4, cudaMemcpyDeviceToHost, st));
      MSGPU_CUDA(cudaMemcpyAsync(h + o_start, c.frag_start, F * 4, cudaMemcpyDeviceToHost, st));
      MSGPU_CUDA(cudaMemcpyAsync(h + o_end, c.frag_end, F * 4, cudaMemcpyDeviceToHost, st));
    }
    MSGPU_CUDA(cudaStreamSynchronize(st));
    out->n_blocks      = B;
    out->n_fragments   = F;
    out->block_outlier = reinterpret_cast<const uint8_t *>(h);
    out->frag_block    = reinterpret_cast<const uint32_t *>(h + o_block);
    out->frag_start    = reinterpret_cast<const int32_t *>(h + o_start);
    out->frag_end      = reinterpret_cast<const int32_t *>(h + o_end);
  });
}

int msgpu_build_from_records(msgpu_ctx *ctx, const msgpu_match_soa *r, uint32_t n_reads, uint32_t n_unitigs) {
  return guarded(ctx, [&] {
    if (!r) throw BuildError{MSGPU_ERR_INVALID_ARG, "null records"};
    cudaStream_t st  = ctx->stream;
    ctx->have_result = false;
    ctx->host_valid  = false;
    ctx->orders      = OrderTables{};
    ctx->filters     = FilterTables{};
    ctx->filter_view = FilterView{};
    ctx->arena.reset();
    ctx->profiler.clear();
    ctx->n_parts = 1;
    ctx->reads   = ReadTable{};
    ctx->unitigs = UnitigTable{};
    ctx->reads.n = n_reads;
    ctx->unitigs.n = n_unitigs;
    ctx->n_lines = 0;
    MatchSoA &m = ctx->matches;
    m = MatchSoA{};
    m.n      = r->n;
    m.unitig = const_cast<uint32_t *>(r->unitig);
    m.read   = const_cast<uint32_t *>(r->read);
    m.ir0    = const_cast<int32_t *>(r->ir0);
    m.ir1    = const_cast<int32_t *>(r->ir1);
    m.nr0    = const_cast<int32_t *>(r->nr0);
    m.nr1    = const_cast<int32_t *>(r->nr1);
    m.score  = const_cast<int32_t *>(r->score);
    m.line   = const_cast<uint32_t *>(r->line);
    m.flags  = const_cast<uint8_t *>(r->flags);
    MSGPU_CUDA(cudaEventRecord(ctx->ev[1], st));
    build_group_offsets(ctx->arena, st, m, ctx->unitigs);
    MSGPU_CUDA(cudaEventRecord(ctx->ev[2], st));
    pair_or_plan(ctx, 0, 1);
    MSGPU_CUDA(cudaEventRecord(ctx->ev[4], st));
    MSGPU_CUDA(cudaStreamSynchronize(st));
    ctx->timings           = msgpu_timings{};
    ctx->timings.ingest_ms = elapsed(ctx->ev[1], ctx->ev[2]);
    ctx->timings.pair_ms   = elapsed(ctx->ev[2], ctx->ev[3]);
    ctx->timings.merge_ms  = elapsed(ctx->ev[3], ctx->ev[4]);
    ctx->timings.total_ms  = elapsed(ctx->ev[1], ctx->ev[4]);
    ctx->have_result       = true;
  });
}

int msgpu_shard_ingest_device(msgpu_ctx *ctx, const char *dev_paf, size_t n_bytes, int is_last_shard) {
  return guarded(ctx, [&] {
    if (!dev_paf && n_bytes) throw BuildError{MSGPU_ERR_INVALID_ARG, "null PAF buffer"};
    build_device(ctx, dev_paf, n_bytes, false, is_last_shard != 0);
  });
}

int msgpu_shard_ingest_buffer(msgpu_ctx *ctx, const char *host_paf, size_t n_bytes, int is_last_shard) {
  return guarded(ctx, [&] {
    if (!host_paf && n_bytes) throw BuildError{MSGPU_ERR_INVALID_ARG, "null PAF buffer"};
    build_host(ctx, host_paf, n_bytes, false, is_last_shard != 0);
  });
}

int msgpu_shard_describe(const msgpu_ctx *ctx, msgpu_shard_header *out) {
  if (!ctx || !out) return MSGPU_ERR_INVALID_ARG;
  if (!ctx->have_result || !ctx->reads.name_off) return MSGPU_ERR_NO_RESULT;
  *out = shard_describe(ctx->n_lines, ctx->reads, ctx->unitigs, ctx->matches);
  return MSGPU_OK;
}

int msgpu_shard_export(msgpu_ctx *ctx, void *dev_blob, uint64_t capacity) {
  return guarded(ctx, [&] {
    if (!ctx->have_result || !ctx->reads.name_off) throw BuildError{MSGPU_ERR_NO_RESULT, "no ingest result to export"};
    msgpu_shard_header h = shard_describe(ctx->n_lines, ctx->reads, ctx->unitigs, ctx->matches);
    if (!dev_blob || capacity < h.blob_bytes) throw BuildError{MSGPU_ERR_INVALID_ARG, "blob buffer too small"};
    shard_export(ctx->stream, h, ctx->reads, ctx->unitigs, ctx->matches, dev_blob);
  });
}

int msgpu_shard_export_peers(msgpu_ctx *ctx, void *const *dev_slots, uint32_t n_slots, uint64_t capacity) {
  return guarded(ctx, [&] {
    if (!ctx->have_result || !ctx->reads.name_off) throw BuildError{MSGPU_ERR_NO_RESULT, "no ingest result to export"};
    if (!dev_slots || !n_slots || n_slots > kMaxPeers || capacity < sizeof(msgpu_shard_header))
      throw BuildError{MSGPU_ERR_INVALID_ARG, "bad peer slots"};
    for (uint32_t d = 0; d < n_slots; ++d)
      if (!dev_slots[d] || (reinterpret_cast<uintptr_t>(dev_slots[d]) & 15u))
        throw BuildError{MSGPU_ERR_INVALID_ARG, "peer slots must be 16-byte aligned device pointers"};
    msgpu_shard_header h = shard_describe(ctx->n_lines, ctx->reads, ctx->unitigs, ctx->matches);
    shard_export_peers(ctx->stream, h, ctx->reads, ctx->unitigs, ctx->matches, dev_slots, n_slots, h.blob_bytes <= capacity);
  });
}

int msgpu_build_from_shards(msgpu_ctx *ctx, uint32_t n_shards, const msgpu_shard_header *headers,
                            const void *const *dev_blobs, uint32_t part, uint32_t n_parts) {
  return guarded(ctx, [&] {
    if (!n_shards || !headers || !dev_blobs || !n_parts || part >= n_parts)
      throw BuildError{MSGPU_ERR_INVALID_ARG, "bad shard arguments"};
    cudaStream_t st  = ctx->stream;
    ctx->have_result = false;
    ctx->host_valid  = false;
    ctx->orders      = OrderTables{};
    ctx->filters     = FilterTables{};
    ctx->filter_view = FilterView{};
    ctx->arena.reset();
    ctx->profiler.clear();
    MSGPU_CUDA(cudaEventRecord(ctx->ev[1], st));
    run_shard_merge(ctx->arena, st, n_shards, headers, dev_blobs, &ctx->n_lines, ctx->reads, ctx->unitigs,
                    ctx->matches);
    MSGPU_CUDA(cudaEventRecord(ctx->ev[2], st));
    ctx->n_parts = n_parts;
    pair_or_plan(ctx, part, n_parts);
    MSGPU_CUDA(cudaEventRecord(ctx->ev[4], st));
    MSGPU_CUDA(cudaStreamSynchronize(st));
    ctx->timings           = msgpu_timings{};
    ctx->timings.ingest_ms = elapsed(ctx->ev[1], ctx->ev[2]);
    ctx->timings.pair_ms   = elapsed(ctx->ev[2], ctx->ev[3]);
    ctx->timings.merge_ms  = elapsed(ctx->ev[3], ctx->ev[4]);
    ctx->timings.total_ms  = elapsed(ctx->ev[1], ctx->ev[4]);
    ctx->have_result       = true;
  });
}

int msgpu_result_counts(const msgpu_ctx *ctx, msgpu_counts *out) {
  if (!ctx || !out) return MSGPU_ERR_INVALID_ARG;
  if (!ctx->have_result) return MSGPU_ERR_NO_RESULT;
  out->n_lines        = ctx->n_lines;
  out->n_reads        = ctx->reads.n;
  out->n_unitigs      = ctx->unitigs.n;
  out->n_matches      = ctx->matches.n;
  out->n_pairs        = ctx->edges.n_pairs;
  out->n_edges        = ctx->edges.n_edges;
  out->n_edge_matches = ctx->edges.n_em;
  return MSGPU_OK;
}

int msgpu_result_device(const msgpu_ctx *ctx, msgpu_tables *t) {
  if (!ctx || !t) return MSGPU_ERR_INVALID_ARG;
  if (!ctx->have_result) return MSGPU_ERR_NO_RESULT;
  std::memset(t, 0, sizeof(*t));
  t->read_nlen       = ctx->reads.nlen;
  t->read_line       = ctx->reads.line;
  t->read_name_off   = ctx->reads.name_off;
  t->read_names      = ctx->reads.names;
  t->unitig_name_off = ctx->unitigs.name_off;
  t->unitig_names    = ctx->unitigs.names;
  t->m_unitig        = ctx->matches.unitig;
  t->m_read          = ctx->matches.read;
  t->m_nr0           = ctx->matches.nr0;
  t->m_nr1           = ctx->matches.nr1;
  t->m_ir0           = ctx->matches.ir0;
  t->m_ir1           = ctx->matches.ir1;
  t->m_rratio        = ctx->matches.rratio;
  t->m_score         = ctx->matches.score;
  t->m_line          = ctx->matches.line;
  t->m_flags         = ctx->matches.flags;
  t->edge_first      = ctx->edges.first;
  t->edge_second     = ctx->edges.second;
  t->edge_em_begin   = ctx->edges.em_begin;
  t->em_unitig       = ctx->edges.em_unitig;
  t->em_ov0          = ctx->edges.em_ov0;
  t->em_ov1          = ctx->edges.em_ov1;
  t->em_score        = ctx->edges.em_score;
  t->em_line         = ctx->edges.em_line;
  t->em_flags        = ctx->edges.em_flags;
  return MSGPU_OK;
}

int msgpu_result_host(msgpu_ctx *ctx, msgpu_tables *out) {
  if (!ctx || !out) return MSGPU_ERR_INVALID_ARG;
  if (!ctx->have_result) return MSGPU_ERR_NO_RESULT;
  if (ctx->host_valid) {
    *out = ctx->host_tables;
    return MSGPU_OK;
  }
  return guarded(ctx, [&] {
    cudaStream_t st = ctx->stream;
    const bool   chunked = ctx->chunks.active;
    if (chunked && ctx->chunks.taken)
      throw BuildError{MSGPU_ERR_NO_RESULT, "chunks of this build were already taken with msgpu_chunk_next"};
    msgpu_tables d;
    msgpu_result_device(ctx, &d);
    const uint64_t R = ctx->reads.n, U = ctx->unitigs.n, M = ctx->matches.n;
    // chunked: the row counts are not known yet; lay the slab out for what the device's tiles can produce at most
    const uint64_t cap_rows = chunked ? pairing_rows_bound(ctx->chunks.plan, ctx->chunks.plan.tile_hi - ctx->chunks.plan.tile_lo) : 0;
    const uint64_t E = chunked ? cap_rows : ctx->edges.n_edges, EM = chunked ? cap_rows : ctx->edges.n_em;
    const bool have_names = ctx->reads.name_off != nullptr;
    // lay every table out in one pinned slab
    size_t off = 0;
    struct Copy {
      const void *src;
      size_t      at, bytes;
      const void **dst;
    };
    std::vector<Copy> copies;
    msgpu_tables     &h = ctx->host_tables;
    std::memset(&h, 0, sizeof(h));
    auto add = [&](auto const *src, size_t n, auto const **dst) {
      using T   = std::remove_cv_t<std::remove_pointer_t<decltype(src)>>;
      size_t at = slab_reserve<T>(off, n);
      copies.push_back(Copy{src, at, n * sizeof(T), reinterpret_cast<const void **>(dst)});
    };
    if (have_names) {
      add(d.read_nlen, R, &h.read_nlen);
      add(d.read_line, R, &h.read_line);
      add(d.read_name_off, R + 1, &h.read_name_off);
      add(d.read_names, ctx->reads.n_bytes, &h.read_names);
      add(d.unitig_name_off, U + 1, &h.unitig_name_off);
      add(d.unitig_names, ctx->unitigs.n_bytes, &h.unitig_names);
    }
    add(d.m_unitig, M, &h.m_unitig);
    add(d.m_read, M, &h.m_read);
    add(d.m_nr0, M, &h.m_nr0);
    add(d.m_nr1, M, &h.m_nr1);
    add(d.m_ir0, M, &h.m_ir0);
    add(d.m_ir1, M, &h.m_ir1);
    if (d.m_rratio) add(d.m_rratio, M, &h.m_rratio);
    add(d.m_score, M, &h.m_score);
    add(d.m_line, M, &h.m_line);
    add(d.m_flags, M, &h.m_flags);
    const size_t n_small = copies.size(); // what follows are the edge / edge-match columns
    add(d.edge_first, E, &h.edge_first);
    add(d.edge_second, E, &h.edge_second);
    add(d.edge_em_begin, E + 1, &h.edge_em_begin);
    add(d.em_unitig, EM, &h.em_unitig);
    add(d.em_ov0, EM, &h.em_ov0);
    add(d.em_ov1, EM, &h.em_ov1);
    add(d.em_score, EM, &h.em_score);
    add(d.em_line, EM, &h.em_line);
    add(d.em_flags, EM, &h.em_flags);
    char *slab = static_cast<char *>(ctx->host_result.get(off + 64));
    for (auto &c : copies) *c.dst = slab + c.at;
    if (!chunked) {
      MSGPU_CUDA(cudaEventRecord(ctx->ev[4], st));
      for (auto &c : copies)
        if (c.bytes) MSGPU_CUDA(cudaMemcpyAsync(slab + c.at, c.src, c.bytes, cudaMemcpyDeviceToHost, st));
      MSGPU_CUDA(cudaEventRecord(ctx->ev[5], st));
      MSGPU_CUDA(cudaStreamSynchronize(st));
      ctx->timings.d2h_ms = elapsed(ctx->ev[4], ctx->ev[5]);
    } else {
      // Chunks are downloaded on a second stream while the next ones are computed: the rows half of chunk c + 2
      // waits (on the device) for the download of chunk c, which used the same buffer set.
      if (!ctx->copy_stream) MSGPU_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
      for (auto &e : ctx->copy_done)
        if (!e) MSGPU_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      cudaStream_t cs = ctx->copy_stream;
      MSGPU_CUDA(cudaEventRecord(ctx->ev[4], st)); // everything queued so far (the plan, the read / match tables)
      MSGPU_CUDA(cudaStreamWaitEvent(cs, ctx->ev[4], 0));
      for (size_t i = 0; i < n_small; ++i)
        if (copies[i].bytes)
          MSGPU_CUDA(cudaMemcpyAsync(slab + copies[i].at, copies[i].src, copies[i].bytes, cudaMemcpyDeviceToHost, cs));
      auto at = [](const void *hp) { return static_cast<char *>(const_cast<void *>(hp)); };
      EdgeTables e;
      uint64_t   index = 0;
      bool       last  = false;
      int        b     = ctx->chunks.cur; // buffer set of the chunk about to be delivered
      uint64_t   eb = 0, mb = 0;
      ctx->copy_pending[0] = ctx->copy_pending[1] = false;
      while (next_chunk(ctx, e, &index, &last)) {
        auto put = [&](char *dst, const void *src, size_t bytes) {
          if (bytes) MSGPU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, cs));
        };
        put(at(h.edge_first) + eb * 4, e.first, e.n_edges * 4);
        put(at(h.edge_second) + eb * 4, e.second, e.n_edges * 4);
        put(at(h.edge_em_begin) + eb * 8, e.em_begin, (e.n_edges + 1) * 8);
        put(at(h.em_unitig) + mb * 4, e.em_unitig, e.n_em * 4);
        put(at(h.em_ov0) + mb * 4, e.em_ov0, e.n_em * 4);
        put(at(h.em_ov1) + mb * 4, e.em_ov1, e.n_em * 4);
        put(at(h.em_score) + mb * 8, e.em_score, e.n_em * 8);
        put(at(h.em_line) + mb * 4, e.em_line, e.n_em * 4);
        put(at(h.em_flags) + mb, e.em_flags, e.n_em);
        MSGPU_CUDA(cudaEventRecord(ctx->copy_done[b], cs));
        ctx->copy_pending[b] = true;
        eb += e.n_edges;
        mb += e.n_em;
        b = ctx->chunks.cur;
      }
      MSGPU_CUDA(cudaStreamSynchronize(cs));
      MSGPU_CUDA(cudaStreamSynchronize(st));
      ctx->copy_pending[0] = ctx->copy_pending[1] = false;
      ctx->chunks.taken = false; // (the chunks went to the host tables, not to the caller)
      ctx->timings.d2h_ms = 0.f;
    }
    ctx->timings.total_ms += ctx->timings.d2h_ms;
    ctx->host_valid = true;
    *out            = h;
  });
}

int msgpu_set_chunk_tiles(msgpu_ctx *ctx, uint32_t tiles_per_chunk) {
  if (!ctx) return MSGPU_ERR_INVALID_ARG;
  ctx->chunk_tiles = tiles_per_chunk;
  return MSGPU_OK;
}

int msgpu_chunk_next(msgpu_ctx *ctx, msgpu_chunk *out) {
  if (!ctx || !out) return MSGPU_ERR_INVALID_ARG;
  return guarded(ctx, [&] {
    if (!ctx->have_result || !ctx->chunks.active) throw BuildError{MSGPU_ERR_NO_RESULT, "no chunked build in progress"};
    EdgeTables e;
    uint64_t   index = 0;
    bool       last  = false;
    const uint64_t em_base = ctx->chunks.em_base, edge_base = ctx->chunks.edge_base;
    if (!next_chunk(ctx, e, &index, &last)) throw BuildError{MSGPU_ERR_NO_RESULT, "every chunk has been delivered"};
    ctx->chunks.taken = true;
    std::memset(out, 0, sizeof(*out));
    out->index          = index;
    out->n_edges        = e.n_edges;
    out->n_edge_matches = e.n_em;
    out->edge_base      = edge_base;
    out->em_base        = em_base;
    out->last           = last ? 1 : 0;
    out->edge_first     = e.first;
    out->edge_second    = e.second;
    out->edge_em_begin  = e.em_begin;
    out->em_unitig      = e.em_unitig;
    out->em_ov0         = e.em_ov0;
    out->em_ov1         = e.em_ov1;
    out->em_score       = e.em_score;
    out->em_line        = e.em_line;
    out->em_flags       = e.em_flags;
  });
}

int msgpu_chunk_digest(msgpu_ctx *ctx, msgpu_digest *out) {
  if (!ctx || !out) return MSGPU_ERR_INVALID_ARG;
  return guarded(ctx, [&] {
    if (!ctx->chunks.active || !ctx->chunks.have_last) throw BuildError{MSGPU_ERR_NO_RESULT, "no chunk delivered yet"};
    run_digest(ctx->arena, ctx->stream, ctx->chunks.last, out);
  });
}

int msgpu_result_digest(msgpu_ctx *ctx, msgpu_digest *out) {
  if (!ctx || !out) return MSGPU_ERR_INVALID_ARG;
  return guarded(ctx, [&] {
    if (!ctx->have_result) throw BuildError{MSGPU_ERR_NO_RESULT, "no construction to digest"};
    if (ctx->chunks.active) throw BuildError{MSGPU_ERR_NO_RESULT, "chunked build: use msgpu_chunk_digest per chunk"};
    run_digest(ctx->arena, ctx->stream, ctx->edges, out);
  });
}

int msgpu_chain_overlaps(msgpu_ctx *ctx, uint64_t wiggle_room) {
  return guarded(ctx, [&] {
    if (!ctx->have_result) throw BuildError{MSGPU_ERR_NO_RESULT, "no construction to chain"};
    if (ctx->chunks.active) throw BuildError{MSGPU_ERR_NO_RESULT, "the chaining stage needs the resident edge tables (chunking is on)"};
    ctx->orders             = OrderTables{};
    ctx->orders_host_valid  = false;
    ctx->filters            = FilterTables{};
    ctx->filter_view = FilterView{};
    ctx->filters_host_valid = false;
    ctx->profiler.clear();
    run_chain_overlaps(ctx->arena, ctx->stream, ctx->reads, ctx->unitigs, ctx->matches, ctx->edges, wiggle_room,
                       ctx->orders, &ctx->chain_ms);
  });
}

int msgpu_orders_device(const msgpu_ctx *ctx, msgpu_orders *o) {
  if (!ctx || !o) return MSGPU_ERR_INVALID_ARG;
  if (!ctx->have_result || !ctx->orders.valid) return MSGPU_ERR_NO_RESULT;
  o->n_edges       = ctx->edges.n_edges;
  o->n_slots       = ctx->edges.n_em;
  o->edge_shadow   = ctx->orders.edge_shadow;
  o->edge_n_orders = ctx->orders.edge_n_orders;
  o->ord_left      = ctx->orders.left;
  o->ord_right     = ctx->orders.right;
  o->ord_score     = ctx->orders.score;
  o->ord_n_ids     = ctx->orders.n_ids;
  o->ord_flags     = ctx->orders.flags;
  o->ord_ids       = ctx->orders.ids;
  return MSGPU_OK;
}

int msgpu_orders_host(msgpu_ctx *ctx, msgpu_orders *out) {
  if (!ctx || !out) return MSGPU_ERR_INVALID_ARG;
  if (!ctx->have_result || !ctx->orders.valid) return MSGPU_ERR_NO_RESULT;
  if (ctx->orders_host_valid) {
    *out = ctx->host_orders_view;
    return MSGPU_OK;
  }
  return guarded(ctx, [&] {
    msgpu_orders d;
    msgpu_orders_device(ctx, &d);
    const uint64_t E = d.n_edges, S = d.n_slots;
    size_t         off = 0;
    const size_t a_sh = slab_reserve<uint8_t>(off, E), a_no = slab_reserve<uint32_t>(off, E),
                 a_l = slab_reserve<double>(off, S), a_r = slab_reserve<double>(off, S),
                 a_sc = slab_reserve<uint64_t>(off, S), a_ni = slab_reserve<uint32_t>(off, S),
                 a_fl = slab_reserve<uint8_t>(off, S), a_id = slab_reserve<uint32_t>(off, S);
    char        *slab = static_cast<char *>(ctx->host_orders.get(off + 64));
    cudaStream_t st   = ctx->stream;
    auto copy = [&](size_t at, const void *src, size_t bytes) {
      if (bytes) MSGPU_CUDA(cudaMemcpyAsync(slab + at, src, bytes, cudaMemcpyDeviceToHost, st));
    };
    copy(a_sh, d.edge_shadow, E);
    copy(a_no, d.edge_n_orders, E * 4);
    copy(a_l, d.ord_left, S * 8);
    copy(a_r, d.ord_right, S * 8);
    copy(a_sc, d.ord_score, S * 8);
    copy(a_ni, d.ord_n_ids, S * 4);
    copy(a_fl, d.ord_flags, S);
    copy(a_id, d.ord_ids, S * 4);
    MSGPU_CUDA(cudaStreamSynchronize(st));
    msgpu_orders &h = ctx->host_orders_view;
    h               = d;
    h.edge_shadow   = reinterpret_cast<const uint8_t *>(slab + a_sh);
    h.edge_n_orders = reinterpret_cast<const uint32_t *>(slab + a_no);
    h.ord_left      = reinterpret_cast<const double *>(slab + a_l);
    h.ord_right     = reinterpret_cast<const double *>(slab + a_r);
    h.ord_score     = reinterpret_cast<const uint64_t *>(slab + a_sc);
    h.ord_n_ids     = reinterpret_cast<const uint32_t *>(slab + a_ni);
    h.ord_flags     = reinterpret_cast<const uint8_t *>(slab + a_fl);
    h.ord_ids       = reinterpret_cast<const uint32_t *>(slab + a_id);
    ctx->orders_host_valid = true;
    *out                   = h;
  });
}

int msgpu_chain_ms(const msgpu_ctx *ctx, float *out) {
  if (!ctx || !out) return MSGPU_ERR_INVALID_ARG;
  if (!ctx->orders.valid) return MSGPU_ERR_NO_RESULT;
  *out = ctx->chain_ms;
  return MSGPU_OK;
}

int msgpu_edge_filters(msgpu_ctx *ctx, uint64_t wiggle_room) {
  return guarded(ctx, [&] {
    if (!ctx->have_result || !ctx->orders.valid)
      throw BuildError{MSGPU_ERR_NO_RESULT, "the edge filters need msgpu_chain_overlaps first"};
    // findContractionEdges looks at the edges AROUND the start vertex of an order and at the third edge of every
    // triangle (main.cpp:432-443): on one slice of a multi-GPU build those rows live on other ranks, hasEdge would
    // answer "no" and the verdict would flip.  Refuse instead of answering wrongly.
    if (ctx->n_parts > 1)
      throw BuildError{MSGPU_ERR_LIMIT,
                       "msgpu_edge_filters needs the whole edge table on one device; this context holds one slice of " +
                           std::to_string(ctx->n_parts) + " (multi-GPU build: exchange the views and call msgpu_edge_filters_sharded)"};
    ctx->filters            = FilterTables{};
    ctx->filters_host_valid = false;
    ctx->profiler.clear();
    run_edge_filters(ctx->arena, ctx->stream, ctx->reads.n, ctx->edges, ctx->orders, wiggle_room, ctx->filters,
                     &ctx->filter_ms);
  });
}

int msgpu_filter_view_describe(msgpu_ctx *ctx, msgpu_filter_view_header *out) {
  if (!out) return MSGPU_ERR_INVALID_ARG;
  return guarded(ctx, [&] {
    if (!ctx->have_result || !ctx->orders.valid)
      throw BuildError{MSGPU_ERR_NO_RESULT, "the filter view needs msgpu_chain_overlaps first"};
    filter_view_describe(ctx->arena, ctx->stream, ctx->edges, ctx->orders, ctx->filter_view);
    *out = msgpu_filter_view_header{ctx->filter_view.n_edges, ctx->filter_view.n_orders, ctx->filter_view.bytes, 0};
  });
}

int msgpu_filter_view_export(msgpu_ctx *ctx, void *dev_dst, uint64_t capacity) {
  return guarded(ctx, [&] {
    if (!ctx->have_result || !ctx->orders.valid || !ctx->filter_view.valid)
      throw BuildError{MSGPU_ERR_NO_RESULT, "msgpu_filter_view_describe first"};
    if (!dev_dst || capacity < ctx->filter_view.bytes || (reinterpret_cast<uintptr_t>(dev_dst) & 15))
      throw BuildError{MSGPU_ERR_INVALID_ARG, "filter view: destination missing, misaligned or smaller than blob_bytes"};
    filter_view_export(ctx->stream, ctx->edges, ctx->orders, ctx->filter_view, dev_dst);
  });
}

int msgpu_edge_filters_sharded(msgpu_ctx *ctx, uint32_t n_parts, uint32_t part, const msgpu_filter_view_header *headers,
                               const void *const *dev_blobs, uint64_t wiggle_room) {
  if (!headers || !dev_blobs || n_parts == 0) return MSGPU_ERR_INVALID_ARG;
  return guarded(ctx, [&] {
    if (!ctx->have_result || !ctx->orders.valid)
      throw BuildError{MSGPU_ERR_NO_RESULT, "the edge filters need msgpu_chain_overlaps first"};
    ctx->filters            = FilterTables{};
    ctx->filters_host_valid = false;
    ctx->profiler.clear();
    run_edge_filters_sharded(ctx->arena, ctx->stream, ctx->reads.n, ctx->edges, ctx->orders, wiggle_room, n_parts, part,
                             headers, dev_blobs, ctx->filters, &ctx->filter_ms);
  });
}

int msgpu_edge_filters_device(const msgpu_ctx *ctx, msgpu_edge_filter_tables *o) {
  if (!ctx || !o) return MSGPU_ERR_INVALID_ARG;
  if (!ctx->have_result || !ctx->filters.valid) return MSGPU_ERR_NO_RESULT;
  o->n_edges           = ctx->filters.n_edges;
  o->contraction_order = ctx->filters.contraction;
  o->deletable         = ctx->filters.deletable;
  o->n_kept_orders     = ctx->filters.n_kept;
  o->weight            = ctx->filters.weight;
  o->consensus         = ctx->filters.consensus;
  return MSGPU_OK;
}

int msgpu_edge_filters_host(msgpu_ctx *ctx, msgpu_edge_filter_tables *out) {
  if (!ctx || !out) return MSGPU_ERR_INVALID_ARG;
  if (!ctx->have_result || !ctx->filters.valid) return MSGPU_ERR_NO_RESULT;
  if (ctx->filters_host_valid) {
    *out = ctx->host_filters_view;
    return MSGPU_OK;
  }
  return guarded(ctx, [&] {
    msgpu_edge_filter_tables d;
    msgpu_edge_filters_device(ctx, &d);
    const uint64_t E   = d.n_edges;
    size_t         off = 0;
    const size_t a_c = slab_reserve<int32_t>(off, E), a_d = slab_reserve<uint8_t>(off, E),
                 a_k = slab_reserve<uint32_t>(off, E), a_w = slab_reserve<uint64_t>(off, E),
                 a_s = slab_reserve<uint8_t>(off, E);
    char        *slab = static_cast<char *>(ctx->host_filters.get(off + 64));
    cudaStream_t st   = ctx->stream;
    auto copy = [&](size_t at, const void *src, size_t bytes) {
      if (bytes) MSGPU_CUDA(cudaMemcpyAsync(slab + at, src, bytes, cudaMemcpyDeviceToHost, st));
    };
    copy(a_c, d.contraction_order, E * 4);
    copy(a_d, d.deletable, E);
    copy(a_k, d.n_kept_orders, E * 4);
    copy(a_w, d.weight, E * 8);
    copy(a_s, d.consensus, E);
    MSGPU_CUDA(cudaStreamSynchronize(st));
    msgpu_edge_filter_tables &h = ctx->host_filters_view;
    h                   = d;
    h.contraction_order = reinterpret_cast<const int32_t *>(slab + a_c);
    h.deletable         = reinterpret_cast<const uint8_t *>(slab + a_d);
    h.n_kept_orders     = reinterpret_cast<const uint32_t *>(slab + a_k);
    h.weight            = reinterpret_cast<const uint64_t *>(slab + a_w);
    h.consensus         = reinterpret_cast<const uint8_t *>(slab + a_s);
    ctx->filters_host_valid = true;
    *out                    = h;
  });
}

int msgpu_edge_filters_ms(const msgpu_ctx *ctx, float *out) {
  if (!ctx || !out) return MSGPU_ERR_INVALID_ARG;
  if (!ctx->filters.valid) return MSGPU_ERR_NO_RESULT;
  *out = ctx->filter_ms;
  return MSGPU_OK;
}

int msgpu_profile_enable(msgpu_ctx *ctx, int enable) {
  if (!ctx) return MSGPU_ERR_INVALID_ARG;
  ctx->profiler.enabled = enable != 0;
  return MSGPU_OK;
}

uint64_t msgpu_launch_count(const msgpu_ctx *ctx) { return ctx ? ctx->profiler.launches : 0; }

int msgpu_profile_read(msgpu_ctx *ctx, msgpu_kernel_stat *out, size_t cap, size_t *n_out) {
  if (!ctx || !n_out || (!out && cap)) return MSGPU_ERR_INVALID_ARG;
  return guarded(ctx, [&] {
    MSGPU_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<msgpu_kernel_stat> stats;
    for (auto const &r : ctx->profiler.recs) {
      float ms = 0.f;
      MSGPU_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
      auto it = std::find_if(stats.begin(), stats.end(),
                             [&](msgpu_kernel_stat const &s) { return std::strncmp(s.name, r.name, 47) == 0; });
      if (it == stats.end()) {
        msgpu_kernel_stat s{};
        std::strncpy(s.name, r.name, 47);
        stats.push_back(s);
        it = stats.end() - 1;
      }
      it->launches += 1;
      it->total_ms += ms;
    }
    *n_out = stats.size();
    for (size_t i = 0; i < stats.size() && i < cap; ++i) out[i] = stats[i];
  });
}

int msgpu_result_timings(const msgpu_ctx *ctx, msgpu_timings *out) {
  if (!ctx || !out) return MSGPU_ERR_INVALID_ARG;
  if (!ctx->have_result) return MSGPU_ERR_NO_RESULT;
  *out = ctx->timings;
  return MSGPU_OK;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// sequence entry points: thin forwarding to seq.cu (the store lives in the context)
// ---------------------------------------------------------------------------------------------------------------------
namespace {

// names of the last build, in id order, so the sequence index continues the construction's numbering
// (the reference hands the same Registry objects to SequenceAccessor, src/main.cpp:161-162)
void prime_names(msgpu_ctx *ctx, std::vector<std::string> &reads, std::vector<std::string> &unitigs) {
  if (!ctx->have_result || !ctx->reads.name_off) return;
  msgpu_tables h;
  int          rc = msgpu_result_host(ctx, &h);
  if (rc != MSGPU_OK) throw BuildError{rc, ctx->error};
  for (uint64_t i = 0; i < ctx->reads.n; ++i)
    reads.emplace_back(h.read_names + h.read_name_off[i], h.read_names + h.read_name_off[i + 1]);
  for (uint64_t i = 0; i < ctx->unitigs.n; ++i)
    unitigs.emplace_back(h.unitig_names + h.unitig_name_off[i], h.unitig_names + h.unitig_name_off[i + 1]);
}

} // namespace

extern "C" {

int msgpu_seq_load_buffers(msgpu_ctx *ctx, const char *unitigs_fasta, size_t n_u, const char *nanopore, size_t n_n,
                           int nanopore_is_fastq) {
  return guarded(ctx, [&] {
    std::vector<std::string> rn, un;
    prime_names(ctx, rn, un);
    seq_destroy(ctx->seq);
    ctx->seq = nullptr;
    ctx->seq = seq_create(ctx->stream, unitigs_fasta, n_u, nanopore, n_n, nanopore_is_fastq != 0, rn, un);
  });
}

int msgpu_seq_load(msgpu_ctx *ctx, const char *unitigs_path, const char *nanopore_path) {
  return guarded(ctx, [&] {
    if (!unitigs_path || !nanopore_path) throw BuildError{MSGPU_ERR_INVALID_ARG, "null path"};
    MappedFile fu(unitigs_path), fn(nanopore_path);
    if (!fu.ok || !fn.ok) throw BuildError{MSGPU_ERR_IO, "Can't open sequence file(s)."}; // SequenceAccessor.cpp:110
    std::vector<std::string> rn, un;
    prime_names(ctx, rn, un);
    seq_destroy(ctx->seq);
    ctx->seq = nullptr;
    ctx->seq = seq_create(ctx->stream, fu.data, fu.size, fn.data, fn.size, path_is_fastq(nanopore_path), rn, un);
  });
}

int msgpu_seq_id(const msgpu_ctx *ctx, uint32_t which, const char *name, uint32_t *out_id) {
  if (!ctx || !name || !out_id || which > 1) return MSGPU_ERR_INVALID_ARG;
  if (!ctx->seq) return MSGPU_ERR_NO_RESULT;
  return seq_lookup(ctx->seq, which, name, out_id) ? MSGPU_OK : MSGPU_ERR_OUT_OF_RANGE;
}

int msgpu_seq_batch_bound(const msgpu_ctx *ctx, const msgpu_seq_req *reqs, size_t n, uint64_t *out_bytes) {
  if (!ctx || (!reqs && n) || !out_bytes) return MSGPU_ERR_INVALID_ARG;
  if (!ctx->seq) return MSGPU_ERR_NO_RESULT;
  *out_bytes = seq_batch_bound(ctx->seq, reqs, n);
  return MSGPU_OK;
}

int msgpu_seq_fetch_batch(msgpu_ctx *ctx, const msgpu_seq_req *reqs, size_t n, char *out, uint64_t cap,
                          uint64_t *out_off, uint8_t *status) {
  if (!ctx || (!reqs && n) || !out_off || !status || (!out && cap)) return MSGPU_ERR_INVALID_ARG;
  if (!ctx->seq) return MSGPU_ERR_NO_RESULT;
  // thread-safe: the store is immutable after load and every call uses its own stream + scratch
  try {
    DeviceGuard g(ctx->device);
    seq_fetch_host(ctx->seq, reqs, n, out, cap, out_off, status);
    return MSGPU_OK;
  } catch (BuildError const &e) {
    return e.status;
  } catch (std::exception const &) {
    return MSGPU_ERR_CUDA;
  }
}

int msgpu_seq_fetch_batch_device(msgpu_ctx *ctx, const msgpu_seq_req *dev_reqs, size_t n, char *dev_out, uint64_t cap,
                                 uint64_t *dev_out_off, uint8_t *dev_status) {
  return guarded(ctx, [&] {
    if (!ctx->seq) throw BuildError{MSGPU_ERR_NO_RESULT, "no sequence store loaded"};
    ctx->profiler.clear(); // per-kernel times of THIS batch (msgpu_profile_read)
    seq_fetch_device(ctx->seq, ctx->stream, dev_reqs, n, dev_out, cap, dev_out_off, dev_status);
  });
}

int msgpu_slice_revcomp_device(msgpu_ctx *ctx, const char *dev_text, const uint64_t *dev_text_off,
                               const int32_t *dev_left, const int32_t *dev_right, const uint8_t *dev_revcomp, size_t n,
                               char *dev_out, uint64_t cap, uint64_t *dev_out_off, uint8_t *dev_status) {
  return guarded(ctx, [&] {
    slice_revcomp_device(ctx->stream, dev_text, dev_text_off, dev_left, dev_right, dev_revcomp, n, dev_out, cap,
                         dev_out_off, dev_status);
  });
}

} // extern "C"
