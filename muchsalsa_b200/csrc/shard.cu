// shard.cu — multi-GPU construction: export of one rank's ingest result as a flat blob, and the merge of all ranks'
// blobs into the global id space.
//
// The reference has no distributed mode; what has to be reproduced is its single-threaded result on the WHOLE PAF:
//   read / unitig id = rank of the first accepted line naming it            (Registry.cpp:36-45)
//   a read's vertex carries the line index and length of that first line    (Graph.cpp:141-149)
//   per (read, unitig) the accepted line with the smallest index is kept    (MatchMap.cpp:64-73)
// Each rank ingests one contiguous chunk of the PAF (split at line boundaries) with the single-GPU kernels, which
// gives chunk-local ids and chunk-local line numbers.  Because the chunks are contiguous and in rank order, the
// concatenation of the ranks' read tables (each ordered by local first line) is ordered by GLOBAL first line, so the
// same "first appearance wins, id = rank of first appearances" interning, run over table ENTRIES instead of lines,
// yields the global ids.  Records are then renumbered, deduplicated across chunks and grouped by unitig — all of it
// O(matches), replicated on every rank.  What crosses NVLink is the match table (tens of bytes per match); the
// candidate pairs (two orders of magnitude more) never leave the GPU that owns their rows (pair_fused.cu).
#include <cstring>
#include <vector>

#include "../../include/msgpu.h"
#include "internal.h"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace msgpu {

constexpr uint64_t kShardMagic = 0x3144524853534d47ull; // "GMSSHRD1"

static uint64_t align16(uint64_t x) { return (x + 15) & ~uint64_t(15); }

ShardLayout shard_layout(const msgpu_shard_header &h) {
  ShardLayout L{};
  uint64_t    o = align16(sizeof(msgpu_shard_header));
  auto take = [&](uint64_t bytes) {
    uint64_t at = o;
    o           = align16(o + bytes);
    return at;
  };
  const uint64_t R = h.n_reads, U = h.n_unitigs, M = h.n_matches;
  L.read_nlen       = take(R * 4);
  L.read_line       = take(R * 4);
  L.read_name_off   = take((R + 1) * 8);
  L.read_names      = take(h.read_name_bytes);
  L.unitig_name_off = take((U + 1) * 8);
  L.unitig_names    = take(h.unitig_name_bytes);
  L.m_unitig        = take(M * 4);
  L.m_read          = take(M * 4);
  L.m_ir0           = take(M * 4);
  L.m_ir1           = take(M * 4);
  L.m_nr0           = take(M * 4);
  L.m_nr1           = take(M * 4);
  L.m_score         = take(M * 4);
  L.m_line          = take(M * 4);
  L.m_rratio        = take(0); // not sent: rRatio is recomputed from the ranges, bit for bit (BlastFileReader.cpp:118-119)
  L.m_flags         = take(M);
  L.total           = o;
  return L;
}

msgpu_shard_header shard_describe(uint64_t n_lines, const ReadTable &reads, const UnitigTable &unitigs,
                                  const MatchSoA &m) {
  msgpu_shard_header h{};
  h.magic             = kShardMagic;
  h.n_lines           = n_lines;
  h.n_reads           = reads.n;
  h.n_unitigs         = unitigs.n;
  h.n_matches         = m.n;
  h.read_name_bytes   = reads.n_bytes;
  h.unitig_name_bytes = unitigs.n_bytes;
  h.blob_bytes        = shard_layout(h).total;
  return h;
}

// all sections of the blob in ONE launch (sixteen cudaMemcpyAsync calls cost more host time than the copy itself)
struct PackJob {
  const void *src[16];
  uint64_t    dst_off[16], bytes[16];
  int         n;
};

__global__ void __launch_bounds__(256) pack_blob_kernel(PackJob job, char *__restrict__ blob) {
  pdl_sync();
  for (int k = 0; k < job.n; ++k) {
    const char    *src = static_cast<const char *>(job.src[k]);
    char          *dst = blob + job.dst_off[k];
    const uint64_t nb  = job.bytes[k];
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) { // sections start 16-byte aligned in the blob
      const uint64_t nv = nb >> 4;
      for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nv; i += uint64_t(gridDim.x) * blockDim.x)
        reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(src)[i];
      for (uint64_t i = (nv << 4) + uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nb;
           i += uint64_t(gridDim.x) * blockDim.x)
        dst[i] = src[i];
    } else {
      for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nb; i += uint64_t(gridDim.x) * blockDim.x)
        dst[i] = src[i];
    }
  }
}

static PackJob pack_job(const msgpu_shard_header &h, const ReadTable &reads, const UnitigTable &unitigs, const MatchSoA &m) {
  const ShardLayout L = shard_layout(h);
  PackJob           job{};
  auto put = [&](uint64_t off, const void *src, uint64_t bytes) {
    if (!bytes) return;
    job.src[job.n]     = src;
    job.dst_off[job.n] = off;
    job.bytes[job.n]   = bytes;
    ++job.n;
  };
  const uint64_t R = h.n_reads, U = h.n_unitigs, M = h.n_matches;
  put(L.read_nlen, reads.nlen, R * 4);
  put(L.read_line, reads.line, R * 4);
  put(L.read_name_off, reads.name_off, (R + 1) * 8);
  put(L.read_names, reads.names, h.read_name_bytes);
  put(L.unitig_name_off, unitigs.name_off, (U + 1) * 8);
  put(L.unitig_names, unitigs.names, h.unitig_name_bytes);
  put(L.m_unitig, m.unitig, M * 4);
  put(L.m_read, m.read, M * 4);
  put(L.m_ir0, m.ir0, M * 4);
  put(L.m_ir1, m.ir1, M * 4);
  put(L.m_nr0, m.nr0, M * 4);
  put(L.m_nr1, m.nr1, M * 4);
  put(L.m_score, m.score, M * 4);
  put(L.m_line, m.line, M * 4);
  put(L.m_flags, m.flags, M);
  return job;
}

void shard_export(cudaStream_t st, const msgpu_shard_header &h, const ReadTable &reads, const UnitigTable &unitigs,
                  const MatchSoA &m, void *dev_blob) {
  char *b = static_cast<char *>(dev_blob);
  MSGPU_CUDA(cudaMemcpyAsync(b, &h, sizeof(h), cudaMemcpyHostToDevice, st));
  const PackJob job = pack_job(h, reads, unitigs, m);
  if (job.n) MSGPU_LAUNCH("pack_blob", st, launch_k(pack_blob_kernel, grid_for(h.blob_bytes / 16 + 1, 256, 2), 256, 0, st, job, b));
}

// The export AND the all-gather in one kernel: every 16-byte unit of the blob is read once from this GPU's tables and
// stored into this rank's slot of EVERY rank's receive buffer — peer-mapped device memory, so the stores to the other
// ranks travel over NVLink / NVSwitch as posted writes while the kernel keeps reading.  No staging copy of the blob, no
// collective launch; the caller orders the step with one barrier across the ranks afterwards.  The header rides in the
// kernel's arguments and is written first (always: a rank whose blob does not fit its slot still tells the others how
// big it is).
struct PeerSlots {
  char    *slot[kMaxPeers];
  uint32_t n;
};

__global__ void __launch_bounds__(256) pack_blob_peers_kernel(PackJob job, PeerSlots dst, msgpu_shard_header h) {
  pdl_sync();
  const uint64_t tid = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x, nth = uint64_t(gridDim.x) * blockDim.x;
  if (tid < sizeof(msgpu_shard_header) / 16) { // the header: four 16-byte units
    const uint4 v = reinterpret_cast<const uint4 *>(&h)[tid];
    for (uint32_t d = 0; d < dst.n; ++d) reinterpret_cast<uint4 *>(dst.slot[d])[tid] = v;
  }
  for (int k = 0; k < job.n; ++k) {
    const char    *src = static_cast<const char *>(job.src[k]);
    const uint64_t off = job.dst_off[k], nb = job.bytes[k];
    uint64_t       done = 0;
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) { // sections start 16-byte aligned in the blob
      const uint64_t nv = nb >> 4;
      for (uint64_t i = tid; i < nv; i += nth) {
        const uint4 v = reinterpret_cast<const uint4 *>(src)[i];
        for (uint32_t d = 0; d < dst.n; ++d) reinterpret_cast<uint4 *>(dst.slot[d] + off)[i] = v;
      }
      done = nv << 4;
    }
    for (uint64_t i = done + tid; i < nb; i += nth) {
      const char v = src[i];
      for (uint32_t d = 0; d < dst.n; ++d) dst.slot[d][off + i] = v;
    }
  }
}

void shard_export_peers(cudaStream_t st, const msgpu_shard_header &h, const ReadTable &reads, const UnitigTable &unitigs,
                        const MatchSoA &m, void *const *dev_slots, uint32_t n_slots, bool with_sections) {
  PeerSlots dst{};
  dst.n = n_slots;
  for (uint32_t d = 0; d < n_slots; ++d) dst.slot[d] = static_cast<char *>(dev_slots[d]);
  PackJob job = pack_job(h, reads, unitigs, m);
  if (!with_sections) job.n = 0;
  MSGPU_LAUNCH("pack_blob_peers", st, launch_k(pack_blob_peers_kernel, grid_for((with_sections ? h.blob_bytes : 0) / 16 + 1, 256, 2), 256, 0, st, job, dst, h));
}

// ---------------------------------------------------------------------------------------------------------------------
// merge
// ---------------------------------------------------------------------------------------------------------------------
struct ShardView { // one rank's blob, as device pointers, plus where its rows start in the concatenated tables
  const int32_t  *read_nlen;
  const uint32_t *read_line;
  const uint64_t *name_off[2]; // [0] reads, [1] unitigs
  const char     *names[2];
  const uint32_t *m_unitig, *m_read;
  const int32_t  *m_ir0, *m_ir1, *m_nr0, *m_nr1, *m_score;
  const uint32_t *m_line;
  const double   *m_rratio;
  const uint8_t  *m_flags;
  uint64_t        base[2]; // first entry of this shard in the concatenated read / unitig tables
  uint64_t        n[2];
  uint64_t        match_base, n_matches, line_base;
};

// entry e of table `which` -> (shard, local index)
__device__ __forceinline__ uint32_t shard_of_entry(const ShardView *sh, uint32_t n_shards, int which, uint64_t e,
                                                   uint64_t *local) {
  uint32_t s = 0;
  while (s + 1 < n_shards && sh[s + 1].base[which] <= e) ++s;
  *local = e - sh[s].base[which];
  return s;
}

__device__ __forceinline__ uint64_t entry_hash(const char *s, uint32_t n, uint64_t seed) {
  uint64_t h = 1469598103934665603ull ^ seed;
  for (uint32_t i = 0; i < n; ++i) h = (h ^ (unsigned char)s[i]) * 1099511628211ull;
  return mix64(h) | 1ull; // 0 is the empty-slot marker
}

struct EntryTable { // open addressing on the 64-bit name hash; the slot keeps the smallest entry index naming it
  unsigned long long *keys;
  uint32_t           *min_entry;
  uint32_t           *id;
  uint64_t            mask;
};

__global__ void __launch_bounds__(256)
    entry_insert_kernel(const ShardView *__restrict__ sh, uint32_t n_shards, int which, uint64_t n_entries,
                        uint64_t seed, EntryTable t, const char **__restrict__ name_ptr, uint32_t *__restrict__ name_len,
                        uint32_t *__restrict__ slot) {
  pdl_sync();
  for (uint64_t e = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n_entries;
       e += uint64_t(gridDim.x) * blockDim.x) {
    uint64_t       i;
    const uint32_t s   = shard_of_entry(sh, n_shards, which, e, &i);
    const uint64_t b   = sh[s].name_off[which][i], en = sh[s].name_off[which][i + 1];
    const char    *p   = sh[s].names[which] + b;
    const uint32_t len = uint32_t(en - b);
    name_ptr[e]        = p;
    name_len[e]        = len;
    unsigned long long h = entry_hash(p, len, seed);
    uint64_t           q = (h >> 1) & t.mask;
    for (;;) {
      unsigned long long prev = atomicCAS(&t.keys[q], 0ull, h);
      if (prev == 0ull || prev == h) break;
      q = (q + 1) & t.mask;
    }
    atomicMin(&t.min_entry[q], uint32_t(e));
    slot[e] = uint32_t(q);
  }
}

// every entry compares its name with the first entry of its slot (a 64-bit hash collision between different names
// restarts the interning with another seed) and learns whether it is the first appearance
__global__ void __launch_bounds__(256)
    entry_verify_kernel(uint64_t n_entries, EntryTable t, const char *const *__restrict__ name_ptr,
                        const uint32_t *__restrict__ name_len, const uint32_t *__restrict__ slot,
                        uint8_t *__restrict__ first, uint32_t *__restrict__ collision) {
  pdl_sync();
  for (uint64_t e = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n_entries;
       e += uint64_t(gridDim.x) * blockDim.x) {
    const uint32_t rep = t.min_entry[slot[e]];
    first[e]           = rep == uint32_t(e);
    if (rep != uint32_t(e)) {
      const uint32_t n     = name_len[e];
      bool           clash = name_len[rep] != n;
      const char    *a = name_ptr[e], *b = name_ptr[rep];
      for (uint32_t i = 0; i < n && !clash; ++i) clash = a[i] != b[i];
      if (clash) *collision = 1;
    }
  }
}

__global__ void __launch_bounds__(256)
    entry_assign_kernel(const ShardView *__restrict__ sh, uint32_t n_shards, int which, uint64_t n_entries,
                        EntryTable t, const uint32_t *__restrict__ slot, const uint8_t *__restrict__ first,
                        const uint32_t *__restrict__ rank, const char *const *__restrict__ name_ptr,
                        const uint32_t *__restrict__ name_len, const char **__restrict__ out_name_ptr,
                        uint32_t *__restrict__ out_name_len, int32_t *__restrict__ read_nlen,
                        uint32_t *__restrict__ read_line) {
  pdl_sync();
  for (uint64_t e = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n_entries;
       e += uint64_t(gridDim.x) * blockDim.x) {
    if (!first[e]) continue;
    const uint32_t id = rank[e];
    t.id[slot[e]]     = id;
    out_name_ptr[id]  = name_ptr[e];
    out_name_len[id]  = name_len[e];
    if (which == 0) {
      uint64_t       i;
      const uint32_t s = shard_of_entry(sh, n_shards, 0, e, &i);
      read_nlen[id]    = sh[s].read_nlen[i];                             // Vertex(id, nanoporeLength, lineIdx)
      read_line[id]    = uint32_t(sh[s].read_line[i] + sh[s].line_base); // of the globally first accepted line
    }
  }
}

__global__ void __launch_bounds__(256)
    entry_gid_kernel(uint64_t n_entries, EntryTable t, const uint32_t *__restrict__ slot, uint32_t *__restrict__ gid,
                     const uint8_t *__restrict__ first, uint8_t *__restrict__ shared /* may be null */) {
  pdl_sync();
  for (uint64_t e = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n_entries;
       e += uint64_t(gridDim.x) * blockDim.x) {
    const uint32_t id = t.id[slot[e]];
    gid[e]            = id;
    if (shared && !first[e]) shared[id] = 1; // the name occurs in more than one shard
  }
}

__global__ void __launch_bounds__(256)
    gather_names_abs_kernel(const char *const *__restrict__ src, const uint64_t *__restrict__ dst_off, uint64_t n,
                            char *__restrict__ out) {
  pdl_sync();
  uint64_t w  = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; // one warp per name
  uint64_t nw = (uint64_t(gridDim.x) * blockDim.x) >> 5;
  for (; w < n; w += nw) {
    uint64_t b = dst_off[w], e = dst_off[w + 1];
    for (uint64_t i = lane_id(); i < e - b; i += 32) out[b + i] = src[w][i];
  }
}

// concatenated records with global ids and global line numbers, shard after shard (= global line order inside a unitig)
__global__ void __launch_bounds__(256)
    remap_matches_kernel(const ShardView *__restrict__ sh, uint32_t n_shards, uint64_t n_total,
                         const uint32_t *__restrict__ gid_read, const uint32_t *__restrict__ gid_unitig, MatchSoA c) {
  pdl_sync();
  for (uint64_t g = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; g < n_total;
       g += uint64_t(gridDim.x) * blockDim.x) {
    uint32_t s = 0;
    while (s + 1 < n_shards && sh[s + 1].match_base <= g) ++s;
    const ShardView &v = sh[s];
    const uint64_t   k = g - v.match_base;
    c.unitig[g] = gid_unitig[v.base[1] + v.m_unitig[k]];
    c.read[g]   = gid_read[v.base[0] + v.m_read[k]];
    c.ir0[g]    = v.m_ir0[k];
    c.ir1[g]    = v.m_ir1[k];
    c.nr0[g]    = v.m_nr0[k];
    c.nr1[g]    = v.m_nr1[k];
    c.score[g]  = v.m_score[k];
    c.line[g]   = uint32_t(v.m_line[k] + v.line_base);
    c.flags[g]  = v.m_flags[k];
    // rRatio = double(illuminaLength) / double(c8 - c7) (BlastFileReader.cpp:118-119): the same division the ingest
    // did on the same integers, so it travels as those integers only
    c.rratio[g] = __ddiv_rn(double(v.m_ir1[k] - v.m_ir0[k] + 1), double(v.m_nr1[k] - v.m_nr0[k] + 1));
  }
}

// per (unitig, read) keep the record with the smallest global line (MatchMap.cpp:64-73) — across chunks this time.
// Every chunk is deduplicated already, so only the records of unitigs that occur in several chunks (the groups cut
// by a chunk boundary, or everything if the PAF is not grouped) can have a duplicate: the others skip the table.
constexpr uint32_t kNoSlot = 0xffffffffu;
__global__ void __launch_bounds__(256)
    shard_dedup_insert_kernel(MatchSoA c, EntryTable t, const uint8_t *__restrict__ shared_unitig,
                              uint32_t *__restrict__ slot) {
  pdl_sync();
  for (uint64_t g = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; g < c.n; g += uint64_t(gridDim.x) * blockDim.x) {
    if (!shared_unitig[c.unitig[g]]) {
      slot[g] = kNoSlot;
      continue;
    }
    unsigned long long key = ((unsigned long long)c.unitig[g] << 32 | c.read[g]) + 1ull;
    uint64_t           q   = mix64(key) & t.mask;
    for (;;) {
      unsigned long long prev = atomicCAS(&t.keys[q], 0ull, key);
      if (prev == 0ull || prev == key) break;
      q = (q + 1) & t.mask;
    }
    atomicMin(&t.min_entry[q], c.line[g]);
    slot[g] = uint32_t(q);
  }
}

struct KeepFlag {
  const uint32_t *line, *slot, *min_line;
  __device__ uint32_t operator()(uint64_t g) const {
    const uint32_t q = slot[g];
    return (q == kNoSlot || min_line[q] == line[g]) ? 1u : 0u;
  }
};

__global__ void __launch_bounds__(256)
    shard_compact_kernel(MatchSoA c, KeepFlag keep, const uint32_t *__restrict__ pos, MatchSoA m) {
  pdl_sync();
  for (uint64_t g = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; g < c.n; g += uint64_t(gridDim.x) * blockDim.x) {
    if (!keep(g)) continue;
    const uint32_t o = pos[g];
    m.unitig[o] = c.unitig[g];
    m.read[o]   = c.read[g];
    m.ir0[o]    = c.ir0[g];
    m.ir1[o]    = c.ir1[g];
    m.nr0[o]    = c.nr0[g];
    m.nr1[o]    = c.nr1[g];
    m.score[o]  = c.score[g];
    m.line[o]   = c.line[g];
    m.flags[o]  = c.flags[g];
    m.rratio[o] = c.rratio[g];
  }
}

static uint64_t table_capacity(uint64_t expect) {
  uint64_t p = 1024;
  while (p < expect * 2 + 2) p <<= 1;
  return p;
}

static EntryTable make_entry_table(Arena &arena, cudaStream_t st, uint64_t expect) {
  EntryTable t;
  uint64_t   cap = table_capacity(expect);
  t.keys         = arena.alloc<unsigned long long>(cap);
  t.min_entry    = arena.alloc<uint32_t>(cap);
  t.id           = arena.alloc<uint32_t>(cap);
  t.mask         = cap - 1;
  MSGPU_CUDA(cudaMemsetAsync(t.keys, 0, cap * sizeof(unsigned long long), st));
  MSGPU_CUDA(cudaMemsetAsync(t.min_entry, 0xff, cap * sizeof(uint32_t), st));
  return t;
}

template <class T> static T read_back(cudaStream_t st, const T *d) {
  T h;
  MSGPU_CUDA(cudaMemcpyAsync(&h, d, sizeof(T), cudaMemcpyDeviceToHost, st));
  MSGPU_CUDA(cudaStreamSynchronize(st));
  return h;
}

struct MergeStatus { // device-resident, read back once per merge
  unsigned long long name_bytes[2];
  unsigned           n_ids[2];
  unsigned           collision, unsorted, n_matches, pad;
};

struct IdNameLen { // lengths of the first *d_n names, 0 behind them (the scan runs over the capacity)
  const uint32_t *len;
  const unsigned *d_n;
  __device__ uint64_t operator()(uint64_t i) const { return i < *d_n ? uint64_t(len[i]) : 0ull; }
};

__global__ void __launch_bounds__(256)
    gather_names_abs_n_kernel(const char *const *__restrict__ src, const uint64_t *__restrict__ dst_off,
                              const unsigned *__restrict__ d_n, char *__restrict__ out) {
  pdl_sync();
  const uint64_t n  = *d_n;
  uint64_t       w  = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; // one warp per name
  uint64_t       nw = (uint64_t(gridDim.x) * blockDim.x) >> 5;
  for (; w < n; w += nw) {
    uint64_t b = dst_off[w], e = dst_off[w + 1];
    for (uint64_t i = lane_id(); i < e - b; i += 32) out[b + i] = src[w][i];
  }
}

__global__ void __launch_bounds__(256)
    shard_check_grouped_kernel(const uint32_t *__restrict__ unitig, const unsigned *__restrict__ d_n,
                               unsigned *unsorted) {
  pdl_sync();
  const uint64_t n = *d_n;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x + 1; i < n; i += uint64_t(gridDim.x) * blockDim.x)
    if (unitig[i] < unitig[i - 1]) *unsorted = 1;
}

// interning of one concatenated name table: fills gid[entry] and `out`; every size is bounded by the entry count and
// the name bytes of the headers, so nothing comes back to the host here (counts and the collision flag: `status`)
static void merge_names(Arena &arena, cudaStream_t st, const ShardView *d_sh, uint32_t n_shards, int which,
                        uint64_t n_entries, uint64_t name_bytes_bound, uint64_t seed, MergeStatus *d_status,
                        uint32_t *gid, NameTable &out, int32_t *&read_nlen, uint32_t *&read_line,
                        uint8_t *shared = nullptr /* per global id: 1 iff the name occurs in several shards */) {
  const char **name_ptr = arena.alloc<const char *>(n_entries);
  uint32_t    *name_len = arena.alloc<uint32_t>(n_entries), *slot = arena.alloc<uint32_t>(n_entries);
  uint32_t    *rank = arena.alloc<uint32_t>(n_entries);
  uint8_t     *first = arena.alloc<uint8_t>(n_entries);
  const int    g = grid_for(n_entries, 256);
  EntryTable   t = make_entry_table(arena, st, n_entries);
  const char **id_name_ptr = arena.alloc<const char *>(n_entries);
  uint32_t    *id_name_len = arena.alloc<uint32_t>(n_entries);
  if (which == 0) {
    read_nlen = arena.alloc<int32_t>(n_entries);
    read_line = arena.alloc<uint32_t>(n_entries);
  }
  if (n_entries) {
    MSGPU_LAUNCH("entry_insert", st, launch_k(entry_insert_kernel, g, 256, 0, st, d_sh, n_shards, which, n_entries, seed, t, name_ptr, name_len, slot));
    MSGPU_LAUNCH("entry_verify", st, launch_k(entry_verify_kernel, g, 256, 0, st, n_entries, t, name_ptr, name_len, slot, first, &d_status->collision));
  }
  exclusive_scan<uint32_t>(arena, st, LoadAs<uint8_t, uint32_t>{first}, n_entries, rank, &d_status->n_ids[which]);
  if (n_entries) {
    MSGPU_LAUNCH("entry_assign", st, launch_k(entry_assign_kernel, g, 256, 0, st, d_sh, n_shards, which, n_entries, t, slot, first, rank, name_ptr, name_len, id_name_ptr, id_name_len, read_nlen, read_line));
    MSGPU_LAUNCH("entry_gid", st, launch_k(entry_gid_kernel, g, 256, 0, st, n_entries, t, slot, gid, first, shared));
  }
  out.name_off = arena.alloc<uint64_t>(n_entries + 1);
  exclusive_scan<uint64_t>(arena, st, IdNameLen{id_name_len, &d_status->n_ids[which]}, n_entries + 1, out.name_off,
                           reinterpret_cast<uint64_t *>(&d_status->name_bytes[which]));
  out.names = arena.alloc<char>(name_bytes_bound);
  if (n_entries) MSGPU_LAUNCH("gather_names", st, launch_k(gather_names_abs_n_kernel, grid_for(n_entries * 32, 256, 2), 256, 0, st, id_name_ptr, out.name_off, &d_status->n_ids[which], out.names));
}

void run_shard_merge(Arena &arena, cudaStream_t st, uint32_t n_shards, const msgpu_shard_header *headers,
                     const void *const *dev_blobs, uint64_t *n_lines_out, ReadTable &reads, UnitigTable &unitigs,
                     MatchSoA &m) {
  reads   = ReadTable{};
  unitigs = UnitigTable{};
  m       = MatchSoA{};
  std::vector<ShardView> hv(n_shards);
  uint64_t               rb = 0, ub = 0, mb = 0, lb = 0, rbytes = 0, ubytes = 0;
  for (uint32_t s = 0; s < n_shards; ++s) {
    const msgpu_shard_header &h = headers[s];
    if (h.magic != kShardMagic || h.blob_bytes != shard_layout(h).total)
      throw BuildError{MSGPU_ERR_INVALID_ARG, "shard " + std::to_string(s) + ": not a msgpu shard blob"};
    const ShardLayout L = shard_layout(h);
    const char       *b = static_cast<const char *>(dev_blobs[s]);
    ShardView        &v = hv[s];
    v.read_nlen   = reinterpret_cast<const int32_t *>(b + L.read_nlen);
    v.read_line   = reinterpret_cast<const uint32_t *>(b + L.read_line);
    v.name_off[0] = reinterpret_cast<const uint64_t *>(b + L.read_name_off);
    v.names[0]    = b + L.read_names;
    v.name_off[1] = reinterpret_cast<const uint64_t *>(b + L.unitig_name_off);
    v.names[1]    = b + L.unitig_names;
    v.m_unitig    = reinterpret_cast<const uint32_t *>(b + L.m_unitig);
    v.m_read      = reinterpret_cast<const uint32_t *>(b + L.m_read);
    v.m_ir0       = reinterpret_cast<const int32_t *>(b + L.m_ir0);
    v.m_ir1       = reinterpret_cast<const int32_t *>(b + L.m_ir1);
    v.m_nr0       = reinterpret_cast<const int32_t *>(b + L.m_nr0);
    v.m_nr1       = reinterpret_cast<const int32_t *>(b + L.m_nr1);
    v.m_score     = reinterpret_cast<const int32_t *>(b + L.m_score);
    v.m_line      = reinterpret_cast<const uint32_t *>(b + L.m_line);
    v.m_rratio    = reinterpret_cast<const double *>(b + L.m_rratio);
    v.m_flags     = reinterpret_cast<const uint8_t *>(b + L.m_flags);
    v.base[0]     = rb;
    v.base[1]     = ub;
    v.n[0]        = h.n_reads;
    v.n[1]        = h.n_unitigs;
    v.match_base  = mb;
    v.n_matches   = h.n_matches;
    v.line_base   = lb;
    rb += h.n_reads;
    ub += h.n_unitigs;
    mb += h.n_matches;
    lb += h.n_lines;
    rbytes += h.read_name_bytes;
    ubytes += h.unitig_name_bytes;
  }
  if (lb >= 0xfffffff0ull || rb >= 0xfffffff0ull || mb >= 0xfffffff0ull)
    throw BuildError{MSGPU_ERR_LIMIT, "more than 2^32-16 lines, read entries or matches over all shards"};
  *n_lines_out = lb;
  ShardView *d_sh = arena.alloc<ShardView>(n_shards);
  if (n_shards * sizeof(ShardView) + 256 <= Arena::kHostScratch) {
    // through the second part of the pinned scratch buffer: no synchronisation (the first 256 bytes are where the
    // status comes back at the end of the merge, after this copy in stream order)
    char *pin = static_cast<char *>(arena.host_scratch()) + 256;
    memcpy(pin, hv.data(), n_shards * sizeof(ShardView));
    MSGPU_CUDA(cudaMemcpyAsync(d_sh, pin, n_shards * sizeof(ShardView), cudaMemcpyHostToDevice, st));
  } else {
    MSGPU_CUDA(cudaMemcpyAsync(d_sh, hv.data(), n_shards * sizeof(ShardView), cudaMemcpyHostToDevice, st));
    MSGPU_CUDA(cudaStreamSynchronize(st)); // hv is pageable host memory
  }

  MergeStatus *d_status = arena.alloc<MergeStatus>(1);
  MergeStatus  h{};
  for (int attempt = 0;; ++attempt) {
    if (attempt == 4) throw BuildError{MSGPU_ERR_INTERNAL, "name hash collisions persisted over 4 seeds"};
    MSGPU_CUDA(cudaMemsetAsync(d_status, 0, sizeof(MergeStatus), st));
    const uint64_t seed = 0x243f6a8885a308d3ull * uint64_t(attempt + 1);

    // ---- global ids
    uint32_t *gid_r = arena.alloc<uint32_t>(rb), *gid_u = arena.alloc<uint32_t>(ub);
    int32_t  *no_nlen = nullptr;
    uint32_t *no_line = nullptr;
    merge_names(arena, st, d_sh, n_shards, 0, rb, rbytes, seed, d_status, gid_r, reads, reads.nlen, reads.line);
    uint8_t *shared_u = arena.alloc<uint8_t>(ub + 1);
    MSGPU_CUDA(cudaMemsetAsync(shared_u, 0, ub + 1, st));
    merge_names(arena, st, d_sh, n_shards, 1, ub, ubytes, seed, d_status, gid_u, unitigs, no_nlen, no_line, shared_u);

    // ---- records: renumber, deduplicate across chunks (capacity = all records), group by unitig.  The final table is
    // allocated first; everything after the mark is scratch of the merge and is handed back before the pairing
    // allocates (at 10^8 .. 10^9 records the scratch — a second record table, the dedup table — is the larger part)
    alloc_match_soa(arena, m, mb);
    const Arena::Mark scratch = arena.mark();
    MatchSoA c;
    alloc_match_soa(arena, c, mb);
    const int g = grid_for(mb, 256);
    if (mb) MSGPU_LAUNCH("remap_matches", st, launch_k(remap_matches_kernel, g, 256, 0, st, d_sh, n_shards, mb, gid_r, gid_u, c));
    EntryTable dt   = make_entry_table(arena, st, mb);
    uint32_t  *slot = arena.alloc<uint32_t>(mb), *pos = arena.alloc<uint32_t>(mb);
    if (mb) MSGPU_LAUNCH("dedup_insert", st, launch_k(shard_dedup_insert_kernel, g, 256, 0, st, c, dt, shared_u, slot));
    KeepFlag keep{c.line, slot, dt.min_entry};
    exclusive_scan<uint32_t>(arena, st, keep, mb, pos, &d_status->n_matches);
    if (mb) {
      MSGPU_LAUNCH("compact_matches", st, launch_k(shard_compact_kernel, g, 256, 0, st, c, keep, pos, m));
      MSGPU_LAUNCH("check_grouped", st, launch_k(shard_check_grouped_kernel, g, 256, 0, st, m.unitig, &d_status->n_matches, &d_status->unsorted));
    }
    // ---- the one round trip of the merge
    static_assert(sizeof(MergeStatus) <= 256, "host scratch layout");
    void *pinned = arena.host_scratch();
    MSGPU_CUDA(cudaMemcpyAsync(pinned, d_status, sizeof(h), cudaMemcpyDeviceToHost, st));
    MSGPU_CUDA(cudaStreamSynchronize(st));
    memcpy(&h, pinned, sizeof(h));
    arena.release_to(scratch);
    if (!h.collision) break;
  }
  reads.n         = h.n_ids[0];
  reads.n_bytes   = h.name_bytes[0];
  unitigs.n       = h.n_ids[1];
  unitigs.n_bytes = h.name_bytes[1];
  m.n             = h.n_matches;
  group_by_unitig(arena, st, m, unitigs.n, h.unsorted ? 1 : 0);
  build_group_offsets(arena, st, m, unitigs);
}

} // namespace msgpu
