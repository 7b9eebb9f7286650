// seq.cu — sequence store resident in HBM + batched slice / reverse-complement (kernel K8).
//
// Replaces (file:line relative to the reference tree):
//   SequenceAccessor::buildIndex / _buildNanoporeIdx / _buildIlluminaIdx   SequenceAccessor.cpp:114-127,143-231
//   SequenceAccessor::get{Nanopore,Illumina}Sequence + getSequenceFromFile SequenceAccessor.cpp:54-69,129-139
//   strSlice, getReverseComplement, getIlluminaSequence, getNanoporeSequence   SequenceUtils.cpp:27-85
//
// The reference re-reads a whole FASTA/FASTQ record from disk, under one mutex per file, for every slice it needs
// (a 50-base flank of a 100 kb read costs a 100 kb fread + a whitespace strip).  Here both files are read ONCE: the
// record structure is found on the host (a line state machine with the reference's exact rules), the raw bytes go to
// HBM, a compaction kernel drops the whitespace of the sequence lines (the header, '+' and quality lines are simply
// not inside any record range), and every later fetch is a gather over that resident copy — batched, one warp per
// request, no locks.
//
// Bytes are kept as ASCII so that every byte the reference would return (N, lower case, IUPAC codes) comes back
// unchanged; only upper-case ACGT are complemented (SequenceUtils.cpp:46-57).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "internal.h"
#include "scan.cuh"
#include "seq.h"

namespace msgpu {

// ---------------------------------------------------------------------------------------------------------------------
// host: files and record structure
// ---------------------------------------------------------------------------------------------------------------------
MappedFile::MappedFile(const char *path) {
  int fd = ::open(path, O_RDONLY | O_CLOEXEC);
  if (fd < 0) return;
  struct stat sb {};
  if (fstat(fd, &sb) != 0) {
    ::close(fd);
    return;
  }
  size = size_t(sb.st_size);
  if (size) {
    void *m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) {
      ::close(fd);
      size = 0;
      return;
    }
    data = static_cast<const char *>(m);
  }
  ::close(fd);
  ok = true;
}
MappedFile::~MappedFile() {
  if (data) munmap(const_cast<char *>(data), size);
}

bool path_is_fastq(const char *path) {
  std::string p(path);
  auto        dot = p.find_last_of('.');
  std::string ext = p.substr(dot == std::string::npos ? 0 : dot + 1);
  for (auto &c : ext) c = char(std::tolower((unsigned char)c));
  return ext != "fa" && ext != "fasta";
}

namespace {

struct RawRecord {
  uint64_t name_off, name_len; // id = header after the marker up to the first whitespace (SequenceAccessor.cpp:80-85)
  uint64_t begin, end;         // raw byte range of the sequence lines, newlines included
};

// The reference's index loop as a line state machine.  FASTA: a record runs from the line after a '>' header to the
// next line that starts with '>'.  FASTQ: from the line after an '@' header to the next line that starts with '+';
// then lines are skipped until one starts with '@' (so a quality line that starts with '@' is taken for a header —
// the reference does the same, SequenceAccessor.cpp:184-187).
std::vector<RawRecord> index_records(const char *d, size_t n, bool fastq) {
  const char             desc = fastq ? '@' : '>', split = fastq ? '+' : '>';
  std::vector<RawRecord> recs;
  enum { SEEK, SEQ, SKIP } state = SEEK;
  RawRecord cur{};
  auto      open_record = [&](size_t ls, size_t le) {
    cur          = RawRecord{};
    cur.name_off = ls + 1;
    size_t e     = cur.name_off;
    while (e < le && d[e] != '\0' && !std::isspace((unsigned char)d[e])) ++e;
    cur.name_len = e - cur.name_off;
    cur.begin = cur.end = le;
    state               = SEQ;
  };
  auto close_record = [&](size_t at) {
    cur.end = at;
    // getSequenceFromFile builds a C string from the bytes it read: a NUL ends the record (SequenceAccessor.cpp:63)
    if (const void *z = cur.end > cur.begin ? std::memchr(d + cur.begin, 0, cur.end - cur.begin) : nullptr)
      cur.end = size_t(static_cast<const char *>(z) - d);
    recs.push_back(cur);
  };
  size_t ls = 0;
  while (ls < n) {
    const void *nl = std::memchr(d + ls, '\n', n - ls);
    size_t      le = nl ? size_t(static_cast<const char *>(nl) - d) + 1 : n;
    char        c0 = d[ls];
    switch (state) {
    case SEEK:
    case SKIP:
      if (c0 == desc) open_record(ls, le);
      break;
    case SEQ:
      if (c0 == split) {
        close_record(ls);
        if (fastq) state = SKIP; else open_record(ls, le);
      }
      break;
    }
    ls = le;
  }
  if (state == SEQ) close_record(n); // end of file ends the last record (readline() == -1 branch)
  return recs;
}

} // namespace

// ---------------------------------------------------------------------------------------------------------------------
// device: whitespace strip (stream compaction of the bytes inside record ranges)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kStripBlock = 256;
constexpr int kStripPer   = 16;
constexpr int kStripTile  = kStripBlock * kStripPer;

__device__ __forceinline__ bool is_space_byte(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// index of the last record with begin <= pos, or -1
__device__ __forceinline__ int64_t find_record(const uint64_t *__restrict__ rec_begin, uint64_t n_rec, uint64_t pos) {
  int64_t lo = -1, hi = int64_t(n_rec);
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (rec_begin[mid] <= pos) lo = mid; else hi = mid;
  }
  return lo;
}

// WRITE = false: count kept bytes per tile.  WRITE = true: write them and record where each record starts / ends in
// the packed stream.  Positions run to n INCLUSIVE so that a record ending at end-of-file sees its end event.
template <bool WRITE>
__global__ void __launch_bounds__(kStripBlock)
    strip_kernel(const char *__restrict__ raw, uint64_t n, const uint64_t *__restrict__ rec_begin,
                 const uint64_t *__restrict__ rec_end, uint64_t n_rec, uint64_t *__restrict__ tile_cnt_or_off,
                 char *__restrict__ packed, uint64_t *__restrict__ packed_begin, uint64_t *__restrict__ packed_end) {
  pdl_sync();
  __shared__ uint32_t sm[33];
  uint64_t            b0 = uint64_t(blockIdx.x) * kStripTile + uint64_t(threadIdx.x) * kStripPer;
  unsigned            keep = 0;
  int64_t             r0   = b0 <= n ? find_record(rec_begin, n_rec, b0) : -1;
  {
    int64_t r = r0;
    for (int k = 0; k < kStripPer; ++k) {
      uint64_t pos = b0 + k;
      if (pos >= n) break;
      while (r + 1 < int64_t(n_rec) && rec_begin[r + 1] <= pos) ++r;
      if (r >= 0 && pos < rec_end[r] && !is_space_byte((unsigned char)raw[pos])) keep |= 1u << k;
    }
  }
  uint32_t total;
  uint32_t rank = block_exclusive_scan<uint32_t>(__popc(keep), sm, &total);
  if (!WRITE) {
    if (threadIdx.x == 0) tile_cnt_or_off[blockIdx.x] = total;
    return;
  }
  uint64_t o = tile_cnt_or_off[blockIdx.x] + rank;
  int64_t  r = r0;
  for (int k = 0; k < kStripPer; ++k) {
    uint64_t pos = b0 + k;
    if (pos > n) break;
    while (r + 1 < int64_t(n_rec) && rec_begin[r + 1] <= pos) ++r;
    if (r >= 0) {
      if (pos == rec_begin[r]) packed_begin[r] = o;
      if (pos == rec_end[r]) packed_end[r] = o;
    }
    // a record's end can coincide with the position where the previous record index still applies
    if (r >= 1 && pos == rec_end[r - 1]) packed_end[r - 1] = o;
    if (keep >> k & 1u) packed[o++] = raw[pos];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// device: slice + reverse complement
// ---------------------------------------------------------------------------------------------------------------------
// strSlice(original, i, j) (SequenceUtils.cpp:27-38): negative indices count from the end; then the window is
// [start, end] INCLUSIVE with start = max(0,i), end = max(min(size, max(0,j)), size_t(i)); substr clamps the count
// at the end of the text and throws if start > size.  Returns false where the reference throws.
__host__ __device__ __forceinline__ bool slice_window(uint64_t size, int i, int j, uint64_t *start, uint64_t *count) {
  int isz = int(size);
  if (i < 0) i = isz + i;
  if (j < 0) j = isz + j;
  uint64_t s  = uint64_t(i > 0 ? i : 0);
  uint64_t jj = uint64_t(j > 0 ? j : 0);
  uint64_t e  = jj < size ? jj : size;
  uint64_t ii = uint64_t((long long)i); // a still-negative i wraps like static_cast<std::size_t>(i)
  if (ii > e) e = ii;
  if (s > size) return false;
  uint64_t c = e - s + 1; // wraps exactly like the reference's size_t arithmetic
  uint64_t a = size - s;
  *start     = s;
  *count     = c < a ? c : a;
  return true;
}

__device__ __forceinline__ char complement(char c) { // SequenceUtils.cpp:46-57: upper-case ACGT only
  switch (c) {
  case 'A': return 'T';
  case 'T': return 'A';
  case 'G': return 'C';
  case 'C': return 'G';
  default: return c;
  }
}

struct SeqIndexDev {
  const uint64_t *start[2]; // [which][id] -> offset in packed[which]; ~0 = no such sequence
  const uint64_t *len[2];
  const char     *packed[2];
  uint32_t        n_ids[2];
};

struct Window {
  uint64_t src;   // absolute offset of the first byte in the source text
  uint64_t count;
};

// status: 0 ok, 1 = the reference would throw (unknown id / out_of_range), 2 = output capacity exceeded
__global__ void __launch_bounds__(256)
    seq_window_kernel(SeqIndexDev idx, const msgpu_seq_req *__restrict__ reqs, uint64_t n,
                      Window *__restrict__ win, uint8_t *__restrict__ status) {
  pdl_sync();
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
    msgpu_seq_req q = reqs[i];
    Window        w{0, 0};
    uint8_t       st = 1;
    if (q.which <= 1 && q.id < idx.n_ids[q.which] && idx.start[q.which][q.id] != ~0ull) {
      uint64_t s, c;
      // get{Illumina,Nanopore}Sequence slice [left, right + 1] (SequenceUtils.cpp:66,79)
      if (slice_window(idx.len[q.which][q.id], q.left, q.right + 1, &s, &c)) {
        w.src   = idx.start[q.which][q.id] + s;
        w.count = c;
        st      = 0;
      }
    }
    win[i]    = w;
    status[i] = st;
  }
}

__global__ void __launch_bounds__(256)
    text_window_kernel(const uint64_t *__restrict__ text_off, const int32_t *__restrict__ left,
                       const int32_t *__restrict__ right, uint64_t n, Window *__restrict__ win,
                       uint8_t *__restrict__ status) {
  pdl_sync();
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
    uint64_t s, c;
    Window   w{0, 0};
    uint8_t  st = 1;
    if (slice_window(text_off[i + 1] - text_off[i], left[i], right[i], &s, &c)) {
      w.src   = text_off[i] + s;
      w.count = c;
      st      = 0;
    }
    win[i]    = w;
    status[i] = st;
  }
}

struct WindowCount {
  const Window *w;
  __device__ uint64_t operator()(uint64_t i) const { return w[i].count; }
};

// complement of four packed bytes: only upper-case A<->T (0x41 ^ 0x54 = 0x15) and C<->G (0x43 ^ 0x47 = 0x04) change
__device__ __forceinline__ uint32_t complement4(uint32_t w) {
  const uint32_t at = __vcmpeq4(w, 0x41414141u) | __vcmpeq4(w, 0x54545454u);
  const uint32_t cg = __vcmpeq4(w, 0x43434343u) | __vcmpeq4(w, 0x47474747u);
  return w ^ (at & 0x15151515u) ^ (cg & 0x04040404u);
}

// the 16 bytes that start at the (arbitrarily aligned) address p: two aligned 128-bit loads + funnel shifts.  Every
// lane of a warp works on the same request, so the misalignment is warp-uniform and the switch does not diverge.
__device__ __forceinline__ uint4 load16_unaligned(const char *p) {
  const uintptr_t a   = reinterpret_cast<uintptr_t>(p);
  const uint4    *q   = reinterpret_cast<const uint4 *>(a & ~uintptr_t(15));
  const unsigned  off = unsigned(a & 15u);
  const uint4     x   = __ldg(q);
  if (off == 0) return x;
  const uint4    y  = __ldg(q + 1); // holds `off` of the wanted bytes, so it lies inside the text
  const unsigned sh = (off & 3u) * 8u;
  uint32_t       a0, a1, a2, a3, a4;
  switch (off >> 2) {
  case 0: a0 = x.x, a1 = x.y, a2 = x.z, a3 = x.w, a4 = y.x; break;
  case 1: a0 = x.y, a1 = x.z, a2 = x.w, a3 = y.x, a4 = y.y; break;
  case 2: a0 = x.z, a1 = x.w, a2 = y.x, a3 = y.y, a4 = y.z; break;
  default: a0 = x.w, a1 = y.x, a2 = y.y, a3 = y.z, a4 = y.w; break;
  }
  return make_uint4(__funnelshift_r(a0, a1, sh), __funnelshift_r(a1, a2, sh), __funnelshift_r(a2, a3, sh),
                    __funnelshift_r(a3, a4, sh));
}

// One warp per request.  Forward: out[k] = src[k].  Reverse complement: out[k] = comp(src[count-1-k])
// (SequenceUtils.cpp:41-61).  The bulk moves as aligned 128-bit stores — lane l writes 16 output bytes per step, read
// from the (generally misaligned) source with two aligned 128-bit loads; in reverse the 16 bytes are byte-reversed and
// complemented four at a time in registers.  The few bytes before the first / after the last aligned output vector
// go one per lane.  src_sel(i) picks the source text of request i (0 / 1 = unitigs / reads for the store; always 0
// for caller text).
template <class SrcSel, class RevSel>
__global__ void __launch_bounds__(256)
    slice_copy_kernel(SrcSel src_sel, RevSel rev_sel, const Window *__restrict__ win,
                      const uint64_t *__restrict__ out_off, uint64_t n, uint64_t cap, char *__restrict__ out,
                      uint8_t *__restrict__ status) {
  pdl_sync();
  uint64_t       w  = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nw = (uint64_t(gridDim.x) * blockDim.x) >> 5;
  const unsigned l  = lane_id();
  for (; w < n; w += nw) {
    const Window   q = win[w];
    const uint64_t o = out_off[w];
    if (q.count == 0) continue;
    if (o + q.count > cap) {
      if (l == 0) status[w] = 2;
      continue;
    }
    const char    *src = src_sel(w) + q.src;
    char          *dst = out + o;
    const bool     rev = rev_sel(w);
    uint64_t       head = (16u - unsigned(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u; // bytes before the first aligned vector
    if (head > q.count) head = q.count;
    const uint64_t nvec = (q.count - head) >> 4, tail0 = head + (nvec << 4);
    if (!rev) {
      if (l < head) dst[l] = src[l];
      for (uint64_t v = l; v < nvec; v += 32)
        *reinterpret_cast<uint4 *>(dst + head + (v << 4)) = load16_unaligned(src + head + (v << 4));
      if (tail0 + l < q.count) dst[tail0 + l] = src[tail0 + l];
    } else {
      const char *end = src + q.count; // out[k] = comp(end[-1 - k])
      if (l < head) dst[l] = complement(end[-1 - int64_t(l)]);
      for (uint64_t v = l; v < nvec; v += 32) {
        const uint4 t = load16_unaligned(end - head - ((v + 1) << 4));
        uint4       r;
        r.x = complement4(__byte_perm(t.w, 0u, 0x0123u));
        r.y = complement4(__byte_perm(t.z, 0u, 0x0123u));
        r.z = complement4(__byte_perm(t.y, 0u, 0x0123u));
        r.w = complement4(__byte_perm(t.x, 0u, 0x0123u));
        *reinterpret_cast<uint4 *>(dst + head + (v << 4)) = r;
      }
      if (tail0 + l < q.count) dst[tail0 + l] = complement(end[-1 - int64_t(tail0 + l)]);
    }
  }
}

struct StoreSrc {
  const msgpu_seq_req *reqs;
  const char          *packed0, *packed1;
  __device__ const char *operator()(uint64_t i) const { return reqs[i].which ? packed1 : packed0; }
};
struct StoreRev {
  const msgpu_seq_req *reqs;
  __device__ bool operator()(uint64_t i) const { return reqs[i].direction == 0; } // !direction -> reverse complement
};
struct TextSrc {
  const char *text;
  __device__ const char *operator()(uint64_t) const { return text; }
};
struct TextRev {
  const uint8_t *rev;
  __device__ bool operator()(uint64_t i) const { return rev[i] != 0; }
};

// ---------------------------------------------------------------------------------------------------------------------
// the store
// ---------------------------------------------------------------------------------------------------------------------
struct Workspace {
  cudaStream_t   stream = nullptr;
  Arena          arena;
  void          *pinned = nullptr;
  size_t         pinned_cap = 0;
  void          *host(size_t bytes) {
    if (bytes > pinned_cap) {
      if (pinned) cudaFreeHost(pinned);
      pinned = nullptr;
      MSGPU_CUDA(cudaMallocHost(&pinned, bytes + bytes / 2 + 4096));
      pinned_cap = bytes + bytes / 2 + 4096;
    }
    return pinned;
  }
  ~Workspace() {
    if (pinned) cudaFreeHost(pinned);
    if (stream) cudaStreamDestroy(stream);
  }
};

struct SeqStore {
  Arena                                     arena; // packed sequences + index, lives as long as the store
  SeqIndexDev                               dev{};
  std::vector<uint64_t>                     h_len[2], h_start[2]; // by id
  std::unordered_map<std::string, uint32_t> ids[2];               // which = 0 unitigs, 1 reads
  std::mutex                                mu;
  std::vector<Workspace *>                  free_ws;
  ~SeqStore() {
    for (auto *w : free_ws) delete w;
  }
};

namespace {

// uploads one file, strips it, returns per-record packed (begin, end)
void load_file(SeqStore *s, cudaStream_t st, int which, const char *data, size_t n, bool fastq,
               const std::vector<std::string> &prime) {
  auto &ids = s->ids[which];
  for (auto const &nm : prime) ids.emplace(nm, uint32_t(ids.size())); // Registry: first request gets the next id
  std::vector<RawRecord> recs = index_records(data, n, fastq);
  size_t                 R    = recs.size();
  std::vector<uint64_t>  rb(R), re(R);
  for (size_t i = 0; i < R; ++i) {
    rb[i] = recs[i].begin;
    re[i] = recs[i].end;
  }
  Arena scratch;
  char *d_raw = scratch.alloc<char>(n + 1);
  if (n) MSGPU_CUDA(cudaMemcpyAsync(d_raw, data, n, cudaMemcpyHostToDevice, st));
  uint64_t *d_rb = scratch.alloc<uint64_t>(R), *d_re = scratch.alloc<uint64_t>(R);
  uint64_t *d_pb = scratch.alloc<uint64_t>(R), *d_pe = scratch.alloc<uint64_t>(R);
  if (R) {
    MSGPU_CUDA(cudaMemcpyAsync(d_rb, rb.data(), R * 8, cudaMemcpyHostToDevice, st));
    MSGPU_CUDA(cudaMemcpyAsync(d_re, re.data(), R * 8, cudaMemcpyHostToDevice, st));
  }
  uint64_t  n_tiles = (uint64_t(n) + 1 + kStripTile - 1) / kStripTile;
  uint64_t *tile    = scratch.alloc<uint64_t>(n_tiles);
  uint64_t *d_total = scratch.alloc<uint64_t>(1);
  MSGPU_LAUNCH("strip", st, strip_kernel<false><<<unsigned(n_tiles), kStripBlock, 0, st>>>(d_raw, n, d_rb, d_re, R, tile, nullptr, nullptr,
                                                                 nullptr));
  exclusive_scan<uint64_t>(scratch, st, LoadAs<uint64_t, uint64_t>{tile}, n_tiles, tile, d_total);
  uint64_t total = 0;
  MSGPU_CUDA(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, st));
  MSGPU_CUDA(cudaStreamSynchronize(st));
  char *packed = s->arena.alloc<char>(total + 16);
  MSGPU_LAUNCH("strip", st, strip_kernel<true><<<unsigned(n_tiles), kStripBlock, 0, st>>>(d_raw, n, d_rb, d_re, R, tile, packed, d_pb, d_pe));
  std::vector<uint64_t> pb(R), pe(R);
  if (R) {
    MSGPU_CUDA(cudaMemcpyAsync(pb.data(), d_pb, R * 8, cudaMemcpyDeviceToHost, st));
    MSGPU_CUDA(cudaMemcpyAsync(pe.data(), d_pe, R * 8, cudaMemcpyDeviceToHost, st));
  }
  MSGPU_CUDA(cudaStreamSynchronize(st));

  // ids: the registry numbers a name on first sight; m_idx*.emplace keeps the FIRST record of an id
  auto &hl = s->h_len[which];
  auto &hs = s->h_start[which];
  for (size_t i = 0; i < R; ++i) {
    std::string nm(data + recs[i].name_off, recs[i].name_len);
    uint32_t    id = ids.emplace(nm, uint32_t(ids.size())).first->second;
    if (id >= hl.size()) {
      hl.resize(size_t(id) + 1, 0);
      hs.resize(size_t(id) + 1, ~0ull);
    }
    if (hs[id] == ~0ull) {
      hs[id] = pb[i];
      hl[id] = pe[i] - pb[i];
    }
  }
  if (hl.size() < ids.size()) {
    hl.resize(ids.size(), 0);
    hs.resize(ids.size(), ~0ull);
  }
  uint64_t *d_start = s->arena.alloc<uint64_t>(hs.size()), *d_len = s->arena.alloc<uint64_t>(hl.size());
  if (!hs.empty()) {
    MSGPU_CUDA(cudaMemcpyAsync(d_start, hs.data(), hs.size() * 8, cudaMemcpyHostToDevice, st));
    MSGPU_CUDA(cudaMemcpyAsync(d_len, hl.data(), hl.size() * 8, cudaMemcpyHostToDevice, st));
  }
  MSGPU_CUDA(cudaStreamSynchronize(st));
  s->dev.start[which]  = d_start;
  s->dev.len[which]    = d_len;
  s->dev.packed[which] = packed;
  s->dev.n_ids[which]  = uint32_t(hs.size());
}

void fetch_on_stream(SeqStore *s, Arena &arena, cudaStream_t st, const msgpu_seq_req *d_reqs, size_t n, char *d_out,
                     uint64_t cap, uint64_t *d_out_off, uint8_t *d_status) {
  Window *win = arena.alloc<Window>(n);
  if (n) {
    MSGPU_LAUNCH("seq_window", st, seq_window_kernel<<<grid_for(n, 256), 256, 0, st>>>(s->dev, d_reqs, n, win, d_status));
  }
  uint64_t *d_total = arena.alloc<uint64_t>(1);
  exclusive_scan<uint64_t>(arena, st, WindowCount{win}, n, d_out_off, d_total);
  MSGPU_CUDA(cudaMemcpyAsync(d_out_off + n, d_total, 8, cudaMemcpyDeviceToDevice, st));
  if (n) {
    MSGPU_LAUNCH("slice_copy", st, slice_copy_kernel<<<grid_for(n * 32, 256), 256, 0, st>>>(StoreSrc{d_reqs, s->dev.packed[0], s->dev.packed[1]},
                                                             StoreRev{d_reqs}, win, d_out_off, n, cap, d_out,
                                                             d_status));
  }
}

} // namespace

SeqStore *seq_create(cudaStream_t st, const char *unitigs, size_t n_u, const char *nanopore, size_t n_n,
                     bool nanopore_is_fastq, const std::vector<std::string> &prime_reads,
                     const std::vector<std::string> &prime_unitigs) {
  auto *s = new SeqStore;
  try {
    load_file(s, st, 1, nanopore, n_n, nanopore_is_fastq, prime_reads);
    load_file(s, st, 0, unitigs, n_u, false, prime_unitigs); // the unitig file is always FASTA
  } catch (...) {
    delete s;
    throw;
  }
  return s;
}

void seq_destroy(SeqStore *s) { delete s; }

bool seq_lookup(const SeqStore *s, uint32_t which, const char *name, uint32_t *id) {
  auto it = s->ids[which].find(name);
  if (it == s->ids[which].end()) return false;
  *id = it->second;
  return true;
}

uint64_t seq_batch_bound(const SeqStore *s, const msgpu_seq_req *reqs, size_t n) {
  uint64_t total = 0;
  for (size_t i = 0; i < n; ++i) {
    const auto &q = reqs[i];
    if (q.which > 1 || q.id >= s->h_len[q.which].size() || s->h_start[q.which][q.id] == ~0ull) continue;
    uint64_t st, c;
    if (slice_window(s->h_len[q.which][q.id], q.left, q.right + 1, &st, &c)) total += c;
  }
  return total;
}

void seq_fetch_device(SeqStore *s, cudaStream_t st, const msgpu_seq_req *d_reqs, size_t n, char *d_out, uint64_t cap,
                      uint64_t *d_out_off, uint8_t *d_status) {
  // scratch for the windows comes from a pooled workspace so concurrent callers never share it
  Workspace *ws = nullptr;
  {
    std::lock_guard<std::mutex> lk(s->mu);
    if (!s->free_ws.empty()) {
      ws = s->free_ws.back();
      s->free_ws.pop_back();
    }
  }
  if (!ws) ws = new Workspace;
  try {
    ws->arena.reset();
    fetch_on_stream(s, ws->arena, st, d_reqs, n, d_out, cap, d_out_off, d_status);
    MSGPU_CUDA(cudaStreamSynchronize(st)); // the workspace goes back to the pool only when the kernels are done
  } catch (...) {
    delete ws;
    throw;
  }
  std::lock_guard<std::mutex> lk(s->mu);
  s->free_ws.push_back(ws);
}

void seq_fetch_host(SeqStore *s, const msgpu_seq_req *reqs, size_t n, char *out, uint64_t cap, uint64_t *out_off,
                    uint8_t *status) {
  uint64_t need = seq_batch_bound(s, reqs, n);
  if (need > cap) throw BuildError{MSGPU_ERR_INVALID_ARG, "output buffer too small (see msgpu_seq_batch_bound)"};
  Workspace *ws = nullptr;
  {
    std::lock_guard<std::mutex> lk(s->mu);
    if (!s->free_ws.empty()) {
      ws = s->free_ws.back();
      s->free_ws.pop_back();
    }
  }
  if (!ws) ws = new Workspace;
  try {
    if (!ws->stream) MSGPU_CUDA(cudaStreamCreateWithFlags(&ws->stream, cudaStreamNonBlocking));
    cudaStream_t st = ws->stream;
    ws->arena.reset();
    size_t req_bytes = n * sizeof(msgpu_seq_req), off_bytes = (n + 1) * 8, st_bytes = n;
    size_t o_req = 0, o_off = (o_req + req_bytes + 63) & ~size_t(63), o_st = (o_off + off_bytes + 63) & ~size_t(63),
           o_out = (o_st + st_bytes + 63) & ~size_t(63);
    char  *h = static_cast<char *>(ws->host(o_out + need + 64));
    std::memcpy(h + o_req, reqs, req_bytes);
    auto *d_reqs = ws->arena.alloc<msgpu_seq_req>(n);
    auto *d_off  = ws->arena.alloc<uint64_t>(n + 1);
    auto *d_st   = ws->arena.alloc<uint8_t>(n);
    auto *d_out  = ws->arena.alloc<char>(need + 16);
    if (n) MSGPU_CUDA(cudaMemcpyAsync(d_reqs, h + o_req, req_bytes, cudaMemcpyHostToDevice, st));
    fetch_on_stream(s, ws->arena, st, d_reqs, n, d_out, need, d_off, d_st);
    MSGPU_CUDA(cudaMemcpyAsync(h + o_off, d_off, off_bytes, cudaMemcpyDeviceToHost, st));
    if (n) MSGPU_CUDA(cudaMemcpyAsync(h + o_st, d_st, st_bytes, cudaMemcpyDeviceToHost, st));
    if (need) MSGPU_CUDA(cudaMemcpyAsync(h + o_out, d_out, need, cudaMemcpyDeviceToHost, st));
    MSGPU_CUDA(cudaStreamSynchronize(st));
    std::memcpy(out_off, h + o_off, off_bytes);
    std::memcpy(status, h + o_st, st_bytes);
    std::memcpy(out, h + o_out, need);
  } catch (...) {
    delete ws;
    throw;
  }
  std::lock_guard<std::mutex> lk(s->mu);
  s->free_ws.push_back(ws);
}

void slice_revcomp_device(cudaStream_t st, const char *d_text, const uint64_t *d_text_off, const int32_t *d_left,
                          const int32_t *d_right, const uint8_t *d_revcomp, size_t n, char *d_out, uint64_t cap,
                          uint64_t *d_out_off, uint8_t *d_status) {
  Arena   arena;
  Window *win = arena.alloc<Window>(n);
  if (n) {
    MSGPU_LAUNCH("text_window", st, text_window_kernel<<<grid_for(n, 256), 256, 0, st>>>(d_text_off, d_left, d_right, n, win, d_status));
  }
  uint64_t *d_total = arena.alloc<uint64_t>(1);
  exclusive_scan<uint64_t>(arena, st, WindowCount{win}, n, d_out_off, d_total);
  MSGPU_CUDA(cudaMemcpyAsync(d_out_off + n, d_total, 8, cudaMemcpyDeviceToDevice, st));
  if (n) {
    MSGPU_LAUNCH("slice_copy", st, slice_copy_kernel<<<grid_for(n * 32, 256), 256, 0, st>>>(TextSrc{d_text}, TextRev{d_revcomp}, win, d_out_off, n,
                                                             cap, d_out, d_status));
  }
  MSGPU_CUDA(cudaStreamSynchronize(st)); // `arena` is released on return
}

} // namespace msgpu
