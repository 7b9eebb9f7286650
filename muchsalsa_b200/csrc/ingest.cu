// ingest.cu — PAF text resident in HBM -> packed structure-of-arrays match records (kernels K1-K5 of SURVEY.md §2).
//
// Replaces, for the reference run with one worker thread (file:line relative to the reference tree):
//   K1 line index            BlastFileAccessor::_buildIndex   BlastFileAccessor.cpp:77-91 (+ readline, IO.cpp:54-97)
//   K1 tokenise              BlastFileReader::parseLine       BlastFileReader.cpp:86-104
//   K2 filter + derive       BlastFileReader::parseLine       BlastFileReader.cpp:101-125
//   K3 id interning, vertex  Registry::operator[]             Registry.cpp:36-45 ; Graph::_addVertex Graph.cpp:141-149
//   K4 keep-min-line dedup   MatchMap::addVertexMatch         MatchMap.cpp:53-81
//   K5 group by unitig       m_scaffolds + sort by line       MatchMap.cpp:77-79,178-183
//
// Everything order-dependent in the reference is made order-free here:
//   "first request gets the next id"  ->  atomicMin of the line index per name, then id = rank of that line among the
//                                         first-appearance lines (one prefix sum over lines)
//   "first Vertex of an id wins"      ->  the vertex takes its fields from that same first-appearance line
//   "smaller line replaces"           ->  atomicMin of the line index per (unitig, read) key
// Names are hashed (64 bit) for the open-addressing tables, and every line then compares its name bytes with the
// representative of its slot; a mismatch (a 64-bit collision) restarts interning with another hash seed.
#include <algorithm>
#include <cstdlib>

#include "../../include/msgpu.h"
#include "internal.h"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace msgpu {

// thresholds: BlastFileReader.cpp:48-50
constexpr int      kMinimumMatches = 400;
constexpr int      kThLength       = 500;
constexpr uint64_t kThMatches      = 500;

constexpr uint8_t kFAccepted = 1, kFDir = 2, kFPrim = 4, kFFirstRead = 8, kFFirstUnitig = 16, kFSurvivor = 32;

// Everything the host needs to know about an ingest, kept in device memory and read back ONCE.  All buffers are
// sized from byte-count bounds (a line with 10 fields has >= 20 bytes), so no count has to reach the host earlier.
struct IngestStatus {
  unsigned long long first_bad_line; // smallest line index the reference would have thrown on; ~0 = none
  unsigned long long n_lines, n_parse;
  unsigned long long read_name_bytes, unitig_name_bytes;
  unsigned           n_accepted, collision, unsorted, overflow; // overflow: more lines than n_bytes / 20 (all invalid)
  unsigned           n_reads, n_unitigs, n_matches, pad;
};

// ---------------------------------------------------------------------------------------------------------------------
// K1a: line index.  A thread owns 32 consecutive bytes (two aligned 128-bit loads), a block a tile of 8 KB.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kNlBlock  = 256;
constexpr int kNlPieces = 8;                        // a thread owns 8 pieces of 32 bytes, 8 KB apart (coalesced)
constexpr int kNlPiece  = kNlBlock * 32;            // bytes one sweep of the block covers
constexpr int kNlTile   = kNlPiece * kNlPieces;     // 64 KB per block

__device__ __forceinline__ unsigned nl_mask4(unsigned w) {
  unsigned y = (__vcmpeq4(w, 0x0a0a0a0au) & 0x80808080u) >> 7; // bit 0/8/16/24 set where the byte is '\n'
  return ((y * 0x00204081u) >> 21) & 0xfu;
}
__device__ __forceinline__ unsigned nl_mask16(uint4 v) {
  return nl_mask4(v.x) | (nl_mask4(v.y) << 4) | (nl_mask4(v.z) << 8) | (nl_mask4(v.w) << 12);
}

// newline bit mask of the 32 virtual bytes [vb, vb+32) of the buffer that starts `mis` bytes after a 16-byte boundary
__device__ __forceinline__ unsigned nl_mask32(const uint4 *aligned, uint64_t vb, uint64_t mis, uint64_t nv) {
  if (vb >= nv) return 0;
  unsigned m = nl_mask16(ld_stream_u4(aligned + (vb >> 4)));
  if (vb + 16 < nv) m |= nl_mask16(ld_stream_u4(aligned + (vb >> 4) + 1)) << 16;
  if (vb < mis) m &= ~0u << (mis - vb);                        // bytes before the buffer
  if (vb + 32 > nv) m &= (nv - vb >= 32) ? ~0u : ((1u << (nv - vb)) - 1u); // bytes after it
  return m;
}

__global__ void __launch_bounds__(kNlBlock)
    newline_count_kernel(const uint4 *__restrict__ aligned, uint64_t mis, uint64_t nv, uint32_t *__restrict__ tile_cnt) {
  pdl_sync();
  __shared__ uint32_t sm[33];
  const uint64_t      base = uint64_t(blockIdx.x) * kNlTile + uint64_t(threadIdx.x) * 32;
  uint32_t            c    = 0;
#pragma unroll
  for (int i = 0; i < kNlPieces; ++i) c += __popc(nl_mask32(aligned, base + uint64_t(i) * kNlPiece, mis, nv));
  uint32_t total;
  block_exclusive_scan<uint32_t>(c, sm, &total);
  if (threadIdx.x == 0) tile_cnt[blockIdx.x] = total;
}

// line_start[k+1] = offset of the byte after the k-th newline (k < cap_lines; beyond that the file is invalid anyway)
__global__ void __launch_bounds__(kNlBlock)
    newline_write_kernel(const uint4 *__restrict__ aligned, uint64_t mis, uint64_t nv,
                         const uint32_t *__restrict__ tile_off, uint64_t cap_lines, uint64_t *__restrict__ line_start) {
  pdl_sync();
  __shared__ uint32_t sm[33];
  const uint64_t      base  = uint64_t(blockIdx.x) * kNlTile + uint64_t(threadIdx.x) * 32;
  uint32_t            carry = tile_off[blockIdx.x];
  for (int i = 0; i < kNlPieces; ++i) {
    const uint64_t vb = base + uint64_t(i) * kNlPiece;
    unsigned       m  = nl_mask32(aligned, vb, mis, nv);
    uint32_t       total;
    uint32_t       rank = carry + block_exclusive_scan<uint32_t>(__popc(m), sm, &total);
    carry += total;
    while (m) {
      int b = __ffs(m) - 1;
      m &= m - 1;
      if (uint64_t(rank) + 1 <= cap_lines) line_start[uint64_t(rank) + 1] = vb + b - mis + 1;
      ++rank;
    }
  }
}

__global__ void line_index_finish_kernel(const char *__restrict__ paf, uint64_t n, const uint32_t *n_newlines,
                                         uint64_t cap_lines, int drop_last_line, uint64_t *line_start,
                                         IngestStatus *status) {
  pdl_sync();
  // a trailing run of bytes without '\n' is one more line (IO.cpp:92-95); give it a virtual terminator so that
  // length(line i) = line_start[i+1] - line_start[i] - 1 holds for every line
  uint64_t nn   = *n_newlines;
  bool     open = n > 0 && paf[n - 1] != '\n';
  uint64_t nl   = nn + (open ? 1 : 0);
  line_start[0] = 0;
  if (nl > cap_lines) { // more lines than a file of n bytes can hold with >= 20 bytes each: some parsed line is invalid
    status->overflow = 1;
    nl               = cap_lines;
  } else if (open) {
    line_start[nn + 1] = n + 1;
  }
  status->n_lines = nn + (open ? 1 : 0);
  // BlastFileReader::read(): lines 0 .. lineCount-2 (BlastFileReader.cpp:76).  A shard of a PAF that is not the last
  // one keeps its last line: only the last line of the whole file is never parsed.
  status->n_parse = drop_last_line ? (nl ? nl - 1 : 0) : nl;
}

// ---------------------------------------------------------------------------------------------------------------------
// K1b + K2: tokenise, std::stoi the numeric fields, accept filter.  One thread per line; a warp covers ~2-3 KB of
// contiguous text, which L1 serves after the first touch of each 128-byte line.
// ---------------------------------------------------------------------------------------------------------------------
struct LineFields { // per-line scratch, structure-of-arrays, indexed by line
  int32_t  *c2, *ir1, *c6, *nr0, *nr1, *c9;
  uint32_t *rname_rel; // offset of the read name (field 5) from the line start
  uint32_t *rname_len, *uname_len;
  uint64_t *rhash, *uhash;
  uint8_t  *flags;
};

__device__ __forceinline__ bool dev_isspace(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// std::stoi on [s, s+n): false if it would throw (no digits / out of int range).  32-bit arithmetic: up to 9
// significant digits cannot overflow, the 10th is checked once, more are out of range.
__device__ __forceinline__ bool dev_stoi(const char *s, uint32_t n, int *out) {
  uint32_t i = 0;
  if (n && (unsigned char)s[0] <= ' ')  // leading white space is rare: one compare on the common path
    while (i < n && dev_isspace(s[i])) ++i;
  bool neg = false;
  if (i < n && (s[i] == '+' || s[i] == '-')) {
    neg = s[i] == '-';
    ++i;
  }
  if (i >= n) return false;
  uint32_t d = uint32_t(s[i]) - uint32_t('0');
  if (d > 9u) return false;
  // common path: at most 9 digits cannot overflow and need no bookkeeping
  {
    uint32_t v = 0, j = i, nd = 0;
    do {
      v = v * 10u + d;
      ++j;
      ++nd;
      if (j >= n) break;
      d = uint32_t(s[j]) - uint32_t('0');
    } while (d <= 9u && nd < 9u);
    if (j >= n || d > 9u) { // the number ended within 9 digits
      *out = neg ? int(0u - v) : int(v);
      return true;
    }
    d = uint32_t(s[i]) - uint32_t('0'); // 10 digits or more: the exact walk below, from the first digit
  }
  uint32_t v = 0, sig = 0; // sig: digits since the first non-zero one
  do {
    if (sig == 9) { // the 10th significant digit decides
      if (v > 214748364u || (v == 214748364u && d > (neg ? 8u : 7u))) return false;
    } else if (sig > 9) {
      return false;
    }
    v   = v * 10u + d;
    sig = (sig | v) ? sig + 1 : 0;
    ++i;
    if (i >= n) break;
    d = uint32_t(s[i]) - uint32_t('0');
  } while (d <= 9u);
  *out = neg ? int(0u - v) : int(v);
  return true;
}

// 64-bit name hash from two independent 32-bit multiplicative streams (a 64-bit multiply per byte costs four times
// as much on this machine); 0 is the empty-slot marker of the tables, so the low bit is forced
__device__ __forceinline__ uint64_t name_hash(const char *s, uint32_t n, uint64_t seed) {
  if (seed == 0) return 1ull; // test hook (MSGPU_TEST_FORCE_COLLISION): every name collides, the verify pass must notice
  uint32_t h1 = 2166136261u ^ uint32_t(seed), h2 = 0x9e3779b9u ^ uint32_t(seed >> 32);
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t c = (unsigned char)s[i];
    h1               = (h1 ^ c) * 16777619u;
    h2               = (h2 + c) * 0x85ebca6bu;
    h2 ^= h2 >> 15;
  }
  return mix64((uint64_t(h1) << 32 | h2) ^ (uint64_t(n) << 20)) | 1ull;
}

// the tokenising + std::stoi + filter of one line whose text starts at `s` (shared or global memory)
__device__ __forceinline__ bool parse_one_line(const char *s, uint64_t len, uint64_t li, uint64_t seed, LineFields &f,
                                               bool *bad_out) {
  // std::getline(iss, token, '\t'): a field exists iff at least one character is left
  // (the loop over the ten fields is unrolled so that fo / fl stay in registers; lines of 4 GiB and more do not exist
  //  on this path: the whole buffer is below 80 GB and a line is counted in 32 bits)
  uint32_t       fo[10], fl[10];
  int            ntok = 0;
  const uint32_t n    = len < 0xfffffff0ull ? uint32_t(len) : 0xfffffff0u;
  uint32_t       p    = 0;
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    fo[k] = 0;
    fl[k] = 0;
    if (p < n) {
      uint32_t q = p;
      while (q < n && s[q] != '\t') ++q;
      fo[k] = p;
      fl[k] = q - p;
      ntok  = k + 1;
      p     = q + 1;
    } else {
      p = 0xffffffffu; // no further field
    }
  }
  uint8_t flags = 0;
  bool    bad   = ntok < 10; // BlastFileReader.cpp:97 ("Invalid BLAST file.")
  bool    acc   = false;
  int     c2 = 0, c3 = 0, c9 = 0, c6 = 0, c7 = 0, c8 = 0;
  if (!bad)
    bad = !(dev_stoi(s + fo[2], fl[2], &c2) && dev_stoi(s + fo[3], fl[3], &c3) && dev_stoi(s + fo[9], fl[9], &c9) &&
            dev_stoi(s + fo[6], fl[6], &c6));
  if (!bad) {
    int      ir1     = c3 - 1;
    int      ilen    = ir1 - c2 + 1;
    uint64_t matches = uint64_t((long long)c9); // static_cast<std::size_t>(int), BlastFileReader.cpp:102
    bool     add     = matches >= uint64_t(kMinimumMatches) && ilen >= kMinimumMatches;
    if (add) {
      bad = !(dev_stoi(s + fo[7], fl[7], &c7) && dev_stoi(s + fo[8], fl[8], &c8));
      if (!bad) {
        flags = kFAccepted;
        if (fl[4] == 1 && s[fo[4]] == '+') flags |= kFDir;
        if (ilen >= kThLength && matches >= kThMatches) flags |= kFPrim;
        f.c2[li]        = c2;
        f.ir1[li]       = ir1;
        f.c6[li]        = c6;
        f.nr0[li]       = c7;
        f.nr1[li]       = c8 - 1;
        f.c9[li]        = c9;
        f.rname_rel[li] = fo[5];
        f.rname_len[li] = fl[5];
        f.uname_len[li] = fl[0];
        f.rhash[li]     = name_hash(s + fo[5], fl[5], seed);
        f.uhash[li]     = name_hash(s + fo[0], fl[0], seed ^ 0x9e3779b97f4a7c15ull);
        acc             = true;
      }
    }
  }
  *bad_out    = bad;
  f.flags[li] = flags;
  return acc;
}

// ---------------------------------------------------------------------------------------------------------------------
// The pipeline's scrubber reads the same PAF with other rules (pipeline/scrubber_bfs.py:57-100):
//   line.rstrip().split("\t")            trailing white space is not part of the line; a line with ONE field is skipped
//   linetemp[0..8], int() of 1 2 3 6 7 8  fewer than nine fields -> IndexError, a non-integer -> ValueError (both: invalid)
//   e_1 - s_1 >= 500                      the only accept rule (unitig span; no match count, no last-line drop)
// Ranges are kept RAW here ((s_2, e_2) is what the script stores, scrubber_bfs.py:100), and int() is Python's: optional
// white space on both sides, an optional sign, digits — nothing may follow (std::stoi stops at the first non-digit).
// Python's digit-group underscores ("1_000") and values beyond 32 bits are reported as invalid.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool py_isspace(char c) { return c == ' ' || (c >= '\t' && c <= '\r') || (c >= 0x1c && c <= 0x1f); }

__device__ __forceinline__ bool dev_pyint(const char *s, uint32_t n, int *out) {
  uint32_t i = 0;
  while (i < n && py_isspace(s[i])) ++i;
  while (n > i && py_isspace(s[n - 1])) --n;
  bool neg = false;
  if (i < n && (s[i] == '+' || s[i] == '-')) {
    neg = s[i] == '-';
    ++i;
  }
  if (i >= n) return false;
  unsigned long long v = 0;
  for (; i < n; ++i) {
    const uint32_t d = uint32_t(s[i]) - uint32_t('0');
    if (d > 9u) return false;
    v = v * 10ull + d;
    if (v > 2147483648ull) return false; // beyond int: not representable in the tables
  }
  if (!neg && v > 2147483647ull) return false;
  *out = neg ? int(0ll - (long long)v) : int(v);
  return true;
}

__device__ __forceinline__ bool parse_one_line_scrubber(const char *s, uint64_t len, uint64_t li, uint64_t seed,
                                                        LineFields &f, bool *bad_out) {
  uint32_t n = len < 0xfffffff0ull ? uint32_t(len) : 0xfffffff0u;
  while (n && py_isspace(s[n - 1])) --n; // str.rstrip()
  uint32_t fo[9], fl[9];
  int      ntok = 0;
  uint32_t p    = 0;
  bool     more = n > 0; // split("\t"): one field per tab, plus one (an empty line is one empty field: skipped below)
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    fo[k] = 0;
    fl[k] = 0;
    if (more) {
      uint32_t q = p;
      while (q < n && s[q] != '\t') ++q;
      fo[k] = p;
      fl[k] = q - p;
      ntok  = k + 1;
      more  = q < n; // a tab follows (the stripped line cannot END in a tab, so another field does)
      p     = q + 1;
    }
  }
  uint8_t flags = 0;
  bool    bad = false, acc = false;
  if (ntok >= 2) { // (one field: the line is skipped, scrubber_bfs.py:76-77)
    int len_1 = 0, len_2 = 0, s_1 = 0, s_2 = 0, e_1 = 0, e_2 = 0;
    bad = ntok < 9 || !(dev_pyint(s + fo[1], fl[1], &len_1) && dev_pyint(s + fo[6], fl[6], &len_2) &&
                        dev_pyint(s + fo[2], fl[2], &s_1) && dev_pyint(s + fo[7], fl[7], &s_2) &&
                        dev_pyint(s + fo[3], fl[3], &e_1) && dev_pyint(s + fo[8], fl[8], &e_2));
    if (!bad && (long long)e_1 - (long long)s_1 >= 500ll) { // threshold_length, scrubber_bfs.py:50,85
      flags = kFAccepted;
      if (fl[4] == 1 && s[fo[4]] == '+') flags |= kFDir;
      f.c2[li]        = s_1;
      f.ir1[li]       = e_1;
      f.c6[li]        = len_2;
      f.nr0[li]       = s_2;
      f.nr1[li]       = e_2;
      f.c9[li]        = len_1;
      f.rname_rel[li] = fo[5];
      f.rname_len[li] = fl[5];
      f.uname_len[li] = fl[0];
      f.rhash[li]     = name_hash(s + fo[5], fl[5], seed);
      f.uhash[li]     = name_hash(s + fo[0], fl[0], seed ^ 0x9e3779b97f4a7c15ull);
      acc             = true;
    }
  }
  *bad_out    = bad;
  f.flags[li] = flags;
  return acc;
}

// The pipeline's unitig filter reads the PAF with yet other rules (pipeline/unitig_filter.py:42-50): every line counts
// (no accept rule, no skipped lines), fields 0, 1, 2, 3, 5, 6 must exist (IndexError otherwise) and 1, 2, 3, 6 must be
// integers for Python's int().  Kept per line: illuminaStart = c2, illuminaEnd = c3 - 1, the unitig length (field 1).
__device__ __forceinline__ bool parse_one_line_coverage(const char *s, uint64_t len, uint64_t li, uint64_t seed,
                                                        LineFields &f, bool *bad_out) {
  uint32_t n = len < 0xfffffff0ull ? uint32_t(len) : 0xfffffff0u;
  while (n && py_isspace(s[n - 1])) --n; // str.rstrip()
  uint32_t fo[7], fl[7];
  int      ntok = 0;
  uint32_t p    = 0;
  bool     more = n > 0;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    fo[k] = 0;
    fl[k] = 0;
    if (more) {
      uint32_t q = p;
      while (q < n && s[q] != '\t') ++q;
      fo[k] = p;
      fl[k] = q - p;
      ntok  = k + 1;
      more  = q < n;
      p     = q + 1;
    }
  }
  int  ulen = 0, c2 = 0, c3 = 0, rlen = 0;
  bool bad  = ntok < 7 || !(dev_pyint(s + fo[1], fl[1], &ulen) && dev_pyint(s + fo[2], fl[2], &c2) &&
                           dev_pyint(s + fo[3], fl[3], &c3) && dev_pyint(s + fo[6], fl[6], &rlen));
  // positions index a list of `ulen` counters: a start below 0 or an end beyond the unitig raises IndexError (or,
  // negative, silently wraps around in Python: reported as invalid here); an empty unitig has no maximum (ValueError)
  if (!bad) bad = ulen <= 0 || c2 < 0 || c3 - 1 >= ulen;
  uint8_t flags = 0;
  if (!bad) {
    flags           = kFAccepted;
    f.c2[li]        = c2;
    f.ir1[li]       = c3 - 1;
    f.c6[li]        = rlen;
    f.nr0[li]       = 0;
    f.nr1[li]       = 0;
    f.c9[li]        = ulen;
    f.rname_rel[li] = fo[5];
    f.rname_len[li] = fl[5];
    f.uname_len[li] = fl[0];
    f.rhash[li]     = name_hash(s + fo[5], fl[5], seed);
    f.uhash[li]     = name_hash(s + fo[0], fl[0], seed ^ 0x9e3779b97f4a7c15ull);
  }
  *bad_out    = bad;
  f.flags[li] = flags;
  return !bad;
}

// One thread per line, a block per run of 256 consecutive lines.  A thread walks its line byte by byte; straight from
// global memory that costs one L1 wavefront per lane and byte (32 lanes, 32 different cache lines).  So the block
// first copies the text of its 256 lines — one contiguous piece of the buffer — into shared memory with coalesced
// 128-bit loads, when it fits (it does unless lines carry long cg:Z tags; then the lines are read in place).
constexpr int      kParseBlock = 256;
constexpr uint32_t kParseStage = 40 * 1024;

template <int kMode> // 0: BlastFileReader's rules, 1: the scrubber's, 2: the unitig filter's
__global__ void __launch_bounds__(kParseBlock)
    parse_lines_kernel(const char *__restrict__ paf, uint64_t n_bytes, const uint64_t *__restrict__ line_start,
                       uint64_t seed, LineFields f, IngestStatus *status) {
  pdl_sync();
  __shared__ __align__(16) char stage[kParseStage + 32];
  const uint64_t n_parse = status->n_parse;
  unsigned       n_acc   = 0;
  for (uint64_t l0 = uint64_t(blockIdx.x) * kParseBlock; l0 < n_parse; l0 += uint64_t(gridDim.x) * kParseBlock) {
    const uint64_t l1 = l0 + kParseBlock < n_parse ? l0 + kParseBlock : n_parse;
    const uint64_t b0 = line_start[l0];
    uint64_t       b1 = line_start[l1]; // one past the terminator of the last line of the run
    if (b1 > n_bytes) b1 = n_bytes;     // the virtual terminator of an unterminated last line
    const uint64_t a0     = (reinterpret_cast<uintptr_t>(paf) + b0) & 15u; // distance to the 16-byte boundary below
    const bool     staged = b1 - b0 + a0 <= kParseStage;
    __syncthreads(); // the previous run is done with the staging buffer
    if (staged) {
      const uint4   *src = reinterpret_cast<const uint4 *>(paf + b0 - a0);
      const uint64_t nv  = (b1 - b0 + a0 + 15) >> 4;
      for (uint64_t v = threadIdx.x; v < nv; v += kParseBlock) reinterpret_cast<uint4 *>(stage)[v] = ld_stream_u4(src + v);
    }
    __syncthreads();
    const uint64_t li = l0 + threadIdx.x;
    if (li < l1) {
      const uint64_t ls  = line_start[li];
      const uint64_t len = line_start[li + 1] - ls - 1; // getLine() drops the last character of the line
      const char    *s   = staged ? stage + a0 + (ls - b0) : paf + ls;
      bool           bad;
      const bool acc = kMode == 1   ? parse_one_line_scrubber(s, len, li, seed, f, &bad)
                       : kMode == 2 ? parse_one_line_coverage(s, len, li, seed, f, &bad)
                                    : parse_one_line(s, len, li, seed, f, &bad);
      if (acc) ++n_acc;
      if (bad) atomicMin(&status->first_bad_line, (unsigned long long)li);
    }
  }
  n_acc = warp_reduce_sum(n_acc);
  if (lane_id() == 0 && n_acc) atomicAdd(&status->n_accepted, n_acc);
}

// ---------------------------------------------------------------------------------------------------------------------
// K3: interning.  Open addressing on the 64-bit name hash; the slot keeps the smallest accepted line naming it.
// ---------------------------------------------------------------------------------------------------------------------
struct InternTable {
  unsigned long long *keys;    // 0 = empty
  uint32_t           *minline; // 0xffffffff initially
  uint32_t           *id;
  uint64_t            mask;
};

// (both registries in one launch: blockIdx.y = 0 reads, 1 unitigs)
__global__ void __launch_bounds__(256)
    intern_insert_kernel(const uint8_t *__restrict__ flags, const uint64_t *__restrict__ rhash,
                         const uint64_t *__restrict__ uhash, const IngestStatus *__restrict__ status, InternTable rt,
                         InternTable ut, uint32_t *__restrict__ rslot, uint32_t *__restrict__ uslot) {
  pdl_sync();
  const uint64_t *__restrict__ hash = blockIdx.y ? uhash : rhash;
  const InternTable t               = blockIdx.y ? ut : rt;
  uint32_t *__restrict__ slot_of_line = blockIdx.y ? uslot : rslot;
  const uint64_t n_parse = status->n_parse;
  for (uint64_t li = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; li < n_parse;
       li += uint64_t(gridDim.x) * blockDim.x) {
    if (!(flags[li] & kFAccepted)) continue;
    unsigned long long h = hash[li];
    uint64_t           p = (h >> 1) & t.mask;
    for (;;) {
      unsigned long long prev = atomicCAS(&t.keys[p], 0ull, h);
      if (prev == 0ull || prev == h) break;
      p = (p + 1) & t.mask;
    }
    atomicMin(&t.minline[p], uint32_t(li));
    slot_of_line[li] = uint32_t(p);
  }
}

// every accepted line checks its name against the representative line of its slot and learns whether it is the
// first appearance of the name
__global__ void __launch_bounds__(256)
    intern_verify_kernel(const char *__restrict__ paf, const uint64_t *__restrict__ line_start, LineFields f,
                         InternTable rt, const uint32_t *__restrict__ rslot, InternTable ut,
                         const uint32_t *__restrict__ uslot, IngestStatus *status) {
  pdl_sync();
  const uint64_t n_parse = status->n_parse;
  for (uint64_t li = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; li < n_parse;
       li += uint64_t(gridDim.x) * blockDim.x) {
    uint8_t fl = f.flags[li];
    if (!(fl & kFAccepted)) continue;
    bool     clash = false;
    uint32_t rrep  = rt.minline[rslot[li]];
    if (rrep == uint32_t(li)) {
      fl |= kFFirstRead;
    } else {
      uint32_t    n = f.rname_len[li];
      const char *a = paf + line_start[li] + f.rname_rel[li];
      const char *b = paf + line_start[rrep] + f.rname_rel[rrep];
      clash         = f.rname_len[rrep] != n;
      for (uint32_t i = 0; i < n && !clash; ++i) clash = a[i] != b[i];
    }
    uint32_t urep = ut.minline[uslot[li]];
    if (urep == uint32_t(li)) {
      fl |= kFFirstUnitig;
    } else if (!clash) {
      uint32_t    n = f.uname_len[li];
      const char *a = paf + line_start[li];
      const char *b = paf + line_start[urep];
      clash         = f.uname_len[urep] != n;
      for (uint32_t i = 0; i < n && !clash; ++i) clash = a[i] != b[i];
    }
    if (clash) status->collision = 1;
    f.flags[li] = fl;
  }
}

// the first-appearance lines publish the id of their name (= their rank among first-appearance lines) and fill the
// vertex / name tables
__global__ void __launch_bounds__(256)
    intern_assign_kernel(const uint64_t *__restrict__ line_start, const IngestStatus *__restrict__ status,
                         LineFields f, InternTable rt,
                         const uint32_t *__restrict__ rslot, const uint32_t *__restrict__ rrank, InternTable ut,
                         const uint32_t *__restrict__ uslot, const uint32_t *__restrict__ urank,
                         int32_t *__restrict__ read_nlen, uint32_t *__restrict__ read_line,
                         uint64_t *__restrict__ read_name_src, uint32_t *__restrict__ read_name_len,
                         uint64_t *__restrict__ unitig_name_src, uint32_t *__restrict__ unitig_name_len) {
  pdl_sync();
  const uint64_t n_parse = status->n_parse;
  for (uint64_t li = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; li < n_parse;
       li += uint64_t(gridDim.x) * blockDim.x) {
    uint8_t fl = f.flags[li];
    if (fl & kFFirstRead) {
      uint32_t id        = rrank[li];
      rt.id[rslot[li]]   = id;
      read_nlen[id]      = f.c6[li];
      read_line[id]      = uint32_t(li);
      read_name_src[id]  = line_start[li] + f.rname_rel[li];
      read_name_len[id]  = f.rname_len[li];
    }
    if (fl & kFFirstUnitig) {
      uint32_t id         = urank[li];
      ut.id[uslot[li]]    = id;
      unitig_name_src[id] = line_start[li];
      unitig_name_len[id] = f.uname_len[li];
    }
  }
}

__global__ void __launch_bounds__(256)
    gather_names_kernel(const char *__restrict__ paf, const uint64_t *__restrict__ src,
                        const uint64_t *__restrict__ dst_off, const unsigned *__restrict__ d_n,
                        char *__restrict__ out) {
  pdl_sync();
  // one warp per name
  const uint64_t n = *d_n;
  uint64_t w  = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  uint64_t nw = (uint64_t(gridDim.x) * blockDim.x) >> 5;
  for (; w < n; w += nw) {
    uint64_t b = dst_off[w], e = dst_off[w + 1];
    for (uint64_t i = lane_id(); i < e - b; i += 32) out[b + i] = paf[src[w] + i];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// K4: per (unitig, read) keep the accepted line with the smallest index (MatchMap.cpp:64-73).  The key is exact
// (two dense ids), so this table needs no verification.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    dedup_insert_kernel(const IngestStatus *__restrict__ status, LineFields f, InternTable rt,
                        const uint32_t *__restrict__ rslot, InternTable ut, const uint32_t *__restrict__ uslot,
                        InternTable dt, uint32_t *__restrict__ dslot, uint32_t *__restrict__ rid_of_line,
                        uint32_t *__restrict__ uid_of_line) {
  pdl_sync();
  const uint64_t n_parse = status->n_parse;
  for (uint64_t li = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; li < n_parse;
       li += uint64_t(gridDim.x) * blockDim.x) {
    if (!(f.flags[li] & kFAccepted)) continue;
    uint32_t           rid = rt.id[rslot[li]], uid = ut.id[uslot[li]];
    unsigned long long key = ((unsigned long long)uid << 32 | rid) + 1ull;
    uint64_t           p   = mix64(key) & dt.mask;
    for (;;) {
      unsigned long long prev = atomicCAS(&dt.keys[p], 0ull, key);
      if (prev == 0ull || prev == key) break;
      p = (p + 1) & dt.mask;
    }
    atomicMin(&dt.minline[p], uint32_t(li));
    dslot[li]       = uint32_t(p);
    rid_of_line[li] = rid;
    uid_of_line[li] = uid;
  }
}

__global__ void __launch_bounds__(256)
    dedup_flag_kernel(const IngestStatus *__restrict__ status, uint8_t *__restrict__ flags, InternTable dt,
                      const uint32_t *__restrict__ dslot) {
  pdl_sync();
  const uint64_t n_parse = status->n_parse;
  for (uint64_t li = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; li < n_parse;
       li += uint64_t(gridDim.x) * blockDim.x) {
    uint8_t fl = flags[li];
    if ((fl & kFAccepted) && dt.minline[dslot[li]] == uint32_t(li)) flags[li] = fl | kFSurvivor;
  }
}

// the unitig filter keeps EVERY line (its per-read dedup is per block, done after the blocks are known)
__global__ void __launch_bounds__(256)
    all_survive_kernel(const IngestStatus *__restrict__ status, uint8_t *__restrict__ flags, InternTable rt,
                       const uint32_t *__restrict__ rslot, InternTable ut, const uint32_t *__restrict__ uslot,
                       uint32_t *__restrict__ rid_of_line, uint32_t *__restrict__ uid_of_line) {
  pdl_sync();
  const uint64_t n_parse = status->n_parse;
  for (uint64_t li = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; li < n_parse;
       li += uint64_t(gridDim.x) * blockDim.x) {
    const uint8_t fl = flags[li];
    if (!(fl & kFAccepted)) continue;
    rid_of_line[li] = rt.id[rslot[li]];
    uid_of_line[li] = ut.id[uslot[li]];
    flags[li]       = fl | kFSurvivor;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// K5: compaction of the surviving lines into the packed match records (line order), rRatio included
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    compact_matches_kernel(const IngestStatus *__restrict__ status, LineFields f, const uint32_t *__restrict__ pos,
                           const uint32_t *__restrict__ rid_of_line, const uint32_t *__restrict__ uid_of_line,
                           MatchSoA m) {
  pdl_sync();
  const uint64_t n_parse = status->n_parse;
  for (uint64_t li = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; li < n_parse;
       li += uint64_t(gridDim.x) * blockDim.x) {
    uint8_t fl = f.flags[li];
    if (!(fl & kFSurvivor)) continue;
    uint32_t o  = pos[li];
    int      c2 = f.c2[li], ir1 = f.ir1[li], nr0 = f.nr0[li], nr1 = f.nr1[li];
    m.unitig[o] = uid_of_line[li];
    m.read[o]   = rid_of_line[li];
    m.ir0[o]    = c2;
    m.ir1[o]    = ir1;
    m.nr0[o]    = nr0;
    m.nr1[o]    = nr1;
    m.score[o]  = f.c9[li];
    m.line[o]   = uint32_t(li);
    m.flags[o]  = ((fl & kFDir) ? 1 : 0) | ((fl & kFPrim) ? 2 : 0);
    // BlastFileReader.cpp:118-119 — one IEEE-754 double division, operands exact
    m.rratio[o] = __ddiv_rn(double(ir1 - c2 + 1), double(nr1 - nr0 + 1));
  }
}

__global__ void __launch_bounds__(256)
    check_grouped_kernel(const uint32_t *__restrict__ unitig, const unsigned *__restrict__ d_n, uint64_t n_host,
                         uint32_t *unsorted) {
  pdl_sync();
  const uint64_t n = d_n ? uint64_t(*d_n) : n_host;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x + 1; i < n;
       i += uint64_t(gridDim.x) * blockDim.x)
    if (unitig[i] < unitig[i - 1]) *unsorted = 1;
}

template <class T>
__global__ void __launch_bounds__(256)
    gather_kernel(const T *__restrict__ in, const uint32_t *__restrict__ idx, uint64_t n, T *__restrict__ out) {
  pdl_sync();
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x)
    out[i] = in[idx[i]];
}

__global__ void __launch_bounds__(256) iota_kernel(uint32_t *out, uint64_t n) {
  pdl_sync();
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x)
    out[i] = uint32_t(i);
}

// grp_off[u] = first row of unitig u.  Every unitig id owns at least one row (its first accepted line survives the
// dedup), so boundaries alone define the table; ids without rows (possible for caller-packed records) are filled by
// running a max-scan backwards on the host-free path below.
__global__ void __launch_bounds__(256)
    group_bounds_kernel(const uint32_t *__restrict__ unitig, uint64_t n, uint32_t n_unitigs,
                        uint32_t *__restrict__ grp_off) {
  pdl_sync();
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i <= n; i += uint64_t(gridDim.x) * blockDim.x) {
    uint32_t lo = i == 0 ? 0u : unitig[i - 1] + 1u;       // first id not yet started before row i
    uint32_t hi = i == n ? n_unitigs : unitig[i];         // ids lo..hi start at row i
    if (i == 0 || i == n || unitig[i] != unitig[i - 1])
      for (uint32_t u = lo; u <= hi && u <= n_unitigs; ++u) grp_off[u] = uint32_t(i);
  }
}

// The scrubber pairs the reads of a BLOCK: consecutive processed lines that name the same anchor
// (scrubber_bfs.py:103-113).  Records are in line order; a record starts a block iff its anchor differs from its
// predecessor's.
struct BlockStart {
  const uint32_t *unitig;
  __device__ uint32_t operator()(uint64_t i) const { return (i == 0 || unitig[i] != unitig[i - 1]) ? 1u : 0u; }
};
__global__ void __launch_bounds__(256)
    block_ids_kernel(const uint32_t *__restrict__ unitig, uint64_t n, const uint32_t *__restrict__ before,
                     uint32_t *__restrict__ block, uint32_t *__restrict__ block_unitig) {
  pdl_sync();
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
    const bool     start = i == 0 || unitig[i] != unitig[i - 1];
    const uint32_t b     = before[i] + (start ? 1u : 0u) - 1u;
    block[i]             = b;
    if (start) block_unitig[b] = unitig[i];
  }
}

template <class T> static T read_back(Arena &arena, cudaStream_t st, const T *d);

uint64_t run_scrubber_blocks(Arena &arena, cudaStream_t st, const MatchSoA &m, uint32_t **block_out,
                             uint32_t **block_unitig_out) {
  const uint64_t M = m.n;
  uint32_t *before = arena.alloc<uint32_t>(M + 1), *block = arena.alloc<uint32_t>(M + 1);
  uint32_t *block_unitig = arena.alloc<uint32_t>(M + 1), *d_n = arena.alloc<uint32_t>(1);
  exclusive_scan<uint32_t>(arena, st, BlockStart{m.unitig}, M, before, d_n);
  if (M) MSGPU_LAUNCH("block_ids", st, launch_k(block_ids_kernel, grid_for(M, 256), 256, 0, st, m.unitig, M, before, block, block_unitig));
  *block_out        = block;
  *block_unitig_out = block_unitig;
  return uint64_t(read_back(arena, st, d_n));
}

static uint64_t next_pow2(uint64_t x) {
  uint64_t p = 1024;
  while (p < x) p <<= 1;
  return p;
}

// the three open-addressing tables of an ingest (reads, unitigs, (unitig, read) keys), same capacity each; their keys
// are one allocation and their min-line words another, so two memsets initialise all of them
static void make_tables(Arena &arena, cudaStream_t st, uint64_t min_slots, InternTable (&t)[3]) {
  const uint64_t      cap     = next_pow2(min_slots);
  unsigned long long *keys    = arena.alloc<unsigned long long>(3 * cap);
  uint32_t           *minline = arena.alloc<uint32_t>(3 * cap);
  for (int k = 0; k < 3; ++k) {
    t[k].keys    = keys + k * cap;
    t[k].minline = minline + k * cap;
    t[k].id      = arena.alloc<uint32_t>(cap);
    t[k].mask    = cap - 1;
  }
  MSGPU_CUDA(cudaMemsetAsync(keys, 0, 3 * cap * sizeof(unsigned long long), st));
  MSGPU_CUDA(cudaMemsetAsync(minline, 0xff, 3 * cap * sizeof(uint32_t), st));
}

template <class T> static T read_back(Arena &arena, cudaStream_t st, const T *d) {
  static_assert(sizeof(T) <= 1024, "host scratch");
  void *h = arena.host_scratch(); // pinned: the copy is a plain DMA, not the driver's staging path
  MSGPU_CUDA(cudaMemcpyAsync(h, d, sizeof(T), cudaMemcpyDeviceToHost, st));
  MSGPU_CUDA(cudaStreamSynchronize(st));
  T out;
  memcpy(&out, h, sizeof(T));
  return out;
}

// lengths of the first *d_n names, 0 behind them: lets a scan run over the capacity without knowing the count
struct NameLen {
  const uint32_t *len;
  const unsigned *d_n;
  __device__ uint64_t operator()(uint64_t i) const { return i < *d_n ? uint64_t(len[i]) : 0ull; }
};

// name_off[i] = exclusive prefix of the name lengths for i <= count (name_off[count] is the total); names gathered
// from the PAF text.  `cap` bounds the counts, `cap_bytes` the total lengths.  Both tables at once: their two scans
// share a launch.
struct NameJob {
  const unsigned     *d_n;
  const uint64_t     *src;
  const uint32_t     *len;
  unsigned long long *d_total_bytes;
  NameTable          *out;
};
static void build_name_tables(Arena &arena, cudaStream_t st, const char *d_paf, uint64_t cap, uint64_t cap_bytes,
                              const NameJob &a, const NameJob &b) {
  a.out->name_off = arena.alloc<uint64_t>(cap + 1);
  b.out->name_off = arena.alloc<uint64_t>(cap + 1);
  run_scans(arena, st,
            scan_job<uint64_t>(arena, st, 0, NameLen{a.len, a.d_n}, cap + 1, a.out->name_off,
                               reinterpret_cast<uint64_t *>(a.d_total_bytes), ScanLive{a.d_n, nullptr, 1}),
            scan_job<uint64_t>(arena, st, 1, NameLen{b.len, b.d_n}, cap + 1, b.out->name_off,
                               reinterpret_cast<uint64_t *>(b.d_total_bytes), ScanLive{b.d_n, nullptr, 1}));
  for (const NameJob *j : {&a, &b}) {
    j->out->names = arena.alloc<char>(cap_bytes);
    if (cap) {
      MSGPU_LAUNCH("gather_names", st, launch_k(gather_names_kernel, grid_for(cap * 32, 256, 2), 256, 0, st, d_paf, j->src, j->out->name_off, j->d_n, j->out->names));
    }
  }
}

void build_group_offsets(Arena &arena, cudaStream_t st, const MatchSoA &m, UnitigTable &unitigs) {
  unitigs.grp_off = arena.alloc<uint32_t>(unitigs.n + 1);
  MSGPU_LAUNCH("group_bounds", st, launch_k(group_bounds_kernel, grid_for(m.n + 1, 256), 256, 0, st, m.unitig, m.n, uint32_t(unitigs.n), unitigs.grp_off));
}

void alloc_match_soa(Arena &arena, MatchSoA &s, uint64_t M) {
  s.n      = M;
  s.unitig = arena.alloc<uint32_t>(M);
  s.read   = arena.alloc<uint32_t>(M);
  s.ir0    = arena.alloc<int32_t>(M);
  s.ir1    = arena.alloc<int32_t>(M);
  s.nr0    = arena.alloc<int32_t>(M);
  s.nr1    = arena.alloc<int32_t>(M);
  s.score  = arena.alloc<int32_t>(M);
  s.line   = arena.alloc<uint32_t>(M);
  s.flags  = arena.alloc<uint8_t>(M);
  s.rratio = arena.alloc<double>(M);
}

// Rows arrive in line order.  If they are not already grouped by unitig id (minimap2 emits by query, so they normally
// are), a stable sort by unitig id groups them and keeps the line order inside a unitig.
void group_by_unitig(Arena &arena, cudaStream_t st, MatchSoA &m, uint64_t n_unitigs, int known_unsorted) {
  const uint64_t M = m.n;
  if (M <= 1 || known_unsorted == 0) return;
  if (known_unsorted < 0) {
    uint32_t *d_unsorted = arena.alloc<uint32_t>(1);
    MSGPU_CUDA(cudaMemsetAsync(d_unsorted, 0, sizeof(uint32_t), st));
    MSGPU_LAUNCH("check_grouped", st, launch_k(check_grouped_kernel, grid_for(M, 256), 256, 0, st, m.unitig, nullptr, M, d_unsorted));
    if (!read_back(arena, st, d_unsorted)) return;
  }
  uint32_t *keys = arena.alloc<uint32_t>(M), *keys2 = arena.alloc<uint32_t>(M);
  uint32_t *idx = arena.alloc<uint32_t>(M), *idx2 = arena.alloc<uint32_t>(M);
  int       g = grid_for(M, 256);
  MSGPU_CUDA(cudaMemcpyAsync(keys, m.unitig, M * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  MSGPU_LAUNCH("iota", st, launch_k(iota_kernel, g, 256, 0, st, idx, M));
  radix_sort_pairs<uint32_t, uint32_t>(arena, st, keys, idx, keys2, idx2, M, {0, bits_for(n_unitigs ? n_unitigs : 1)});
  MatchSoA s;
  alloc_match_soa(arena, s, M);
  MSGPU_LAUNCH("gather", st, gather_kernel<<<g, 256, 0, st>>>(m.unitig, idx, M, s.unitig));
  MSGPU_LAUNCH("gather", st, gather_kernel<<<g, 256, 0, st>>>(m.read, idx, M, s.read));
  MSGPU_LAUNCH("gather", st, gather_kernel<<<g, 256, 0, st>>>(m.ir0, idx, M, s.ir0));
  MSGPU_LAUNCH("gather", st, gather_kernel<<<g, 256, 0, st>>>(m.ir1, idx, M, s.ir1));
  MSGPU_LAUNCH("gather", st, gather_kernel<<<g, 256, 0, st>>>(m.nr0, idx, M, s.nr0));
  MSGPU_LAUNCH("gather", st, gather_kernel<<<g, 256, 0, st>>>(m.nr1, idx, M, s.nr1));
  MSGPU_LAUNCH("gather", st, gather_kernel<<<g, 256, 0, st>>>(m.score, idx, M, s.score));
  MSGPU_LAUNCH("gather", st, gather_kernel<<<g, 256, 0, st>>>(m.line, idx, M, s.line));
  MSGPU_LAUNCH("gather", st, gather_kernel<<<g, 256, 0, st>>>(m.flags, idx, M, s.flags));
  MSGPU_LAUNCH("gather", st, gather_kernel<<<g, 256, 0, st>>>(m.rratio, idx, M, s.rratio));
  m = s;
}

void run_ingest(Arena &arena, cudaStream_t st, const char *d_paf, uint64_t n_bytes, uint64_t *n_lines_out,
                ReadTable &reads, UnitigTable &unitigs, MatchSoA &m, bool drop_last_line, int mode) {
  const bool scrubber = mode != 0; // (1: scrubber, 2: unitig filter — both keep the records in line order)
  reads = ReadTable{};
  unitigs = UnitigTable{};
  m = MatchSoA{};
  *n_lines_out = 0;

  // Capacities from the byte count alone, so that nothing has to come back to the host before the end:
  //   a line the reader accepts or rejects without throwing has >= 10 tab-separated fields = >= 20 bytes, so a valid
  //   file of n bytes has <= n / 20 + 1 lines (more lines: some parsed line is invalid, BlastFileReader.cpp:97-98)
  const uint64_t cap = n_bytes / 20 + 2;
  if (cap >= 0xfffffff0ull) throw BuildError{MSGPU_ERR_LIMIT, "PAF buffers of 80 GB and more must be split into shards"};

  // ---- K1a line index
  uint64_t     mis      = reinterpret_cast<uintptr_t>(d_paf) & 15u;
  const uint4 *aligned  = reinterpret_cast<const uint4 *>(d_paf - mis);
  uint64_t     nv       = mis + n_bytes;
  uint64_t     n_tiles  = (nv + kNlTile - 1) / kNlTile;
  uint32_t    *tile_cnt = arena.alloc<uint32_t>(n_tiles + 1);
  uint32_t    *d_n_nl   = arena.alloc<uint32_t>(1);
  IngestStatus *d_status = arena.alloc<IngestStatus>(1);
  uint64_t    *line_start = arena.alloc<uint64_t>(cap + 2);
  if (n_tiles) {
    MSGPU_LAUNCH("newline_count", st, launch_k(newline_count_kernel, (unsigned)n_tiles, kNlBlock, 0, st, aligned, mis, nv, tile_cnt));
  }
  exclusive_scan<uint32_t>(arena, st, LoadAs<uint32_t, uint32_t>{tile_cnt}, n_tiles, tile_cnt, d_n_nl);
  if (n_tiles) {
    MSGPU_LAUNCH("newline_write", st, launch_k(newline_write_kernel, (unsigned)n_tiles, kNlBlock, 0, st, aligned, mis, nv, tile_cnt, cap, line_start));
  }

  LineFields f;
  f.c2        = arena.alloc<int32_t>(cap);
  f.ir1       = arena.alloc<int32_t>(cap);
  f.c6        = arena.alloc<int32_t>(cap);
  f.nr0       = arena.alloc<int32_t>(cap);
  f.nr1       = arena.alloc<int32_t>(cap);
  f.c9        = arena.alloc<int32_t>(cap);
  f.rname_rel = arena.alloc<uint32_t>(cap);
  f.rname_len = arena.alloc<uint32_t>(cap);
  f.uname_len = arena.alloc<uint32_t>(cap);
  f.rhash     = arena.alloc<uint64_t>(cap);
  f.uhash     = arena.alloc<uint64_t>(cap);
  f.flags     = arena.alloc<uint8_t>(cap);
  uint32_t *rslot = arena.alloc<uint32_t>(cap), *uslot = arena.alloc<uint32_t>(cap);
  uint32_t *dslot = arena.alloc<uint32_t>(cap);
  uint32_t *rid_of_line = arena.alloc<uint32_t>(cap), *uid_of_line = arena.alloc<uint32_t>(cap);
  uint32_t *rank_a = arena.alloc<uint32_t>(cap), *rank_b = arena.alloc<uint32_t>(cap);
  // ids are handed out to accepted lines only; the shortest line the filter accepts is 20 bytes as well (std::getline
  // takes empty fields: "\t\t0\t400\t\t\t1\t0\t1\t400\n" — both names and the strand empty), so the id-indexed tables,
  // the match table and the hash tables (load factor <= 2/3 in the worst case, ~1/4 on real PAF) share the line bound
  const uint64_t cap_acc = cap;
  reads.nlen = arena.alloc<int32_t>(cap_acc);
  reads.line = arena.alloc<uint32_t>(cap_acc);
  uint64_t *rsrc = arena.alloc<uint64_t>(cap_acc), *usrc = arena.alloc<uint64_t>(cap_acc);
  uint32_t *rlen = arena.alloc<uint32_t>(cap_acc), *ulen = arena.alloc<uint32_t>(cap_acc);
  alloc_match_soa(arena, m, cap_acc);

  const int    grid_lines = grid_for(cap, 256);
  IngestStatus h{};
  for (int attempt = 0;; ++attempt) {
    if (attempt == 4) throw BuildError{MSGPU_ERR_INTERNAL, "name hash collisions persisted over 4 seeds"};
    IngestStatus init{};
    init.first_bad_line = ~0ull;
    MSGPU_CUDA(cudaMemcpyAsync(d_status, &init, sizeof(init), cudaMemcpyHostToDevice, st));
    MSGPU_CUDA(cudaMemsetAsync(f.flags, 0, cap, st)); // the scans below run over the capacity
    MSGPU_LAUNCH("line_index_finish", st, launch_k(line_index_finish_kernel, 1, 1, 0, st, d_paf, n_bytes, d_n_nl, cap, drop_last_line ? 1 : 0, line_start, d_status));
    // seeds are odd multiples of a constant, never 0; the test hook makes the first attempt collide on purpose
    const bool     force = attempt == 0 && std::getenv("MSGPU_TEST_FORCE_COLLISION") != nullptr;
    const uint64_t seed  = force ? 0ull : 0x243f6a8885a308d3ull * uint64_t(attempt + 1);
    if (mode == 1)
      MSGPU_LAUNCH("parse_lines", st, launch_k(parse_lines_kernel<1>, grid_for(cap, kParseBlock, 4), kParseBlock, 0, st, d_paf, n_bytes, line_start, seed, f, d_status));
    else if (mode == 2)
      MSGPU_LAUNCH("parse_lines", st, launch_k(parse_lines_kernel<2>, grid_for(cap, kParseBlock, 4), kParseBlock, 0, st, d_paf, n_bytes, line_start, seed, f, d_status));
    else
      MSGPU_LAUNCH("parse_lines", st, launch_k(parse_lines_kernel<0>, grid_for(cap, kParseBlock, 4), kParseBlock, 0, st, d_paf, n_bytes, line_start, seed, f, d_status));

    const ScanLive live_lines{nullptr, &d_status->n_lines, 0}; // the per-line scans stop at the actual line count
    // ---- K3 interning
    InternTable tables[3];
    make_tables(arena, st, cap_acc + cap_acc / 2, tables);
    const InternTable rt = tables[0], ut = tables[1], dt = tables[2];
    MSGPU_LAUNCH("intern_insert", st, launch_k(intern_insert_kernel, dim3(grid_lines, 2), 256, 0, st, f.flags, f.rhash, f.uhash, d_status, rt, ut, rslot, uslot));
    MSGPU_LAUNCH("intern_verify", st, launch_k(intern_verify_kernel, grid_lines, 256, 0, st, d_paf, line_start, f, rt, rslot, ut, uslot, d_status));
    run_scans(arena, st,
              scan_job<uint32_t>(arena, st, 0, FlagBit<uint32_t>{f.flags, kFFirstRead}, cap, rank_a, &d_status->n_reads, live_lines),
              scan_job<uint32_t>(arena, st, 1, FlagBit<uint32_t>{f.flags, kFFirstUnitig}, cap, rank_b, &d_status->n_unitigs, live_lines));
    MSGPU_LAUNCH("intern_assign", st, launch_k(intern_assign_kernel, grid_lines, 256, 0, st, line_start, d_status, f, rt, rslot, rank_a, ut, uslot, rank_b,
                                                     reads.nlen, reads.line, rsrc, rlen, usrc, ulen));
    build_name_tables(arena, st, d_paf, cap_acc, n_bytes,
                      NameJob{&d_status->n_reads, rsrc, rlen, &d_status->read_name_bytes, &reads},
                      NameJob{&d_status->n_unitigs, usrc, ulen, &d_status->unitig_name_bytes, &unitigs});

    // ---- K4 dedup
    if (mode == 2) {
      MSGPU_LAUNCH("all_survive", st, launch_k(all_survive_kernel, grid_lines, 256, 0, st, d_status, f.flags, rt, rslot, ut, uslot, rid_of_line, uid_of_line));
    } else {
      MSGPU_LAUNCH("dedup_insert", st, launch_k(dedup_insert_kernel, grid_lines, 256, 0, st, d_status, f, rt, rslot, ut, uslot, dt, dslot, rid_of_line, uid_of_line));
      MSGPU_LAUNCH("dedup_flag", st, launch_k(dedup_flag_kernel, grid_lines, 256, 0, st, d_status, f.flags, dt, dslot));
    }

    // ---- K5 compaction
    exclusive_scan<uint32_t>(arena, st, FlagBit<uint32_t>{f.flags, kFSurvivor}, cap, rank_a, &d_status->n_matches, live_lines);
    MSGPU_LAUNCH("compact_matches", st, launch_k(compact_matches_kernel, grid_lines, 256, 0, st, d_status, f, rank_a, rid_of_line, uid_of_line, m));
    MSGPU_LAUNCH("check_grouped", st, launch_k(check_grouped_kernel, grid_for(cap_acc, 256, 2), 256, 0, st, m.unitig, &d_status->n_matches, 0, &d_status->unsorted));

    // ---- the one round trip of the ingest
    h = read_back(arena, st, d_status);
    if (h.overflow || h.first_bad_line != ~0ull)
      throw BuildError{MSGPU_ERR_INVALID_PAF,
                       h.first_bad_line != ~0ull ? "Invalid BLAST file. (line " + std::to_string(h.first_bad_line) + ")"
                                                 : std::string("Invalid BLAST file. (lines shorter than 10 fields)")};
    if (!h.collision) break; // else: two different names with the same 64-bit hash (astronomically rare): new seed
  }
  if (h.n_lines + 2 >= 0xffffffffull) throw BuildError{MSGPU_ERR_LIMIT, "more than 2^32-2 lines in one PAF"};
  *n_lines_out    = h.n_lines;
  reads.n         = h.n_reads;
  reads.n_bytes   = h.read_name_bytes;
  unitigs.n       = h.n_unitigs;
  unitigs.n_bytes = h.unitig_name_bytes;
  m.n             = h.n_matches;
  if (scrubber) return; // records stay in line order: the scrubber pairs blocks of consecutive lines (run_scrubber_blocks)
  // grouping sort only when the PAF was not grouped by unitig (minimap2 emits by query, so it normally is)
  group_by_unitig(arena, st, m, h.n_unitigs, h.unsorted ? 1 : 0);
  build_group_offsets(arena, st, m, unitigs);
}

} // namespace msgpu
