/* ms_oracle.c — plain-C CPU restatement of MuCHSALSA's graph-construction hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the checker the CUDA path is compared with; it is never linked into, imported by
 * or called from the product (muchsalsa_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs
 * may use it.
 *
 * PARITY PINNED: the reference's own tests do not cover this path, so this restatement is pinned
 * against (1) the 8-line known-answer PAF recorded from the reference (tests/golden/kat8.*),
 * (2) the reference's sequence-store golden strings (libms/tests/SA_test.cpp:11-136 on test_data/fasta.fa, fastq.fq),
 * (3) outputs of the reference ITSELF, compiled from /root/reference into oracle/_ref/libmsref.so, on the synthetic
 * configurations (tests/test_oracle.py; digests committed under tests/golden/ by tests/golden/make_golden.py).
 *
 * What is restated (all file:line relative to /root/reference), for the reference run with ONE worker thread — its only
 * deterministic mode (SURVEY.md §8a rules i-vii):
 *   line index          BlastFileAccessor.cpp:77-91 + IO.cpp:54-97 (a trailing un-terminated line counts as a line)
 *   read()              BlastFileReader.cpp:72-84   (lines 0 .. lineCount-2; the LAST line is never parsed)
 *   parseLine()         BlastFileReader.cpp:86-130  (tab split, std::stoi fields, accept filter, rRatio, isPrimary)
 *   Registry            Registry.cpp:36-45          (id = rank of first request; reads and unitigs independent)
 *   addVertex           Graph.cpp:141-149           (emplace: first Vertex of a read wins -> min accepted line)
 *   addVertexMatch      MatchMap.cpp:53-81          (per (read, unitig) keep the smallest line)
 *   processScaffold     MatchMap.cpp:175-224        (sort by line; all pairs; overlap > 100; dir; prim; f64 score;
 *                                                    vertex order by Vertex lineIdx)
 *   addEdge             Graph.cpp:212-230,291-311   (an edge exists iff >= 1 accepted pair; first insertion fixes order)
 *   strSlice / getReverseComplement / get*Sequence      SequenceUtils.cpp:27-85
 *   sequence index + fetch                              SequenceAccessor.cpp:54-85,143-231
 *
 * The data structures are NOT the reference's (it uses nested unordered_maps and shared_ptrs); sorting replaces
 * hashing wherever the reference's result does not depend on hash order.
 */
#include <ctype.h>
#include <limits.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ms_tables.h"

/* thresholds: BlastFileReader.cpp:48-50, MatchMap.cpp:41 */
#define MINIMUM_MATCHES 400
#define TH_LENGTH 500
#define TH_MATCHES 500
#define TH_OVERLAP 100

#define MSO_ERR_INVALID 1 /* the reference would throw "Invalid BLAST file." / std::invalid_argument / out_of_range */
#define MSO_ERR_NOMEM 2

void mso_free(void *p) { free(p); }

void ms_tables_free(ms_tables *t) {
  if (!t) return;
  void *ptrs[] = {t->read_name_off, t->read_names,    t->read_nlen,  t->read_line, t->unitig_name_off, t->unitig_names,
                  t->vm_read,       t->vm_unitig,     t->vm_nr0,     t->vm_nr1,    t->vm_ir0,          t->vm_ir1,
                  t->vm_rratio,     t->vm_dir,        t->vm_prim,    t->vm_score,  t->vm_line,         t->edge_first,
                  t->edge_second,   t->edge_em_begin, t->em_unitig,  t->em_ov0,    t->em_ov1,          t->em_score,
                  t->em_dir,        t->em_prim,       t->em_line};
  for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); ++i) free(ptrs[i]);
  memset(t, 0, sizeof(*t));
}

/* std::stoi on a token [s, s+n): strtol base 10 semantics — leading isspace skipped, optional sign, at least one
 * digit, trailing junk ignored, value must fit int.  Returns 0 on success. */
static int mso_stoi(const char *s, size_t n, int *out) {
  size_t i = 0;
  while (i < n && isspace((unsigned char)s[i])) ++i;
  int neg = 0;
  if (i < n && (s[i] == '+' || s[i] == '-')) {
    neg = s[i] == '-';
    ++i;
  }
  if (i >= n || s[i] < '0' || s[i] > '9') return MSO_ERR_INVALID;
  int64_t v = 0;
  for (; i < n && s[i] >= '0' && s[i] <= '9'; ++i) {
    v = v * 10 + (s[i] - '0');
    if (v > (int64_t)INT_MAX + 1) return MSO_ERR_INVALID; /* out_of_range */
  }
  if (neg) v = -v;
  if (v > INT_MAX || v < INT_MIN) return MSO_ERR_INVALID;
  *out = (int)v;
  return 0;
}

/* ---- string interning: open addressing, FNV-1a ------------------------------------------------------------------ */
typedef struct {
  const char *base;
  uint64_t   *slot_off; /* offset of the representative name, UINT64_MAX = empty */
  uint32_t   *slot_len;
  uint32_t   *slot_id;
  uint64_t    cap, n;
} strtab;

static int strtab_init(strtab *t, const char *base, uint64_t expect) {
  t->base = base;
  t->cap  = 16;
  while (t->cap < expect * 2 + 16) t->cap <<= 1;
  t->n        = 0;
  t->slot_off = malloc(t->cap * sizeof(uint64_t));
  t->slot_len = malloc(t->cap * sizeof(uint32_t));
  t->slot_id  = malloc(t->cap * sizeof(uint32_t));
  if (!t->slot_off || !t->slot_len || !t->slot_id) return MSO_ERR_NOMEM;
  memset(t->slot_off, 0xff, t->cap * sizeof(uint64_t));
  return 0;
}
static void strtab_free(strtab *t) {
  free(t->slot_off);
  free(t->slot_len);
  free(t->slot_id);
}
/* Registry::operator[] : returns the id, assigning the next free one on first sight (Registry.cpp:36-45) */
static uint32_t strtab_intern(strtab *t, uint64_t off, uint32_t len, int *is_new) {
  uint64_t h = 1469598103934665603ull;
  for (uint32_t i = 0; i < len; ++i) h = (h ^ (unsigned char)t->base[off + i]) * 1099511628211ull;
  uint64_t p = h & (t->cap - 1);
  for (;;) {
    if (t->slot_off[p] == UINT64_MAX) {
      t->slot_off[p] = off;
      t->slot_len[p] = len;
      t->slot_id[p]  = (uint32_t)t->n++;
      *is_new        = 1;
      return t->slot_id[p];
    }
    if (t->slot_len[p] == len && memcmp(t->base + t->slot_off[p], t->base + off, len) == 0) {
      *is_new = 0;
      return t->slot_id[p];
    }
    p = (p + 1) & (t->cap - 1);
  }
}

/* ---- accepted match rows ---------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t read, unitig;
  int32_t  nr0, nr1, ir0, ir1;
  double   rratio;
  uint64_t score, line;
  uint8_t  dir, prim;
} vmatch;

typedef struct {
  uint32_t first, second, unitig;
  int32_t  ov0, ov1;
  double   score;
  uint64_t line;
  uint8_t  dir, prim;
} ematch;

static int cmp_vm_unitig_read_line(const void *a, const void *b) {
  const vmatch *x = a, *y = b;
  if (x->unitig != y->unitig) return x->unitig < y->unitig ? -1 : 1;
  if (x->read != y->read) return x->read < y->read ? -1 : 1;
  return x->line < y->line ? -1 : (x->line > y->line);
}
static int cmp_vm_unitig_line(const void *a, const void *b) {
  const vmatch *x = a, *y = b;
  if (x->unitig != y->unitig) return x->unitig < y->unitig ? -1 : 1;
  return x->line < y->line ? -1 : (x->line > y->line);
}
static int cmp_vm_read_unitig(const void *a, const void *b) {
  const vmatch *x = a, *y = b;
  if (x->read != y->read) return x->read < y->read ? -1 : 1;
  return x->unitig < y->unitig ? -1 : (x->unitig > y->unitig);
}
static int cmp_em(const void *a, const void *b) {
  const ematch *x = a, *y = b;
  if (x->first != y->first) return x->first < y->first ? -1 : 1;
  if (x->second != y->second) return x->second < y->second ? -1 : 1;
  return x->unitig < y->unitig ? -1 : (x->unitig > y->unitig);
}

#define ALLOC(ptr, n)                                              \
  do {                                                             \
    (ptr) = malloc(((n) ? (n) : 1) * sizeof(*(ptr)));              \
    if (!(ptr)) {                                                  \
      rc = MSO_ERR_NOMEM;                                          \
      goto done;                                                   \
    }                                                              \
  } while (0)

int mso_build(const char *paf, uint64_t nbytes, ms_tables *out) {
  int       rc = 0;
  uint64_t *line_off = NULL, *name_off[2] = {NULL, NULL};
  vmatch   *vm = NULL;
  ematch   *em = NULL;
  uint64_t *read_first_line = NULL, *read_nlen = NULL, *grp_off = NULL;
  strtab    reads, unitigs;
  int       tabs_ready = 0;
  memset(out, 0, sizeof(*out));

  /* -- line index: BlastFileAccessor::_buildIndex + readline: every '\n' ends a line; trailing bytes without '\n'
   *    form one more line (IO.cpp:92-95) */
  uint64_t n_lines = 0;
  for (uint64_t i = 0; i < nbytes; ++i) n_lines += paf[i] == '\n';
  if (nbytes && paf[nbytes - 1] != '\n') ++n_lines;
  ALLOC(line_off, n_lines + 1);
  {
    uint64_t k = 0, start = 0;
    for (uint64_t i = 0; i < nbytes; ++i)
      if (paf[i] == '\n') {
        line_off[k++] = start;
        start         = i + 1;
      }
    if (start < nbytes) line_off[k++] = start;
    line_off[k] = nbytes + (nbytes && paf[nbytes - 1] != '\n'); /* pretend terminator so len = next - off - 1 */
  }

  /* -- read(): lines 0 .. n_lines-2 (BlastFileReader.cpp:76; size_t underflow for an empty file is not restated) */
  uint64_t n_parse = n_lines ? n_lines - 1 : 0;
  ALLOC(vm, n_parse);
  if ((rc = strtab_init(&reads, paf, n_parse)) || (rc = strtab_init(&unitigs, paf, n_parse))) goto done;
  tabs_ready = 1;
  ALLOC(read_first_line, n_parse);
  ALLOC(read_nlen, n_parse);
  ALLOC(name_off[0], 2 * n_parse + 2);
  ALLOC(name_off[1], 2 * n_parse + 2);

  uint64_t n_acc = 0;
  for (uint64_t li = 0; li < n_parse; ++li) {
    /* getLine(): the line without its last character (BlastFileAccessor.cpp:60-63) */
    const char *s   = paf + line_off[li];
    uint64_t    len = line_off[li + 1] - line_off[li] - 1;
    /* std::getline(iss, token, '\t') split: a trailing empty field is not produced; an empty line has no fields */
    uint64_t tok_off[10], tok_len[10];
    int      ntok = 0;
    uint64_t p    = 0;
    while (p < len && ntok < 10) {
      uint64_t q = p;
      while (q < len && s[q] != '\t') ++q;
      tok_off[ntok] = p;
      tok_len[ntok] = q - p;
      ++ntok;
      p = q + 1;
    }
    /* BlastFileReader.cpp:97 checks `< 9` and then reads field 9; fewer than 10 fields is treated as invalid here */
    if (ntok < 10) {
      rc = MSO_ERR_INVALID;
      goto done;
    }
    int c2, c3, c9, c6;
    if (mso_stoi(s + tok_off[2], tok_len[2], &c2) || mso_stoi(s + tok_off[3], tok_len[3], &c3) ||
        mso_stoi(s + tok_off[9], tok_len[9], &c9) || mso_stoi(s + tok_off[6], tok_len[6], &c6)) {
      rc = MSO_ERR_INVALID;
      goto done;
    }
    int      ir0 = c2, ir1 = c3 - 1;               /* :101 */
    uint64_t matches = (uint64_t)(int64_t)c9;      /* :102 static_cast<size_t>(int) */
    int      ilen    = ir1 - ir0 + 1;
    int      add     = matches >= MINIMUM_MATCHES; /* :106 */
    add &= ilen >= MINIMUM_MATCHES;                /* :107 */
    if (!add) continue;

    int      is_new;
    uint32_t rid = strtab_intern(&reads, (uint64_t)(s - paf) + tok_off[5], (uint32_t)tok_len[5], &is_new); /* :110 */
    if (is_new) {
      /* Graph::_addVertex emplace: only the first Vertex of an id is kept (Graph.cpp:148) */
      read_first_line[rid]      = li;
      read_nlen[rid]            = (uint64_t)(int64_t)c6;
      name_off[0][2 * rid]      = (uint64_t)(s - paf) + tok_off[5];
      name_off[0][2 * rid + 1]  = tok_len[5];
    }
    uint32_t uid = strtab_intern(&unitigs, (uint64_t)(s - paf) + tok_off[0], (uint32_t)tok_len[0], &is_new); /* :111 */
    if (is_new) {
      name_off[1][2 * uid]     = (uint64_t)(s - paf) + tok_off[0];
      name_off[1][2 * uid + 1] = tok_len[0];
    }
    int c7, c8;
    if (mso_stoi(s + tok_off[7], tok_len[7], &c7) || mso_stoi(s + tok_off[8], tok_len[8], &c8)) {
      rc = MSO_ERR_INVALID;
      goto done;
    }
    vmatch *m = &vm[n_acc++];
    m->read   = rid;
    m->unitig = uid;
    m->nr0    = c7;
    m->nr1    = c8 - 1;                                                        /* :116 */
    m->ir0    = ir0;
    m->ir1    = ir1;
    m->dir    = tok_len[4] == 1 && s[tok_off[4]] == '+';                       /* :117 */
    m->rratio = (double)(ir1 - ir0 + 1) / (double)(m->nr1 - m->nr0 + 1);       /* :118-119 */
    m->prim   = (ilen >= TH_LENGTH) & (matches >= TH_MATCHES);                 /* :121-122 */
    m->score  = matches;
    m->line   = li;
  }
  uint64_t n_reads = reads.n, n_unitigs = unitigs.n;

  /* -- addVertexMatch: per (read, unitig) the smallest line survives (MatchMap.cpp:64-73) */
  qsort(vm, n_acc, sizeof(vmatch), cmp_vm_unitig_read_line);
  uint64_t n_vm = 0;
  for (uint64_t i = 0; i < n_acc; ++i)
    if (i == 0 || vm[i].unitig != vm[i - 1].unitig || vm[i].read != vm[i - 1].read) vm[n_vm++] = vm[i];

  /* -- scaffolds: members of a unitig sorted by line (MatchMap.cpp:178-183) */
  qsort(vm, n_vm, sizeof(vmatch), cmp_vm_unitig_line);
  ALLOC(grp_off, n_unitigs + 2);
  {
    uint64_t i = 0;
    for (uint64_t u = 0; u <= n_unitigs; ++u) {
      while (i < n_vm && vm[i].unitig < u) ++i;
      grp_off[u] = i;
    }
    grp_off[n_unitigs + 1] = n_vm;
  }

  /* -- processScaffold: count, then emit */
  uint64_t n_pairs = 0;
  for (uint64_t u = 0; u < n_unitigs; ++u) {
    uint64_t k = grp_off[u + 1] - grp_off[u];
    n_pairs += k * (k ? k - 1 : 0) / 2;
  }
  ALLOC(em, n_pairs);
  uint64_t n_em = 0;
  for (uint64_t u = 0; u < n_unitigs; ++u) {
    for (uint64_t o = grp_off[u] + 1; o < grp_off[u + 1]; ++o) {
      const vmatch *outer = &vm[o];
      for (uint64_t i = grp_off[u]; i < o; ++i) {
        const vmatch *inner = &vm[i];
        int ov0 = outer->ir0 > inner->ir0 ? outer->ir0 : inner->ir0;
        int ov1 = outer->ir1 < inner->ir1 ? outer->ir1 : inner->ir1;
        if (!(ov0 <= ov1 && ov1 - ov0 > TH_OVERLAP)) continue;                 /* :192 */
        ematch *e = &em[n_em++];
        e->dir    = outer->dir == inner->dir;                                  /* :193 */
        e->prim   = outer->prim && inner->prim;                                /* :194 */
        double outer_len  = (double)(outer->ir1 - outer->ir0 + 1);
        double inner_len  = (double)(inner->ir1 - inner->ir0 + 1);
        double common     = (double)(ov1 - ov0 + 1);
        double outer_sc   = (double)outer->score * common / outer_len;         /* :199 */
        double inner_sc   = (double)inner->score * common / inner_len;         /* :200 */
        e->score          = outer_sc + inner_sc;
        /* vertex order: smaller Vertex lineIdx first (:204-213) */
        if (read_first_line[outer->read] < read_first_line[inner->read]) {
          e->first  = outer->read;
          e->second = inner->read;
        } else {
          e->first  = inner->read;
          e->second = outer->read;
        }
        e->unitig = (uint32_t)u;
        e->ov0    = ov0;
        e->ov1    = ov1;
        e->line   = outer->line;                                               /* :218 */
      }
    }
  }
  qsort(em, n_em, sizeof(ematch), cmp_em);
  uint64_t n_edges = 0;
  for (uint64_t i = 0; i < n_em; ++i)
    if (i == 0 || em[i].first != em[i - 1].first || em[i].second != em[i - 1].second) ++n_edges;

  /* -- flatten */
  out->n_reads = n_reads;
  out->n_unitigs = n_unitigs;
  out->n_vm = n_vm;
  out->n_edges = n_edges;
  out->n_em = n_em;
  ALLOC(out->read_nlen, n_reads);
  ALLOC(out->read_line, n_reads);
  memcpy(out->read_nlen, read_nlen, n_reads * sizeof(uint64_t));
  memcpy(out->read_line, read_first_line, n_reads * sizeof(uint64_t));
  for (int w = 0; w < 2; ++w) {
    uint64_t  n   = w ? n_unitigs : n_reads, tot = 0;
    uint64_t *off;
    char     *blob;
    ALLOC(off, n + 1);
    if (w) out->unitig_name_off = off; else out->read_name_off = off;
    for (uint64_t i = 0; i < n; ++i) tot += name_off[w][2 * i + 1];
    ALLOC(blob, tot);
    if (w) out->unitig_names = blob; else out->read_names = blob;
    tot = 0;
    for (uint64_t i = 0; i < n; ++i) {
      off[i] = tot;
      memcpy(blob + tot, paf + name_off[w][2 * i], name_off[w][2 * i + 1]);
      tot += name_off[w][2 * i + 1];
    }
    off[n] = tot;
  }
  qsort(vm, n_vm, sizeof(vmatch), cmp_vm_read_unitig);
  ALLOC(out->vm_read, n_vm); ALLOC(out->vm_unitig, n_vm); ALLOC(out->vm_nr0, n_vm); ALLOC(out->vm_nr1, n_vm);
  ALLOC(out->vm_ir0, n_vm); ALLOC(out->vm_ir1, n_vm); ALLOC(out->vm_rratio, n_vm); ALLOC(out->vm_dir, n_vm);
  ALLOC(out->vm_prim, n_vm); ALLOC(out->vm_score, n_vm); ALLOC(out->vm_line, n_vm);
  for (uint64_t i = 0; i < n_vm; ++i) {
    out->vm_read[i] = vm[i].read; out->vm_unitig[i] = vm[i].unitig; out->vm_nr0[i] = vm[i].nr0;
    out->vm_nr1[i] = vm[i].nr1; out->vm_ir0[i] = vm[i].ir0; out->vm_ir1[i] = vm[i].ir1;
    out->vm_rratio[i] = vm[i].rratio; out->vm_dir[i] = vm[i].dir; out->vm_prim[i] = vm[i].prim;
    out->vm_score[i] = vm[i].score; out->vm_line[i] = vm[i].line;
  }
  ALLOC(out->edge_first, n_edges); ALLOC(out->edge_second, n_edges); ALLOC(out->edge_em_begin, n_edges + 1);
  ALLOC(out->em_unitig, n_em); ALLOC(out->em_ov0, n_em); ALLOC(out->em_ov1, n_em); ALLOC(out->em_score, n_em);
  ALLOC(out->em_dir, n_em); ALLOC(out->em_prim, n_em); ALLOC(out->em_line, n_em);
  {
    uint64_t e = 0;
    for (uint64_t i = 0; i < n_em; ++i) {
      if (i == 0 || em[i].first != em[i - 1].first || em[i].second != em[i - 1].second) {
        out->edge_first[e] = em[i].first; out->edge_second[e] = em[i].second; out->edge_em_begin[e] = i;
        ++e;
      }
      out->em_unitig[i] = em[i].unitig; out->em_ov0[i] = em[i].ov0; out->em_ov1[i] = em[i].ov1;
      out->em_score[i] = em[i].score; out->em_dir[i] = em[i].dir; out->em_prim[i] = em[i].prim;
      out->em_line[i] = em[i].line;
    }
    out->edge_em_begin[n_edges] = n_em;
  }

done:
  free(line_off); free(vm); free(em); free(read_first_line); free(read_nlen); free(grp_off);
  free(name_off[0]); free(name_off[1]);
  if (tabs_ready) { strtab_free(&reads); strtab_free(&unitigs); }
  if (rc) ms_tables_free(out);
  return rc;
}

/* ===================================================================================================================
 * Sequence path
 * =================================================================================================================*/

/* strSlice(original, i, j) (SequenceUtils.cpp:27-38): Python-style negative indices, then an INCLUSIVE [start, end]
 * window; returns the slice length, or -1 where std::string_view::substr would throw (start > size). */
int64_t mso_strslice(const char *s, uint64_t size, int i, int j, char *out) {
  int isz = (int)size;
  if (i < 0) i = isz + i;
  if (j < 0) j = isz + j;
  uint64_t start = (uint64_t)(i > 0 ? i : 0);
  uint64_t jj    = (uint64_t)(j > 0 ? j : 0);
  uint64_t end   = jj < size ? jj : size;
  uint64_t ii    = (uint64_t)(int64_t)i; /* static_cast<std::size_t>(i): a still-negative i wraps to a huge value */
  if (ii > end) end = ii;
  if (start > size) return -1;
  uint64_t count = end - start + 1;      /* may wrap exactly like the reference's size_t arithmetic */
  uint64_t avail = size - start;
  if (count > avail) count = avail;
  memcpy(out, s + start, count);
  return (int64_t)count;
}

/* getReverseComplement (SequenceUtils.cpp:41-61): reverse, complement UPPER-CASE ACGT only */
void mso_revcomp(const char *s, uint64_t n, char *out) {
  for (uint64_t k = 0; k < n; ++k) {
    char c = s[n - 1 - k];
    switch (c) {
    case 'A': c = 'T'; break;
    case 'T': c = 'A'; break;
    case 'G': c = 'C'; break;
    case 'C': c = 'G'; break;
    default: break;
    }
    out[k] = c;
  }
}

/* readline() on a memory buffer (IO.cpp:54-97): returns the number of bytes consumed (including the '\n'), or -1 at
 * end of data */
static int64_t mem_readline(const char *d, uint64_t n, uint64_t *pos, uint64_t *line_start) {
  if (*pos >= n) return -1;
  *line_start = *pos;
  uint64_t p  = *pos;
  while (p < n && d[p] != '\n') ++p;
  if (p < n) ++p;
  int64_t r = (int64_t)(p - *pos);
  *pos      = p;
  return r;
}

/* Sequence index (SequenceAccessor.cpp:143-231).  fastq=0: description '>' and a record ends at the next '>' line
 * (used for the unitig file and for FASTA nanopore files); fastq=1: description '@', sequence ends at the '+' line,
 * then lines are skipped until the next one that begins with '@' (SequenceAccessor.cpp:184-187 — a quality line that
 * begins with '@' is therefore taken for a header, exactly like the reference).
 * Output rows: (name offset, name length, sequence byte offset, sequence byte length incl. newlines), file order.
 * Inputs on which the reference does not terminate (a header as the very last line) return an error instead. */
int mso_seq_index(const char *d, uint64_t n, int fastq, uint64_t **rows_out, uint64_t *n_out) {
  char      desc = fastq ? '@' : '>', split = fastq ? '+' : '>';
  uint64_t  cap = 64, cnt = 0, pos = 0, ls = 0;
  uint64_t *rows = malloc(cap * 4 * sizeof(uint64_t));
  if (!rows) return MSO_ERR_NOMEM;
  int64_t  ret = mem_readline(d, n, &pos, &ls);
  uint64_t offset_start = pos;
  while (ret != -1) {
    if (d[ls] == desc) break;
    ret          = mem_readline(d, n, &pos, &ls);
    offset_start = pos;
  }
  if (ret == -1) { /* no header at all: the reference then tests a stale buffer; an empty index is the sane reading */
    *rows_out = rows;
    *n_out    = 0;
    return 0;
  }
  while (d[ls] == desc) {
    /* id = header after the marker up to the first whitespace (cleanSequenceId, SequenceAccessor.cpp:80-85) */
    uint64_t name_off = ls + 1, name_len = 0;
    while (name_off + name_len < n && !isspace((unsigned char)d[name_off + name_len])) ++name_len;
    uint64_t length = 0;
    for (;;) {
      ret                 = mem_readline(d, n, &pos, &ls);
      uint64_t offset_end = pos;
      if (ret == -1 || d[ls] == split) {
        if (cnt == cap) {
          cap *= 2;
          uint64_t *nr = realloc(rows, cap * 4 * sizeof(uint64_t));
          if (!nr) { free(rows); return MSO_ERR_NOMEM; }
          rows = nr;
        }
        rows[4 * cnt + 0] = name_off; rows[4 * cnt + 1] = name_len;
        rows[4 * cnt + 2] = offset_start; rows[4 * cnt + 3] = length;
        ++cnt;
        offset_start = offset_end;
        break;
      }
      length += (uint64_t)ret;
    }
    if (ret == -1) {
      /* the reference re-tests its stale line buffer (the last line read): unless that is a header the loop ends;
       * if it is one (a header is the last line of the file) the reference never terminates */
      if (d[ls] == desc) { free(rows); return MSO_ERR_INVALID; }
      break;
    }
    while (ret != -1 && d[ls] != desc) {
      ret          = mem_readline(d, n, &pos, &ls);
      offset_start = pos;
    }
    if (ret == -1) break;
  }
  *rows_out = rows;
  *n_out    = cnt;
  return 0;
}

/* getSequenceFromFile (SequenceAccessor.cpp:54-69): `length` bytes at `off`, cut at the first NUL, isspace removed */
int64_t mso_seq_extract(const char *d, uint64_t n, uint64_t off, uint64_t length, char *out) {
  if (off > n || length > n - off) return -1; /* fread comes up short -> "Failed to read sequence." */
  uint64_t k = 0;
  for (uint64_t i = 0; i < length; ++i) {
    char c = d[off + i];
    if (c == '\0') break;
    if (!isspace((unsigned char)c)) out[k++] = c;
  }
  return (int64_t)k;
}
