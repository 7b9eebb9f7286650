"""Test helpers: canonical forms shared by the oracle, the reference and the CUDA path."""
from __future__ import annotations

import json
import os

import numpy as np

from oracle.pyoracle import Tables

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

KAT8 = """u1 5000 0    3000 + r1 20000 100  3100 2900 3000 60
u1 5000 1000 4000 - r2 15000 500  3500 2800 3000 60
u1 5000 2500 5000 + r3 18000 0    2500 450  2500 60
u2 4000 0    2000 + r2 15000 6000 8000 1900 2000 60
u2 4000 100  1800 + r3 18000 7000 8700 1600 1700 60
u1 5000 200  900  + r1 20000 9000 9700 650  700  60
u3 900  0    300  + r1 20000 0    300  290  300  60
u2 4000 0    2000 + r1 20000 100  2100 1999 2000 60
"""


def kat8_paf() -> bytes:
    """The 8-line known-answer PAF recorded from the reference, tab separated."""
    return ("\n".join("\t".join(l.split()) for l in KAT8.strip().splitlines()) + "\n").encode()


def paf_line(u, ulen, c2, c3, strand, r, rlen, c7, c8, c9, c10=None, mapq=60, extra=()) -> bytes:
    f = [u, ulen, c2, c3, strand, r, rlen, c7, c8, c9, c10 if c10 is not None else c3 - c2, mapq, *extra]
    return ("\t".join(str(x) for x in f) + "\n").encode()


def gpu_to_tables(res) -> Tables:
    """muchsalsa_b200.construct.Construction -> the canonical oracle Tables (vertex matches by (read, unitig))."""
    m = res.matches
    order = np.lexsort((m["unitig"], m["read"]))
    vm = dict(read=m["read"][order], unitig=m["unitig"][order], nr0=m["nr0"][order], nr1=m["nr1"][order],
              ir0=m["ir0"][order], ir1=m["ir1"][order], rratio=m["rratio"][order],
              dir=(m["flags"][order] & 1).astype(np.uint8), prim=((m["flags"][order] >> 1) & 1).astype(np.uint8),
              score=m["score"][order].astype(np.int64).astype(np.uint64),
              line=m["line"][order].astype(np.uint64))
    e = res.edge_matches
    em = dict(unitig=e["unitig"], ov0=e["ov0"], ov1=e["ov1"], score=e["score"],
              dir=(e["flags"] & 1).astype(np.uint8), prim=((e["flags"] >> 1) & 1).astype(np.uint8),
              line=e["line"].astype(np.uint64))
    return Tables(read_names=res.read_names, read_nlen=res.read_nlen.astype(np.int64).astype(np.uint64),
                  read_line=res.read_line.astype(np.uint64), unitig_names=res.unitig_names, vm=vm,
                  edge_first=res.edge_first, edge_second=res.edge_second, edge_em_begin=res.edge_em_begin, em=em,
                  times={})


def check_invariants(res):
    """Size-independent properties of a construction result (used where the oracle is too slow to run)."""
    m, e, c = res.matches, res.edge_matches, res.counts
    M, E, EM = c["n_matches"], c["n_edges"], c["n_edge_matches"]
    assert np.all(np.diff(m["unitig"].astype(np.int64)) >= 0), "matches not grouped by unitig"
    same = np.diff(m["unitig"].astype(np.int64)) == 0
    assert np.all(np.diff(m["line"].astype(np.int64))[same] > 0), "matches of a unitig not in line order"
    # (unitig, read) unique
    key = m["unitig"].astype(np.uint64) << np.uint64(32) | m["read"].astype(np.uint64)
    assert np.unique(key).size == M
    # edges strictly increasing by (first, second), first < second
    ek = res.edge_first.astype(np.uint64) << np.uint64(32) | res.edge_second.astype(np.uint64)
    assert np.all(np.diff(ek.astype(np.int64)) > 0) if E > 1 else True
    assert np.all(res.edge_first < res.edge_second)
    b = res.edge_em_begin.astype(np.int64)
    assert b[0] == 0 and b[-1] == EM and np.all(np.diff(b) >= 1)
    # unitigs strictly increasing inside an edge
    edge_of = np.repeat(np.arange(E), np.diff(b))
    du = np.diff(e["unitig"].astype(np.int64))
    assert np.all(du[np.diff(edge_of) == 0] > 0)
    assert np.all(e["ov1"] - e["ov0"] > 100)
    # pair count = sum k(k-1)/2
    k = np.bincount(m["unitig"], minlength=c["n_unitigs"]).astype(np.int64)
    assert int((k * (k - 1) // 2).sum()) == c["n_pairs"]
    assert EM <= c["n_pairs"]


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def edge_case_pafs() -> dict:
    """Small hand-built PAFs that hit every rule of the path. Each ends with a sentinel line because
    the reference never parses the last line (BlastFileReader.cpp:76)."""
    L = paf_line
    sentinel = L("zz", 900, 0, 300, "+", "zr", 1000, 0, 300, 290)
    cases = {"kat8": kat8_paf()}
    # last line without a terminating newline is still "the last line" and dropped
    cases["no_trailing_newline"] = kat8_paf()[:-1]
    # one parsed line only / nothing accepted
    cases["single_match"] = L("u1", 5000, 0, 3000, "+", "r1", 20000, 100, 3100, 2900) + sentinel
    cases["nothing_accepted"] = L("u1", 5000, 0, 300, "+", "r1", 20000, 100, 400, 290) * 3 + sentinel
    # keep-min-line per (read, unitig) (MatchMap.cpp:64-73) + a read whose FIRST line is rejected (ids are assigned on
    # the first ACCEPTED line, BlastFileReader.cpp:109-111) + ungrouped unitigs
    cases["duplicates_ungrouped"] = b"".join([
        L("u1", 9000, 0, 300, "+", "rA", 30000, 0, 300, 290),        # rejected (too short): rA not yet known
        L("u2", 9000, 0, 4000, "+", "rB", 30000, 0, 4000, 3900),     # rB = read 0, u2 = unitig 0
        L("u1", 9000, 100, 5000, "-", "rA", 30000, 50, 4950, 4500),  # rA = read 1, u1 = unitig 1
        L("u2", 9000, 500, 4500, "+", "rA", 30000, 7000, 11000, 3800),
        L("u1", 9000, 200, 4000, "+", "rB", 30000, 9000, 12800, 3500),
        L("u1", 9000, 0, 9000, "+", "rA", 30000, 20000, 29000, 8800),  # duplicate (rA, u1): later line loses
        L("u2", 9000, 0, 9000, "-", "rB", 30000, 100, 9100, 8000),     # duplicate (rB, u2): later line loses
        L("u3", 2000, 0, 2000, "+", "rC", 5000, 0, 2000, 1900),
        L("u1", 9000, 4800, 9000, "+", "rC", 5000, 100, 4300, 4000),
        sentinel])
    # thresholds: matches 399/400/499/500, unitig span 399/400/499/500, overlap 100 vs 101 (MatchMap.cpp:192: > 100)
    rows = []
    for k, (span, matches) in enumerate([(399, 500), (400, 399), (400, 400), (499, 500), (500, 499), (500, 500),
                                         (2000, 1900)]):
        rows.append(L("t1", 9000, 1000, 1000 + span, "+", f"q{k}", 20000, 10, 10 + span, matches))
    rows.append(L("t2", 9000, 0, 1000, "+", "qa", 20000, 0, 1000, 900))
    rows.append(L("t2", 9000, 899, 2000, "+", "qb", 20000, 0, 1101, 1000))   # ov = [899, 999]: 100 -> rejected
    rows.append(L("t2", 9000, 898, 2000, "-", "qc", 20000, 0, 1102, 1000))   # ov = [898, 999]: 101 -> kept
    cases["thresholds"] = b"".join(rows) + sentinel
    # std::stoi tolerance: leading blanks, '+', trailing junk; CRLF; long tag fields; strand not '+'
    cases["stoi_crlf_tags"] = b"".join([
        b"u1\t5000\t 0\t+3000\t+\tr1\t20000\t100\t3100x\t2900\t3000\t60\tcg:Z:" + b"120=1X" * 300 + b"\r\n",
        b"u1\t5000\t1000\t4000\t*\tr2\t15000\t500\t3500\t2800\t3000\t60\ttp:A:P\r\n",
        b"u1\t5000\t2500\t5000\t+\tr3\t18000\t0\t2500\t450 \t2500\t60\r\n",
        b"u1\t5000\t2500\t5000\t-\tr4\t-5\t0\t2500\t-1\t2500\t60\r\n",   # negative matches widen to a huge size_t
        sentinel])
    # std::stoi at the edges of int: INT_MAX as a read length, -INT_MAX-1 as a match count, ten and more digits made
    # of leading zeros (the fast path of the device parser handles up to nine digits; these take the exact one)
    cases["stoi_ten_digits"] = b"".join([
        b"u1\t5000\t0000000000\t00000003000\t+\tr1\t2147483647\t100\t3100\t2900\t3000\t60\n",
        b"u1\t5000\t1000\t4000\t-\tr2\t0000000000015000\t500\t3500\t-2147483648\t3000\t60\n",
        b"u1\t5000\t2500\t5000\t+\tr3\t18000\t000000000\t0000002500\t450\t2500\t60\n",
        b"u1\t5000\t2400\t5000\t+\tr4\t999999999\t1\t2601\t1234567890\t2600\t60\n",
        sentinel])
    # one heavy unitig: k = 120 reads -> 7 140 pairs in one group, plus small groups around it
    rng = np.random.Generator(np.random.PCG64(7))
    rows = []
    for u in range(6):
        k = 120 if u == 3 else int(rng.integers(1, 9))
        ulen = 12000
        reads = rng.choice(160, k, replace=False)
        for r in reads:
            a = int(rng.integers(0, ulen - 1500))
            b = int(min(ulen, a + rng.integers(450, 6000)))
            rows.append(L(f"h{u}", ulen, a, b, "+-"[int(rng.integers(0, 2))], f"x{r}", 50000, 1000, 1000 + (b - a),
                          int((b - a) * 0.9)))
    cases["heavy_group"] = b"".join(rows) + sentinel
    # every line carries a long cg:Z tag (minimap2 -c --eqx, pipeline.sh:175): 256 consecutive lines no longer fit the
    # parser's shared-memory stage, so the lines are read in place; a few groups of overlapping reads
    rng = np.random.Generator(np.random.PCG64(9))
    rows = []
    for u in range(40):
        for r in rng.choice(60, int(rng.integers(2, 12)), replace=False):
            a = int(rng.integers(0, 3000))
            b = int(a + rng.integers(420, 5000))
            tag = ("cg:Z:" + "".join(f"{int(x)}=1X" for x in rng.integers(1, 99, 90))).encode()
            rows.append(L(f"c{u}", 9000, a, b, "+-"[int(rng.integers(0, 2))], f"y{r}", 40000, 50, 50 + (b - a),
                          int((b - a) * 0.95), extra=("tp:A:P", tag.decode())))
    cases["long_tags"] = b"".join(rows) + sentinel
    return cases


def invalid_pafs() -> dict:
    """Inputs the reference answers with an exception ("Invalid BLAST file." / std::stoi failures)."""
    L = paf_line
    sentinel = L("zz", 900, 0, 300, "+", "zr", 1000, 0, 300, 290)
    good = L("u1", 5000, 0, 3000, "+", "r1", 20000, 100, 3100, 2900)
    return {
        "too_few_fields": good + b"u1\t5000\t0\t3000\t+\tr1\t20000\t100\t3100\n" + sentinel,
        "empty_line": good + b"\n" + sentinel,
        "not_a_number": good + b"u1\t5000\tabc\t3000\t+\tr1\t20000\t100\t3100\t2900\t3000\t60\n" + sentinel,
        "int_overflow_by_one": good + b"u1\t5000\t0\t3000\t+\tr1\t2147483648\t100\t3100\t2900\t3000\t60\n" + sentinel,
        "int_underflow_by_one": good + b"u1\t5000\t0\t3000\t+\tr1\t20000\t100\t3100\t-2147483649\t3000\t60\n" + sentinel,
        "int_overflow": good + b"u1\t5000\t0\t99999999999\t+\tr1\t20000\t100\t3100\t2900\t3000\t60\n" + sentinel,
        "bad_read_coord_on_accepted": good + b"u1\t5000\t0\t3000\t+\tr1\t20000\t\t3100\t2900\t3000\t60\n" + sentinel,
    }


def heavy_rows_paf(k0=3000, k1=2500, shared_from=1000, seed=11, small_groups=40) -> bytes:
    """Two high-multiplicity unitigs that share most of their reads (a repeat seen by thousands of reads), so that
    rows of the edge table hold far more candidates than one pairing tile (4096): exercises the value-range split of
    heavy rows, with duplicates (two edge matches per edge) on both sides of every split point.  Lines of the two
    unitigs are interleaved and the reads appear in shuffled order, so read ids are not monotone inside a group."""
    rng = np.random.Generator(np.random.PCG64(seed))
    L = paf_line
    rows = []
    ulen = 20000
    for u, (lo, k) in enumerate([(0, k0), (shared_from, k1)]):
        for r in rng.permutation(np.arange(lo, lo + k)):
            a = int(rng.integers(0, ulen - 3000))
            b = int(min(ulen, a + rng.integers(420, 15000)))
            rows.append((int(rng.integers(0, 1 << 30)),
                         L(f"H{u}", ulen, a, b, "+-"[int(rng.integers(0, 2))], f"x{r}", 60000, 500, 500 + (b - a),
                           int((b - a) * 0.93))))
    for g in range(small_groups):
        k = int(rng.integers(2, 30))
        for r in rng.choice(k0 + k1, k, replace=False):
            a = int(rng.integers(0, 4000))
            b = int(a + rng.integers(420, 5000))
            rows.append((int(rng.integers(0, 1 << 30)),
                         L(f"s{g}", 9000, a, b, "+", f"x{r}", 60000, 100, 100 + (b - a), int((b - a) * 0.9))))
    rows.sort(key=lambda t: t[0])
    sentinel = L("zz", 900, 0, 300, "+", "zr", 1000, 0, 300, 290)
    return b"".join(r[1] for r in rows) + sentinel


def group_sizes_paf(sizes, seed=23) -> bytes:
    """One unitig per entry of `sizes`, matched by that many reads drawn from a common pool (so the groups share reads
    and rows get several sources); the lines of all groups are shuffled, so read ids are not monotone inside a group."""
    rng = np.random.Generator(np.random.PCG64(seed))
    L = paf_line
    pool = max(sizes) + 50
    rows = []
    for g, k in enumerate(sizes):
        for r in rng.choice(pool, k, replace=False):
            a = int(rng.integers(0, 3000))
            b = int(a + rng.integers(420, 6000))
            rows.append((int(rng.integers(0, 1 << 30)),
                         L(f"g{g}", 9000, a, b, "+-"[int(rng.integers(0, 2))], f"x{r}", 60000, 100, 100 + (b - a),
                           int((b - a) * 0.9))))
    rows.sort(key=lambda t: t[0])
    sentinel = L("zz", 900, 0, 300, "+", "zr", 1000, 0, 300, 290)
    return b"".join(r[1] for r in rows) + sentinel


def merge_parts(parts):
    """Concatenate the per-rank results of a multi-GPU construction (rank order) into one Construction: the global
    tables are taken from rank 0 (they are replicated), the edge slices are appended."""
    import copy
    out = copy.copy(parts[0])
    out.edge_first = np.concatenate([p.edge_first for p in parts])
    out.edge_second = np.concatenate([p.edge_second for p in parts])
    begins, base = [], 0
    for p in parts:
        begins.append(p.edge_em_begin[:-1].astype(np.uint64) + np.uint64(base))
        base += int(p.edge_em_begin[-1])
    out.edge_em_begin = np.concatenate(begins + [np.array([base], np.uint64)])
    out.edge_matches = {k: np.concatenate([p.edge_matches[k] for p in parts]) for k in parts[0].edge_matches}
    out.counts = dict(parts[0].counts)
    out.counts["n_edges"] = int(out.edge_first.size)
    out.counts["n_edge_matches"] = base
    return out


from oracle.sampled import (checker_sample_edge_matches, gpu_sample_edge_matches, records_equal,  # noqa: E402,F401
                            subset_paf, unitig_in_sample)
