"""Sampled parity — TEST INFRASTRUCTURE (used by tests/, bench.py's verification legs and scripts/run_big.py; never by
the product).

Configurations whose result is too large for the CPU checker (thaliana, human) are compared on a deterministic subset
S of the unitigs.  The edge matches a unitig contributes depend only on that unitig's group (MatchMap::processScaffold
works on one scaffold, MatchMap.cpp:175-224), so the checker run on the PAF lines of S alone must produce exactly the
GPU's edge matches whose unitig is in S — compared by NAME.  Which read of a pair is `first` depends on the global
first accepted line of the reads (MatchMap.cpp:204-213), which a subset cannot see: pairs are compared unordered, and
the GPU's own orientation is checked against its own read table.
"""
from __future__ import annotations

import numpy as np


def unitig_in_sample(name: bytes, modulus: int) -> bool:
    import zlib
    return zlib.crc32(name) % modulus == 0


def subset_paf(paf: bytes, modulus: int):
    """Lines of `paf` (all but its last, which the reader never parses) whose unitig is in the sample, + a sentinel.
    Returns (subset PAF, original line index of every subset line)."""
    lines = paf.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    keep, orig, memo = [], [], {}
    for i, ln in enumerate(lines[:-1]):
        nm = ln[:ln.find(b"\t")]
        hit = memo.get(nm)
        if hit is None:
            hit = memo[nm] = unitig_in_sample(nm, modulus)
        if hit:
            keep.append(ln)
            orig.append(i)
    keep.append(lines[-1] if lines else b"zz\t900\t0\t300\t+\tzr\t1000\t0\t300\t290\t300\t60")
    return b"\n".join(keep) + b"\n", np.asarray(orig, np.int64)


def _canon_em(first_names, second_names, unitig_names, ov0, ov1, score, line, flags):
    lo = np.where(first_names <= second_names, first_names, second_names)
    hi = np.where(first_names <= second_names, second_names, first_names)
    rec = np.rec.fromarrays([lo, hi, unitig_names, ov0.astype(np.int64), ov1.astype(np.int64),
                             np.ascontiguousarray(score, np.float64).view(np.uint64), line.astype(np.int64),
                             flags.astype(np.int64)], names="lo,hi,u,ov0,ov1,score,line,flags")
    return np.sort(rec, order=["lo", "hi", "u"])


def gpu_sample_edge_matches(gb, modulus: int):
    """The edge matches of the context's CURRENT device-resident edge tables whose unitig is in the sample, gathered
    on the device (torch ops over the library's buffers — nothing but the sample and the name tables is downloaded).
    Returns (canonical record array, orientation_ok)."""
    import torch
    from muchsalsa_b200.construct import device_view as dv
    c, t = gb.counts(), gb.device_tables()
    R, U, E, EM = c["n_reads"], c["n_unitigs"], c["n_edges"], c["n_edge_matches"]

    def names(off_ptr, blob_ptr, n):
        off = dv(off_ptr, n + 1, "u8").cpu().numpy()
        blob = dv(blob_ptr, int(off[-1]), "u1").cpu().numpy().tobytes()
        return np.array([blob[int(off[i]):int(off[i + 1])] for i in range(n)], dtype=object)

    un, rn = names(t.unitig_name_off, t.unitig_names, U), names(t.read_name_off, t.read_names, R)
    member = torch.from_numpy(np.array([unitig_in_sample(x, modulus) for x in un], np.bool_)).cuda()
    em_u = dv(t.em_unitig, EM, "u4")
    idx = torch.nonzero(member[em_u.long()]).squeeze(1)
    begin = dv(t.edge_em_begin, E + 1, "u8")
    edge = torch.searchsorted(begin, idx, right=True) - 1
    first, second = dv(t.edge_first, E, "u4")[edge].long(), dv(t.edge_second, E, "u4")[edge].long()
    rl = dv(t.read_line, R, "u4").long()
    orient_ok = bool((rl[first] < rl[second]).all().item()) and bool((first < second).all().item())
    g = lambda ptr, kind: dv(ptr, EM, kind)[idx].cpu().numpy()
    rec = _canon_em(rn[first.cpu().numpy()].astype("S"), rn[second.cpu().numpy()].astype("S"),
                    un[em_u[idx].long().cpu().numpy()].astype("S"), g(t.em_ov0, "i4"), g(t.em_ov1, "i4"),
                    g(t.em_score, "f8"), g(t.em_line, "u4"), g(t.em_flags, "u1"))
    return rec, orient_ok


def checker_sample_edge_matches(checker, paf: bytes, modulus: int):
    """The same canonical records from the CPU checker run on the sampled lines only."""
    sub, orig = subset_paf(paf, modulus)
    t = checker.build(sub)
    b = t.edge_em_begin.astype(np.int64)
    edge = np.repeat(np.arange(t.edge_first.size), np.diff(b))
    rn = np.array([x.encode("latin-1") for x in t.read_names], dtype="S") if t.read_names else np.zeros(0, "S1")
    un = np.array([x.encode("latin-1") for x in t.unitig_names], dtype="S") if t.unitig_names else np.zeros(0, "S1")
    em = t.em
    flags = (em["dir"].astype(np.int64) & 1) | ((em["prim"].astype(np.int64) & 1) << 1)
    return _canon_em(rn[t.edge_first[edge]], rn[t.edge_second[edge]], un[em["unitig"]], em["ov0"], em["ov1"],
                     em["score"], orig[em["line"].astype(np.int64)], flags)


def records_equal(a, b) -> bool:
    return a.shape == b.shape and all(np.array_equal(a[f], b[f]) for f in a.dtype.names)


# ---------------------------------------------------------------------------------------------------------------------
# the same for results that are never resident as a whole (chunked builds, multi-GPU slices): rows are gathered chunk
# by chunk as plain id columns; names are attached once at the end
# ---------------------------------------------------------------------------------------------------------------------
def device_name_table(off_ptr, blob_ptr, n):
    """A name table of the library (offsets + blob on the device) as a numpy bytes array indexed by id."""
    from muchsalsa_b200.construct import device_view as dv
    off = dv(off_ptr, n + 1, "u8").cpu().numpy()
    blob = dv(blob_ptr, int(off[-1]), "u1").cpu().numpy().tobytes() if n else b""
    return np.array([blob[int(off[i]):int(off[i + 1])] for i in range(n)], dtype="S") if n else np.zeros(0, "S1")


def sample_rows_of_tables(first_p, second_p, begin_p, n_edges, n_em, em_base, em_ptrs, member):
    """Rows of one edge-table piece (device pointers; `begin_p` holds slice-relative values starting at em_base)
    whose unitig is flagged in `member` (bool tensor indexed by unitig id).  Returns id columns as numpy arrays."""
    import torch
    from muchsalsa_b200.construct import device_view as dv
    em_u = dv(em_ptrs["unitig"], n_em, "u4")
    idx = torch.nonzero(member[em_u.long()]).squeeze(1)
    begin = dv(begin_p, n_edges + 1, "u8") - em_base
    edge = torch.searchsorted(begin, idx, right=True) - 1
    out = {"first": dv(first_p, n_edges, "u4")[edge].long().cpu().numpy(),
           "second": dv(second_p, n_edges, "u4")[edge].long().cpu().numpy(),
           "unitig": em_u[idx].long().cpu().numpy()}
    for k, kind in (("ov0", "i4"), ("ov1", "i4"), ("score", "f8"), ("line", "u4"), ("flags", "u1")):
        out[k] = dv(em_ptrs[k], n_em, kind)[idx].cpu().numpy()
    return out


def sample_rows_of_chunk(c, member):
    """sample_rows_of_tables for a msgpu_chunk (muchsalsa_b200._lib.Chunk)."""
    return sample_rows_of_tables(c.edge_first, c.edge_second, c.edge_em_begin, int(c.n_edges), int(c.n_edge_matches),
                                 int(c.em_base), dict(unitig=c.em_unitig, ov0=c.em_ov0, ov1=c.em_ov1, score=c.em_score,
                                                      line=c.em_line, flags=c.em_flags), member)


def concat_rows(parts):
    keys = ("first", "second", "unitig", "ov0", "ov1", "score", "line", "flags")
    if not parts:
        return {k: np.zeros(0, np.float64 if k == "score" else np.int64) for k in keys}
    return {k: np.concatenate([p[k] for p in parts]) for k in keys}


def canon_from_ids(rows, read_names, unitig_names):
    return _canon_em(read_names[rows["first"]], read_names[rows["second"]], unitig_names[rows["unitig"]], rows["ov0"],
                     rows["ov1"], rows["score"], rows["line"], rows["flags"])


def checker_records_for_lines(checker, lines, orig):
    """Canonical records of the checker run on `lines` (bytes objects without terminators, unitig-grouped or not) whose
    original global line indices are `orig`; a sentinel line is appended (the reader drops the last line)."""
    sub = b"\n".join(list(lines) + [b"zz\t900\t0\t300\t+\tzr\t1000\t0\t300\t290\t300\t60"]) + b"\n"
    t = checker.build(sub)
    b = t.edge_em_begin.astype(np.int64)
    edge = np.repeat(np.arange(t.edge_first.size), np.diff(b))
    rn = np.array([x.encode("latin-1") for x in t.read_names], dtype="S") if t.read_names else np.zeros(0, "S1")
    un = np.array([x.encode("latin-1") for x in t.unitig_names], dtype="S") if t.unitig_names else np.zeros(0, "S1")
    em = t.em
    flags = (em["dir"].astype(np.int64) & 1) | ((em["prim"].astype(np.int64) & 1) << 1)
    orig = np.asarray(orig, np.int64)
    return _canon_em(rn[t.edge_first[edge]], rn[t.edge_second[edge]], un[em["unitig"]], em["ov0"], em["ov1"],
                     em["score"], orig[em["line"].astype(np.int64)], flags)
