"""The synthetic workload generator of muchsalsa_b200.synth, restated with torch ops so that the big configurations
(thaliana, human: 10^7 .. 10^8 PAF lines per rank) are generated and formatted ON THE GPU in seconds instead of
minutes of numpy string work on the host.  Workload generation only — not the measured path, not the product.

Same model as synth.generate_columns: unitigs tile the genome, reads are uniform, a fraction of the
unitigs has extra loci, one PAF line per (locus, read) with >= 100 bp of geometric overlap, lines grouped by unitig,
reads by name inside a group.  The random streams are torch's (Philox on CUDA, mt19937 on CPU), so the numbers differ
from the numpy generator's; a given (config, shard, device type) is deterministic.  The genome-wide draws (unitig
lengths, reads, repeat loci) are made on the CPU generator so that every rank — whatever its device — sees the SAME
genome; only the per-line jitter is drawn on the device.
"""
from __future__ import annotations

import numpy as np
import torch

from .synth import SynthConfig


def _genome(cfg: SynthConfig):
    """Genome-wide tables, identical on every rank: numpy PCG64 exactly like synth.generate_columns."""
    rng = np.random.Generator(np.random.PCG64(cfg.seed))
    G = cfg.genome
    n_guess = int(G / 1500) + 64
    ulen = np.maximum(500, rng.lognormal(7.2, 0.6, n_guess).astype(np.int64))
    while ulen.sum() < G:
        ulen = np.concatenate([ulen, np.maximum(500, rng.lognormal(7.2, 0.6, n_guess).astype(np.int64))])
    ustart = np.cumsum(ulen) - ulen
    n_u = int(np.searchsorted(ustart, G, side="left"))
    ulen, ustart = ulen[:n_u].copy(), ustart[:n_u]
    ulen[-1] = max(500, G - ustart[-1])
    n_r = max(2, int(G * cfg.coverage / cfg.mean_read))
    rlen = np.maximum(1000, rng.gamma(2.0, cfg.mean_read / 2.0, n_r).astype(np.int64))
    rstart = rng.integers(0, max(1, G - 1000), n_r, dtype=np.int64)
    rlen = np.minimum(rlen, G + 500 - rstart)
    rstrand = rng.integers(0, 2, n_r, dtype=np.int64)
    rname = rng.permutation(n_r)
    loc_u = np.arange(n_u, dtype=np.int64)
    loc_s = ustart.copy()
    if cfg.rep_frac > 0:
        n_rep = int(n_u * cfg.rep_frac)
        rep_u = rng.choice(n_u, n_rep, replace=False)
        m = rng.geometric(1.0 / cfg.rep_mean, n_rep) if cfg.rep_kind == "geom" else rng.zipf(cfg.rep_mean, n_rep)
        m = np.minimum(m, cfg.rep_cap).astype(np.int64)
        extra_u = np.repeat(rep_u, m)
        extra_s = rng.integers(0, np.maximum(1, G - ulen[extra_u]), extra_u.size, dtype=np.int64)
        loc_u = np.concatenate([loc_u, extra_u])
        loc_s = np.concatenate([loc_s, extra_s])
    return dict(G=G, n_u=n_u, n_r=n_r, ulen=ulen, ustart=ustart, rlen=rlen, rstart=rstart, rstrand=rstrand,
                rname=rname, loc_u=loc_u, loc_s=loc_s)


def generate_columns_torch(cfg: SynthConfig, shard=None, sentinel=True, device="cuda") -> dict:
    """Integer PAF columns (int64 tensors on `device`), in file order, sentinel line included.  shard=(r, n): the
    unitigs of the r-th of n contiguous genome slices."""
    g = _genome(cfg)
    dev = torch.device(device)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    ulen, rlen, rstart, rstrand, rname = t(g["ulen"]), t(g["rlen"]), t(g["rstart"]), t(g["rstrand"]), t(g["rname"])
    loc_u, loc_s = g["loc_u"], g["loc_s"]
    if shard is not None:
        r, n = shard
        lo_u, hi_u = (g["n_u"] * r) // n, (g["n_u"] * (r + 1)) // n
        keep = (loc_u >= lo_u) & (loc_u < hi_u)
        loc_u, loc_s = loc_u[keep], loc_s[keep]
    loc_u, loc_s = t(loc_u), t(loc_s)
    loc_e = loc_s + ulen[loc_u]
    rend = rstart + rlen
    order = torch.argsort(rstart, stable=True)
    rs_sorted = rstart[order]
    max_rlen = int(rlen.max().item())
    lo = torch.searchsorted(rs_sorted, loc_s - max_rlen, right=False)
    hi = torch.searchsorted(rs_sorted, loc_e - 100, right=True)
    cnt = (hi - lo).clamp_min(0)
    own = torch.repeat_interleave(torch.arange(cnt.numel(), device=dev), cnt)
    first = torch.cumsum(cnt, 0) - cnt
    pos = lo[own] + (torch.arange(own.numel(), device=dev) - first[own])
    rid = order[pos]
    del pos, first, cnt, lo, hi
    a = torch.maximum(loc_s[own], rstart[rid])
    b = torch.minimum(loc_e[own], rend[rid])
    ok = (b - a) >= 100
    own, rid, a, b = own[ok], rid[ok], a[ok], b[ok]
    del ok
    uid = loc_u[own]
    n_lines = int(uid.numel())
    c2 = a - loc_s[own]
    c3 = b - loc_s[own]
    span = b - a
    off = a - rstart[rid]
    del own
    c7 = torch.where(rstrand[rid] == 1, off, rlen[rid] - (off + span))
    gen = torch.Generator(device=dev)
    gen.manual_seed(cfg.seed * 1000003 + (shard[0] if shard else 0))
    jit = torch.randint(-20, 21, (n_lines,), generator=gen, device=dev)
    c8 = torch.minimum(torch.maximum(c7 + span + jit, torch.ones_like(c7)), rlen[rid])
    c7 = torch.minimum(c7.clamp_min(0), c8 - 1)
    frac = 0.85 + 0.13 * torch.rand((n_lines,), generator=gen, device=dev, dtype=torch.float64)
    c9 = torch.floor(span.double() * frac).long()
    c10 = torch.maximum(span, c8 - c7)
    # file order: grouped by unitig, reads by name inside a group
    key = torch.argsort(uid * (g["n_r"] + 1) + rname[rid], stable=True)
    rid = rid[key]
    cols = {"unitig": uid[key], "c2": c2[key], "c3": c3[key], "c7": c7[key], "c8": c8[key], "c9": c9[key],
            "c10": c10[key], "strand": rstrand[rid], "read": rname[rid], "rlen": rlen[rid]}
    cols["ulen"] = ulen[cols["unitig"]]
    if sentinel:  # a copy of the first line (or a dummy when the shard is empty): the reader never parses it
        for k in list(cols):
            tail = cols[k][:1] if n_lines else torch.tensor(
                [1000 if k in ("ulen", "rlen", "c3", "c8", "c10") else 0], dtype=torch.int64, device=dev)
            cols[k] = torch.cat([cols[k], tail])
    cols["n_unitigs"], cols["n_reads"] = g["n_u"], g["n_r"]
    return cols


_POW10 = [10 ** k for k in range(19)]


def _n_digits(v: torch.Tensor) -> torch.Tensor:
    n = torch.ones_like(v)
    for k in range(1, 19):
        n += (v >= _POW10[k]).long()
    return n


def columns_to_paf_torch(cols: dict, unitig_prefix: str = "u", read_prefix: str = "r") -> torch.Tensor:
    """The 12-column PAF text of the columns as one uint8 tensor on their device — byte for byte what
    synth.columns_to_paf writes for the same columns (all values are non-negative)."""
    dev = cols["unitig"].device
    n = int(cols["unitig"].numel())
    if n == 0:
        return torch.zeros(0, dtype=torch.uint8, device=dev)
    # fields: (kind, payload)  kind: "int" tensor | "lit" bytes | "strand"
    fields = [("lit", unitig_prefix.encode()), ("int", cols["unitig"]), ("lit", b"\t"), ("int", cols["ulen"]),
              ("lit", b"\t"), ("int", cols["c2"]), ("lit", b"\t"), ("int", cols["c3"]), ("lit", b"\t"),
              ("strand", cols["strand"]), ("lit", b"\t" + read_prefix.encode()), ("int", cols["read"]), ("lit", b"\t"),
              ("int", cols["rlen"]), ("lit", b"\t"), ("int", cols["c7"]), ("lit", b"\t"), ("int", cols["c8"]),
              ("lit", b"\t"), ("int", cols["c9"]), ("lit", b"\t"), ("int", cols["c10"]), ("lit", b"\t60\n")]
    width = torch.zeros(n, dtype=torch.int64, device=dev)
    digs = {}
    for i, (kind, v) in enumerate(fields):
        if kind == "int":
            digs[i] = _n_digits(v)
            width += digs[i]
        elif kind == "lit":
            width += len(v)
        else:
            width += 1
    start = torch.cumsum(width, 0) - width
    out = torch.empty(int(width.sum().item()), dtype=torch.uint8, device=dev)
    cur = start.clone()
    for i, (kind, v) in enumerate(fields):
        if kind == "lit":
            for j, ch in enumerate(v):
                out[cur + j] = ch
            cur += len(v)
        elif kind == "strand":
            out[cur] = torch.where(v == 1, torch.full_like(v, ord("+")), torch.full_like(v, ord("-"))).to(torch.uint8)
            cur += 1
        else:
            nd = digs[i]
            for d in range(int(nd.max().item())):  # digit d counted from the least significant one
                has = nd > d
                idx = (cur + nd - 1 - d)[has]
                out[idx] = ((v[has] // _POW10[d]) % 10 + ord("0")).to(torch.uint8)
            cur += nd
    return out
