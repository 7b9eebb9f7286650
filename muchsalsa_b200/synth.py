"""Deterministic synthetic unitig/nanopore PAF workloads.

There is no network for real genomes, so every named configuration of BASELINE.json is synthesised: unitigs tile a
genome of length G (length = max(500, lognormal(7.2, 0.6)), 500 = pipeline.sh:29 MINLENGTH), N = G*cov/9000 reads
(length = max(1000, Gamma(2, 4500)), uniform start, random strand), one PAF line per (unitig, read) whose geometric
overlap is >= 100 bp.  Columns follow the pipeline's minimap2 call (pipeline.sh:175: query = unitig, target = read) and
the positions BlastFileReader.cpp:52-60 reads.  Lines are grouped by unitig (minimap2 emits by query) and ONE sentinel
line is appended because BlastFileReader::read() never parses the last line (BlastFileReader.cpp:76).

Repeat model (configs 3-5): a fraction `rep_frac` of the unitigs is additionally placed at m extra genome loci
(m ~ Geometric or Zipf), which yields the high-multiplicity groups that stress the per-unitig k^2 pairing and produces
duplicate (read, unitig) lines that exercise the keep-min-line rule (MatchMap.cpp:64-80).

Host-side numpy only: this is workload generation, not the measured path.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

__all__ = ["SynthConfig", "CONFIGS", "generate_columns", "columns_to_paf", "generate_paf", "generate_sequences",
           "rank_chunk", "scaled"]


@dataclass(frozen=True)
class SynthConfig:
    name: str
    genome: int  # G, bp
    coverage: float
    seed: int
    rep_frac: float = 0.0  # fraction of unitigs with extra loci
    rep_kind: str = "none"  # "geom" | "zipf"
    rep_mean: float = 0.0  # geometric mean count / zipf exponent
    rep_cap: int = 0
    mean_read: int = 9000
    shuffle_groups: bool = False  # emit lines NOT grouped by unitig (exercises the grouping sort)
    extra: dict = field(default_factory=dict)


# The five named shapes of BASELINE.json `configs` (index = seed).
CONFIGS = {
    "small": SynthConfig("small", 1_500_000, 3.0, 1),
    "ecoli": SynthConfig("ecoli", 4_600_000, 30.0, 2),
    "yeast": SynthConfig("yeast", 12_000_000, 50.0, 3, rep_frac=0.05, rep_kind="geom", rep_mean=5.0, rep_cap=64),
    # thaliana: few unitigs, but with up to 200 extra loci each (k up to ~9e3 reads on one unitig): ~1e9 candidate
    # pairs, which is what one B200 holds as a device-resident result (25 B per edge match + 16 B per edge)
    "thaliana": SynthConfig("thaliana", 135_000_000, 40.0, 4, rep_frac=0.005, rep_kind="zipf", rep_mean=1.5,
                            rep_cap=200),
    # human: 3.1 Gbp / 30x: 2.07e6 unitigs, 1.03e7 reads, ~7e7 lines without repeats.  27 % of the unitigs recur at
    # Geometric(mean 12) further loci: 2.8e8 lines, 9e10 candidate pairs (groups of up to ~4 000 reads) — 2.3 TB of
    # edge tables, which is why this configuration exists only sharded over ranks and chunk by chunk
    # (synth_gpu.generate_columns_torch(cfg, shard=(rank, world)); never materialised whole, never on the host).
    # The repeat content is bounded by what eight B200s pair in about a second per step, not by biology: pairs grow
    # with the SQUARE of the multiplicity (10^9 lines from repeats of this kind would be > 10^12 pairs).
    "human": SynthConfig("human", 3_100_000_000, 30.0, 5, rep_frac=0.27, rep_kind="geom", rep_mean=12.0, rep_cap=400),
}


def _expand_ranges(lo: np.ndarray, hi: np.ndarray):
    """For every i, the integers lo[i]..hi[i]-1; returns (owner index, value)."""
    cnt = np.maximum(hi - lo, 0)
    tot = int(cnt.sum())
    owner = np.repeat(np.arange(lo.size, dtype=np.int64), cnt)
    start = np.repeat(lo, cnt)
    first = np.repeat(np.cumsum(cnt) - cnt, cnt)
    return owner, start + (np.arange(tot, dtype=np.int64) - first)


def generate_columns(cfg: SynthConfig, shard: tuple[int, int] | None = None, sentinel: bool = True) -> dict:
    """Integer PAF columns of the workload, in file order (grouped by unitig), sentinel line included.

    shard=(r, n): keep only the unitigs of the r-th of n contiguous genome slices (multi-GPU input sharding happens at
    unitig boundaries); ids stay global so shards can be compared with the unsharded run.
    sentinel=False omits the trailing sentinel line: the chunks of ONE file that is split over ranks carry it only on
    the last chunk (only the last line of the whole file is never parsed).
    """
    rng = np.random.Generator(np.random.PCG64(cfg.seed))
    G = cfg.genome

    # ---- unitigs tile the genome
    n_guess = int(G / 1500) + 64
    ulen = np.maximum(500, rng.lognormal(7.2, 0.6, n_guess).astype(np.int64))
    while ulen.sum() < G:
        ulen = np.concatenate([ulen, np.maximum(500, rng.lognormal(7.2, 0.6, n_guess).astype(np.int64))])
    ustart = np.cumsum(ulen) - ulen
    n_u = int(np.searchsorted(ustart, G, side="left"))
    ulen, ustart = ulen[:n_u].copy(), ustart[:n_u]
    ulen[-1] = max(500, G - ustart[-1])
    uend = ustart + ulen

    # ---- reads
    n_r = max(2, int(G * cfg.coverage / cfg.mean_read))
    rlen = np.maximum(1000, rng.gamma(2.0, cfg.mean_read / 2.0, n_r).astype(np.int64))
    rstart = rng.integers(0, max(1, G - 1000), n_r, dtype=np.int64)
    rlen = np.minimum(rlen, G + 500 - rstart)  # allow a little overhang past the genome end
    rend = rstart + rlen
    rstrand = rng.integers(0, 2, n_r, dtype=np.int64)  # 1 = '+'
    rname = rng.permutation(n_r)  # read names are not spatially ordered

    # ---- loci: (unitig, genome start).  Locus 0 of each unitig is its tile; repeats add more.
    loc_u = np.arange(n_u, dtype=np.int64)
    loc_s = ustart.copy()
    if cfg.rep_frac > 0:
        n_rep = int(n_u * cfg.rep_frac)
        rep_u = rng.choice(n_u, n_rep, replace=False)
        if cfg.rep_kind == "geom":
            m = rng.geometric(1.0 / cfg.rep_mean, n_rep)
        else:
            m = rng.zipf(cfg.rep_mean, n_rep)
        m = np.minimum(m, cfg.rep_cap).astype(np.int64)
        extra_u = np.repeat(rep_u, m)
        extra_s = rng.integers(0, np.maximum(1, G - ulen[extra_u]), extra_u.size, dtype=np.int64)
        loc_u = np.concatenate([loc_u, extra_u])
        loc_s = np.concatenate([loc_s, extra_s])
    loc_e = loc_s + ulen[loc_u]

    if shard is not None:
        r, n = shard
        lo_u, hi_u = (n_u * r) // n, (n_u * (r + 1)) // n
        keep = (loc_u >= lo_u) & (loc_u < hi_u)
        loc_u, loc_s, loc_e = loc_u[keep], loc_s[keep], loc_e[keep]

    # ---- (locus, read) pairs with >= 100 bp geometric overlap
    order = np.argsort(rstart, kind="stable")
    rs_sorted = rstart[order]
    max_rlen = int(rlen.max())
    # candidate reads for a locus start in (loc_s - max_rlen, loc_e - 100]
    lo = np.searchsorted(rs_sorted, loc_s - max_rlen, side="left")
    hi = np.searchsorted(rs_sorted, loc_e - 100, side="right")
    own, pos = _expand_ranges(lo, hi)
    rid = order[pos]
    a = np.maximum(loc_s[own], rstart[rid])
    b = np.minimum(loc_e[own], rend[rid])
    ok = (b - a) >= 100
    own, rid, a, b = own[ok], rid[ok], a[ok], b[ok]
    uid = loc_u[own]

    n = uid.size
    c2 = a - loc_s[own]  # unitig start of the aligned block
    c3 = b - loc_s[own]  # unitig end (exclusive)
    span = b - a
    # read coordinates of the block (on the read's own strand), with a small indel jitter
    off = a - rstart[rid]
    c7 = np.where(rstrand[rid] == 1, off, rlen[rid] - (off + span))
    jit = rng.integers(-20, 21, n, dtype=np.int64)
    c8 = np.clip(c7 + span + jit, 1, rlen[rid])
    c7 = np.clip(c7, 0, c8 - 1)
    c9 = np.floor(span * rng.uniform(0.85, 0.98, n)).astype(np.int64)
    c10 = np.maximum(span, c8 - c7)

    # ---- file order: grouped by unitig (minimap2 emits by query), reads by name inside a group
    key = np.lexsort((rname[rid], uid))
    if cfg.shuffle_groups:
        key = rng.permutation(n)
    cols = {
        "unitig": uid[key], "ulen": ulen[uid[key]], "c2": c2[key], "c3": c3[key], "strand": rstrand[rid[key]],
        "read": rname[rid[key]], "rlen": rlen[rid[key]], "c7": c7[key], "c8": c8[key], "c9": c9[key],
        "c10": c10[key],
    }
    # sentinel (dropped by the reader): a copy of the first line, or a dummy when the shard is empty
    for k in cols:
        tail = cols[k][:1] if n else np.array([1000 if k in ("ulen", "rlen", "c3", "c8", "c10") else 0], np.int64)
        cols[k] = np.concatenate([cols[k], tail] if sentinel else [cols[k]]).astype(np.int64)
    cols["n_unitigs"] = n_u
    cols["n_reads"] = n_r
    # genome geometry, for generate_sequences
    cols["_geom"] = dict(ulen=ulen, ustart=ustart, rlen=rlen, rstart=rstart, rstrand=rstrand, rname=rname, G=G)
    return cols


def _fmt_int_col(a: np.ndarray) -> np.ndarray:
    return np.char.mod("%d", a)


def columns_to_paf(cols: dict, unitig_prefix: str = "u", read_prefix: str = "r") -> bytes:
    """12-column PAF text of `generate_columns` output."""
    n = cols["unitig"].size
    if n == 0:
        return b""
    parts = [
        np.char.add(unitig_prefix, _fmt_int_col(cols["unitig"])), _fmt_int_col(cols["ulen"]),
        _fmt_int_col(cols["c2"]), _fmt_int_col(cols["c3"]), np.where(cols["strand"] == 1, "+", "-"),
        np.char.add(read_prefix, _fmt_int_col(cols["read"])), _fmt_int_col(cols["rlen"]), _fmt_int_col(cols["c7"]),
        _fmt_int_col(cols["c8"]), _fmt_int_col(cols["c9"]), _fmt_int_col(cols["c10"]),
    ]
    line = parts[0]
    for p in parts[1:]:
        line = np.char.add(np.char.add(line, "\t"), p)
    line = np.char.add(line, "\t60\n")
    return "".join(line.tolist()).encode("ascii")


def generate_paf(cfg: SynthConfig | str, shard: tuple[int, int] | None = None, sentinel: bool = True) -> bytes:
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    return columns_to_paf(generate_columns(cfg, shard, sentinel))


def scaled(cfg: SynthConfig | str, factor: int) -> SynthConfig:
    """The same shape on a `factor` times longer genome (weak scaling: one base-size share per GPU)."""
    import dataclasses
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    return cfg if factor == 1 else dataclasses.replace(cfg, name=f"{cfg.name}x{factor}", genome=cfg.genome * factor)


def rank_chunk(workload: str, rank: int, world: int) -> bytes:
    """Chunk `rank` of the `world`-rank weak-scaling workload: the genome is `world` times longer and rank r holds the
    unitigs of the r-th contiguous genome slice; only the last chunk carries the sentinel line.  The chunks
    concatenated in rank order are ONE PAF (the one whose digest tests/golden/sum_digests.json records); for
    world == 1 this is generate_paf(workload)."""
    if world == 1:
        return generate_paf(workload)
    return columns_to_paf(generate_columns(scaled(workload, world), (rank, world), sentinel=(rank == world - 1)))


_COMP = np.arange(256, dtype=np.uint8)
for _a, _b in (("A", "T"), ("C", "G")):
    _COMP[ord(_a)], _COMP[ord(_b)] = ord(_b), ord(_a)


def generate_sequences(cfg: SynthConfig | str, n_frac: float = 0.002, wrap: int = 70, fastq: bool = False):
    """(unitigs FASTA bytes, nanopore FASTA/FASTQ bytes) consistent with generate_columns(cfg) for SMALL configs.

    The genome is random ACGT with a sprinkle of N / lower-case / IUPAC bytes: the reference complements only upper
    case ACGT and passes everything else through (SequenceUtils.cpp:46-57).  Records are wrapped at `wrap` columns so
    the whitespace stripping of SequenceAccessor.cpp:64-66 is exercised.
    """
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    cols = generate_columns(cfg)
    g = cols["_geom"]
    rng = np.random.Generator(np.random.PCG64(cfg.seed + 1000))
    genome = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, g["G"] + 600)].copy()
    odd = rng.random(genome.size) < n_frac
    genome[odd] = np.frombuffer(b"NnacgtRYKM", np.uint8)[rng.integers(0, 10, int(odd.sum()))]

    def wrap_seq(s: bytes) -> bytes:
        return b"\n".join(s[i:i + wrap] for i in range(0, len(s), wrap)) + b"\n"

    fa = bytearray()
    for i in range(g["ulen"].size):
        s = genome[g["ustart"][i]:g["ustart"][i] + g["ulen"][i]].tobytes()
        fa += b">u%d len=%d\n" % (i, len(s)) + wrap_seq(s)
    fx = bytearray()
    for i in range(g["rlen"].size):
        s = genome[g["rstart"][i]:g["rstart"][i] + g["rlen"][i]]
        if g["rstrand"][i] == 0:
            s = _COMP[s[::-1]]
        s = s.tobytes()
        if fastq:
            fx += b"@r%d strand=%d\n" % (g["rname"][i], g["rstrand"][i]) + s + b"\n+\n" + b"F" * len(s) + b"\n"
        else:
            fx += b">r%d strand=%d\n" % (g["rname"][i], g["rstrand"][i]) + wrap_seq(s)
    return bytes(fa), bytes(fx)
