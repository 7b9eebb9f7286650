"""Records golden vectors for the pipeline's unitig filter from the UNMODIFIED script — build container only:

    python tests/golden/make_prefilter_golden.py      ->  tests/golden/unitig_filter.json

/root/reference/pipeline/unitig_filter.py is run whole (runpy) with sys.argv set and a stand-in for
`Bio.SeqIO.index_db` backed by a plain FASTA parser (Biopython is not installed here; the script only uses
`record.seq[a:b]`, `str(...)`, `len(...)` and `record.description`).  Recorded per fixture: the per-block statistics of
pass 1, the cut-offs, the fragments of pass 2, and the SHA-256 of the two files the script writes.  The restatement
oracle/unitig_filter_oracle.py is checked against the script on every fixture.
"""
import dataclasses
import hashlib
import json
import os
import runpy
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from muchsalsa_b200 import synth  # noqa: E402
from oracle import unitig_filter_oracle as ufo  # noqa: E402

SCRIPT = "/root/reference/pipeline/unitig_filter.py"


class _Record:
    def __init__(self, description, seq):
        self.description, self.seq = description, seq


def _index_db(_idx, path, _fmt):
    recs, name, desc, parts = {}, None, None, []
    for line in open(path):
        if line.startswith(">"):
            if name is not None:
                recs[name] = _Record(desc, "".join(parts))
            desc = line[1:].rstrip("\n")
            name, parts = desc.split()[0], []
        else:
            parts.append(line.strip())
    if name is not None:
        recs[name] = _Record(desc, "".join(parts))
    return recs


def run_script(paf: bytes, unitigs_fa: bytes):
    bio, seqio = types.ModuleType("Bio"), types.ModuleType("Bio.SeqIO")
    seqio.index_db = _index_db
    bio.SeqIO = seqio
    sys.modules["Bio"], sys.modules["Bio.SeqIO"] = bio, seqio
    with tempfile.TemporaryDirectory() as td:
        p, u, st, out = (os.path.join(td, n) for n in ("in.paf", "unitigs.fa", "report.txt", "unitigs.corrected.fa"))
        open(p, "wb").write(paf)
        open(u, "wb").write(unitigs_fa)
        argv = sys.argv
        sys.argv = [SCRIPT, p, u, st, out]
        try:
            ns = runpy.run_path(SCRIPT, run_name="unitig_filter")
        finally:
            sys.argv = argv
        return ns, open(out, "rb").read(), open(st, "rb").read()


def fixtures():
    """name -> (PAF, unitigs FASTA).  The repeat-rich ones have blocks above the cut-off."""
    out = {}
    for name, cfg in (("small", synth.CONFIGS["small"]),
                      ("small_repeats", dataclasses.replace(synth.CONFIGS["small"], rep_frac=0.08, rep_kind="geom",
                                                            rep_mean=6.0, rep_cap=40, coverage=12.0)),
                      ("small_shuffled", dataclasses.replace(synth.CONFIGS["small"], shuffle_groups=True, rep_frac=0.05,
                                                             rep_kind="geom", rep_mean=3.0, rep_cap=10))):
        out[name] = (synth.generate_paf(cfg), synth.generate_sequences(cfg)[0])
    # hand-built: unitigs whose coverage peaks are cut out and whose quiet stretches (>= 500 bp) are rescued; a unitig
    # with two blocks; a read that hits a block twice (counted once in pass 1, twice in pass 2)
    sys.path.insert(0, os.path.dirname(HERE))
    import util
    L = util.paf_line
    rng = np.random.Generator(np.random.PCG64(5))
    rows, fa = [], []
    for u in range(40):
        ulen = 6000
        fa.append(b">pk%d some description\n" % u + bytes(rng.choice(list(b"ACGT"), ulen).astype(np.uint8)) + b"\n")
        n_bg = int(rng.integers(2, 6))
        for r in range(n_bg):  # background coverage over the whole unitig
            rows.append(L(f"pk{u}", ulen, 0, ulen, "+", f"bg{u}_{r}", 20000, 0, ulen, 5000))
        if u % 4 == 0:  # a peak in the middle: 30 reads on [2000, 3000)
            for r in range(30):
                rows.append(L(f"pk{u}", ulen, 2000, 3000, "+", f"pile{u}_{r}", 20000, 0, 1000, 900))
            rows.append(L(f"pk{u}", ulen, 100, 400, "-", f"pile{u}_0", 20000, 5000, 5300, 250))  # same read again
    rows.append(L("pk3", 6000, 0, 6000, "+", "late", 20000, 0, 6000, 5000))  # a second block of pk3
    out["peaks"] = (b"".join(rows), b"".join(fa))
    return out


def main():
    res = {}
    for name, (paf, ufa) in fixtures().items():
        ns, out_bytes, stats_bytes = run_script(paf, ufa)
        blocks = ufo.blocks_pass1(paf)
        cutoff, q3, last = ufo.cutoffs(blocks)
        assert {u: v[2] for u, v in ns["mapping_freq"].items()} == last, name
        assert {u: (v[0], v[1]) for u, v in ns["mapping_freq"].items()} == \
            {u: (n, ln) for u, n, ln, _ in blocks}, name  # (last block wins in both)
        assert float(ns["cutoff"]) == cutoff and float(ns["q3"]) == q3, name
        p2 = ufo.blocks_pass2(paf, last, cutoff, q3)
        n_out = sum(1 for _, fr in p2 if fr is not None)
        assert n_out == ns["outlier_counter"] and len(p2) == ns["all_counter"], name
        res[name] = {"blocks": [[u, n, ln, mx] for u, n, ln, mx in blocks], "cutoff": cutoff, "q3": q3,
                     "pass2": [[u, None if fr is None else [list(x) for x in fr]] for u, fr in p2],
                     "outliers": n_out, "rescued": int(ns["rescued_outlier"]),
                     "output_sha256": hashlib.sha256(out_bytes).hexdigest(),
                     "stats_sha256": hashlib.sha256(stats_bytes).hexdigest()}
        print(name, len(blocks), "blocks", n_out, "outliers", ns["rescued_outlier"], "rescued", "cutoff", cutoff,
              "q3", q3, flush=True)
    json.dump(res, open(os.path.join(HERE, "unitig_filter.json"), "w"))


if __name__ == "__main__":
    main()
