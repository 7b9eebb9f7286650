"""Regenerates the golden fixtures of tests/golden/ from the REFERENCE ITSELF (oracle/_ref/libmsref.so, compiled from
/root/reference by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py [--yeast]

Outputs (committed):
    cases.json     full canonical tables of the reference for every hand-built PAF of tests/util.edge_case_pafs()
    digests.json   sha256 of the canonical tables (oracle.pyoracle.tables_digest) + row counts for the synthetic
                   configurations (small, small/shuffled, ecoli, ecoli/shuffled, yeast with --yeast)
    seq.json       what the reference's SequenceAccessor + get*Sequence return for the fixtures below
    ref_fasta.fa / ref_fastq.fq   byte copies of the reference's own sequence fixtures (test_data/, data not code);
                   seq.json is additionally checked against the strings libms/tests/SA_test.cpp expects
"""
import dataclasses
import json
import os
import re
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from muchsalsa_b200 import synth  # noqa: E402
from oracle.pyoracle import Reference, tables_digest  # noqa: E402
import util  # noqa: E402


def tables_json(t):
    d = dict(read_names=t.read_names, read_nlen=t.read_nlen.tolist(), read_line=t.read_line.tolist(),
             unitig_names=t.unitig_names, edge_first=t.edge_first.tolist(), edge_second=t.edge_second.tolist(),
             edge_em_begin=t.edge_em_begin.tolist())
    for grp in ("vm", "em"):
        g = getattr(t, grp)
        d[grp] = {k: (v.view(np.uint64).tolist() if v.dtype == np.float64 else v.tolist()) for k, v in g.items()}
    return d


def seq_cases():
    """(label, unitigs FASTA bytes, nanopore bytes, fastq?) — the reference fixtures + synthetic ones with wrapped
    lines, blank lines, CRLF, N / lower case / IUPAC bytes, duplicate ids, '@'-leading quality, no final newline."""
    fa = open(os.path.join(HERE, "ref_fasta.fa"), "rb").read()
    fq = open(os.path.join(HERE, "ref_fastq.fq"), "rb").read()
    odd_fa = (b"; comment before the first record\n>s1 first record\nACGTNacgtRY\nGGCCTTAA\n\n>s2\tsecond\r\n"
              b"AC GT\r\nTTNN\r\n>s1 duplicate id is ignored\nAAAA\n>s3\n>s4 empty one before\nG\n>s5 no newline at end\n"
              b"ACGTACGTAC")
    odd_fq = (b"@q1 desc\nACGTN\nACGT\n+\nFFFFF\nFFFF\n@q2\nGGCCA\n+q2\n@FFFF\n@q3\nTTAACC\n+\nFFFFFF")
    u, n = synth.generate_sequences("small")
    return [("ref_fasta_both", fa, fa, False), ("ref_fasta_fastq", fa, fq, True), ("odd", odd_fa, odd_fq, True),
            ("odd_fasta_reads", odd_fa, odd_fa, False), ("synth_small", u[:24000], n[:60000], False)]


def orders_digest(o) -> str:
    import hashlib
    h = hashlib.sha256()
    for f in ("edge_shadow", "edge_order_begin", "start_is_first", "left", "right", "contained", "direction", "primary",
              "score", "ids_begin", "ids"):
        h.update(np.ascontiguousarray(getattr(o, f)).tobytes())
    return h.hexdigest()


def filters_digest(f) -> str:
    import hashlib
    from oracle.pyoracle import FILTER_FIELDS
    h = hashlib.sha256()
    for k in FILTER_FIELDS:
        h.update(np.ascontiguousarray(f[k]).tobytes())
    return h.hexdigest()


def main():
    R = Reference()
    # the reference's own fixtures (data, not code) travel with the repo so the GPU box can load them
    for src, dst in (("fasta.fa", "ref_fasta.fa"), ("fastq.fq", "ref_fastq.fq")):
        shutil.copyfile(os.path.join("/root/reference/test_data", src), os.path.join(HERE, dst))
        os.chmod(os.path.join(HERE, dst), 0o644)

    cases = {name: tables_json(R.build(paf)) for name, paf in util.edge_case_pafs().items()}
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(cases, f)

    # known-answer vector of the per-edge jobs after chaining on the SURVEY.md §8c PAF
    fl = R.filters(util.kat8_paf())
    with open(os.path.join(HERE, "filters_kat8.json"), "w") as f:
        json.dump({k: fl[k].tolist() for k in ("contraction_order", "deletable", "n_kept_orders", "weight", "consensus")}, f)

    digests = {}
    names = ["small", "ecoli"] + (["yeast"] if "--yeast" in sys.argv else [])
    old = {}
    if os.path.exists(os.path.join(HERE, "digests.json")):
        old = json.load(open(os.path.join(HERE, "digests.json")))
    for name in names:
        for shuffled in (False, True):
            if name == "yeast" and shuffled:
                continue
            cfg = dataclasses.replace(synth.CONFIGS[name], shuffle_groups=shuffled)
            t = R.build(synth.generate_paf(cfg))
            digests[name + ("_shuffled" if shuffled else "")] = dict(digest=tables_digest(t), counts=t.counts)
            if not shuffled:  # the stages after construction, same inputs (yeast: ~10 min of the reference's CPU time)
                o = R.chain(synth.generate_paf(cfg))
                digests[name].update(orders_digest=orders_digest(o), orders=int(o.left.size),
                                     shadow_edges=int(o.edge_shadow.sum()))
                fl = R.filters(synth.generate_paf(cfg))  # the per-edge jobs after chaining
                digests[name].update(filters_digest=filters_digest(fl),
                                     contraction_edges=int((fl["contraction_order"] >= 0).sum()),
                                     deletable_edges=int(fl["deletable"].sum()))
            print(name, shuffled, t.counts, flush=True)
    for k, v in old.items():
        digests.setdefault(k, v)
    with open(os.path.join(HERE, "digests.json"), "w") as f:
        json.dump(digests, f, indent=1, sort_keys=True)

    # ---- sequences
    sa_src = open("/root/reference/libms/tests/SA_test.cpp").read()
    expected_literals = set(re.findall(r'"([ACGTN]{40,})"', sa_src))
    seq = {}
    rng = np.random.Generator(np.random.PCG64(11))
    for label, ufa, nfx, fastq in seq_cases():
        with tempfile.TemporaryDirectory() as td:
            pu, pn = os.path.join(td, "u.fa"), os.path.join(td, "n.fq" if fastq else "n.fasta")
            open(pu, "wb").write(ufa)
            open(pn, "wb").write(nfx)
            h = R.seq_open(pn, pu)
            entry = {"fastq": fastq, "records": [[], []], "fetch": []}
            for which, data in ((0, ufa), (1, nfx)):
                marker = b"@" if (which and fastq) else b">"
                names = []
                for line in data.split(b"\n"):
                    if line.startswith(marker):
                        nm = line[1:].split()[0].decode() if line[1:].split() else ""
                        if nm not in names:
                            names.append(nm)
                for nm in names:
                    id_ = h.id(which, nm)
                    full = h.full(which, id_)
                    entry["records"][which].append([nm, id_, None if full is None else full.decode("latin-1")])
                    if full is None:
                        continue
                    n = len(full)
                    wins = [(0, n - 2), (0, 0), (3, 10), (n - 5, n + 7), (-4, -2), (5, 1), (n, n), (n + 1, n + 3),
                            (-n - 9, 2), (2, -3)]
                    wins += [(int(a), int(b)) for a, b in rng.integers(-n - 2, n + 3, (6, 2))]
                    for (l, r) in wins:
                        for direction in (True, False):
                            got = h.fetch(which, id_, l, r, direction)
                            entry["fetch"].append([which, id_, l, r, direction,
                                                   None if got is None else got.decode("latin-1")])
            h.close()
            seq[label] = entry
    # the strings SA_test.cpp:11-136 expects must be among what we recorded for the reference fixtures
    recorded = {r[2] for lab in ("ref_fasta_both", "ref_fasta_fastq") for w in (0, 1) for r in seq[lab]["records"][w]}
    missing = [s[:30] for s in expected_literals if not any(s in (x or "") for x in recorded)]
    assert not missing, f"SA_test.cpp expectations not reproduced: {missing}"
    with open(os.path.join(HERE, "seq.json"), "w") as f:
        json.dump(seq, f)
    # strSlice / getReverseComplement known answers
    s = b"ACGTNacgtRYGGCCTTAA"
    kat = {"text": s.decode(), "revcomp": R.revcomp(s).decode(), "slices": []}
    for i in range(-25, 26):
        for j in range(-25, 26):
            got = R.strslice(s, i, j)
            kat["slices"].append([i, j, None if got is None else got.decode()])
    with open(os.path.join(HERE, "slice_kat.json"), "w") as f:
        json.dump(kat, f)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
