"""GPU parity tests of the sequence path (K8): resident store + batched slice / reverse complement through the C ABI,
against the golden vectors recorded from the reference and against the oracle."""
import importlib.util as iu
import os

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def _seq_cases():
    spec = iu.spec_from_file_location("make_golden", os.path.join(util.GOLDEN, "make_golden.py"))
    mg = iu.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return {c[0]: c for c in mg.seq_cases()}


@pytest.mark.parametrize("label", ["ref_fasta_both", "ref_fasta_fastq", "odd", "odd_fasta_reads", "synth_small"])
def test_store_and_fetch_vs_reference_golden(builder, label):
    _, ufa, nfx, fastq = _seq_cases()[label]
    g = util.load_golden("seq.json")[label]
    builder.build_from_buffer(b"")  # a fresh (empty) registry, like the golden run
    builder.seq_load_buffers(ufa, nfx, fastq)
    reqs, want = [], []
    for which in (0, 1):
        for name, id_, full in g["records"][which]:
            got_id = builder.seq_id(which, name)
            if full is None:
                assert got_id is None  # the reference has no such record
                continue
            assert got_id == id_, (label, which, name)
            # the whole record = slice [0, len - 2 + 1] of get*Sequence
            reqs.append((which, id_, 0, max(len(full) - 2, 0), True))
            want.append(full.encode("latin-1") if len(full) >= 1 else b"")
    for which, id_, l, r, direction, exp in g["fetch"]:
        reqs.append((which, id_, l, r, direction))
        want.append(None if exp is None else exp.encode("latin-1"))
    got = builder.seq_fetch_batch(reqs)
    for q, a, b in zip(reqs, got, want):
        assert a == b, (label, q)


def test_unknown_id_and_empty_batch(builder):
    _, ufa, nfx, fastq = _seq_cases()["odd"]
    builder.build_from_buffer(b"")
    builder.seq_load_buffers(ufa, nfx, fastq)
    assert builder.seq_fetch_batch([]) == []
    assert builder.seq_fetch_batch([(0, 999, 0, 3, True), (1, 3, 0, 3, True)]) == [None, None]


def test_ids_continue_the_construction_numbering(builder, oracle):
    """Sequences named by the PAF keep the ids construction gave them; the others are appended in file order
    (the reference shares its Registry objects, src/main.cpp:161-162)."""
    from muchsalsa_b200 import synth
    paf = synth.generate_paf("small")
    ufa, nfa = synth.generate_sequences("small")
    res = builder.build_from_buffer(paf).result()
    builder.seq_load_buffers(ufa, nfa, False)
    for which, names in ((0, res.unitig_names), (1, res.read_names)):
        for id_ in (0, 1, len(names) // 2, len(names) - 1):
            assert builder.seq_id(which, names[id_]) == id_
    # every vertex match: the unitig slice and the read slice the assembler would fetch, against the oracle's
    # slice / revcomp applied to the oracle's own extraction of the record
    stores = []
    for data in (ufa, nfa):
        d = {}
        for name, off, length in oracle.seq_index(data, False):
            d.setdefault(name, oracle.seq_extract(data, off, length))
        stores.append(d)
    m = res.matches
    rng = np.random.Generator(np.random.PCG64(3))
    pick = rng.choice(m["read"].size, 300, replace=False)
    reqs, want = [], []
    for i in pick:
        u, r = int(m["unitig"][i]), int(m["read"][i])
        direction = bool(m["flags"][i] & 1)
        reqs.append((0, u, int(m["ir0"][i]), int(m["ir1"][i]), direction))
        want.append(oracle.seq_fetch(stores[0][res.unitig_names[u]], int(m["ir0"][i]), int(m["ir1"][i]), direction))
        reqs.append((1, r, int(m["nr0"][i]), int(m["nr1"][i]), True))
        want.append(oracle.seq_fetch(stores[1][res.read_names[r]], int(m["nr0"][i]), int(m["nr1"][i]), True))
    assert builder.seq_fetch_batch(reqs) == want


def test_slice_revcomp_on_device_text(builder, oracle):
    """msgpu_slice_revcomp_device: strSlice + getReverseComplement on caller text, all buffers in HBM."""
    import ctypes as C
    import torch
    from muchsalsa_b200 import _lib
    kat = util.load_golden("slice_kat.json")
    s = kat["text"].encode()
    cases = [(i, j, rev) for i, j, _ in kat["slices"] for rev in (0, 1)]
    n = len(cases)
    text = torch.frombuffer(bytearray(s * n), dtype=torch.uint8).cuda()
    off = (torch.arange(n + 1, dtype=torch.int64) * len(s)).cuda()
    left = torch.tensor([c[0] for c in cases], dtype=torch.int32).cuda()
    right = torch.tensor([c[1] for c in cases], dtype=torch.int32).cuda()
    rev = torch.tensor([c[2] for c in cases], dtype=torch.uint8).cuda()
    out = torch.zeros(n * len(s) + 16, dtype=torch.uint8, device="cuda")
    out_off = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    status = torch.zeros(n, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    lib = _lib.load()
    rc = lib.msgpu_slice_revcomp_device(builder.handle, text.data_ptr(), off.data_ptr(), left.data_ptr(),
                                        right.data_ptr(), rev.data_ptr(), n, out.data_ptr(), out.numel(),
                                        out_off.data_ptr(), status.data_ptr())
    assert rc == 0, lib.msgpu_last_error(builder.handle)
    torch.cuda.synchronize()
    o, oo, st = out.cpu().numpy().tobytes(), out_off.cpu().numpy(), status.cpu().numpy()
    for k, (i, j, r) in enumerate(cases):
        want = oracle.strslice(s, i, j)
        if want is None:
            assert st[k] == 1
            continue
        if r:
            want = oracle.revcomp(want)
        assert st[k] == 0 and o[oo[k]:oo[k + 1]] == want, (i, j, r)
