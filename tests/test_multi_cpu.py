"""CPU tests of the multi-GPU host logic: PAF chunking and the two collectives of muchsalsa_b200.multi_gpu, run with two
processes over the gloo backend (the same code path NCCL takes on the GPU box)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from muchsalsa_b200 import _lib, multi_gpu, synth


def test_split_paf_is_a_partition_at_line_boundaries():
    paf = synth.generate_paf("small")
    for n in (1, 2, 3, 5, 8):
        chunks = multi_gpu.split_paf(paf, n)
        assert len(chunks) == n and b"".join(chunks) == paf
        assert all(c.endswith(b"\n") for c in chunks[:-1] if c)
    # more ranks than lines: trailing chunks are empty, nothing is lost
    tiny = b"a\tb\nc\td\n"
    chunks = multi_gpu.split_paf(tiny, 5)
    assert b"".join(chunks) == tiny and len(chunks) == 5
    assert multi_gpu.split_paf(b"", 3) == [b"", b"", b""]


def test_header_tensor_round_trip():
    h = _lib.ShardHeader()
    for i, (name, _) in enumerate(h._fields_):
        setattr(h, name, 0x3144524853534d47 if name == "magic" else (1 << 40) + i)
    h.magic = 0xF144524853534D47  # top bit set: must survive the int64 transport
    back = multi_gpu.tensor_to_headers(multi_gpu.header_to_tensor(h, "cpu").view(1, -1))[0]
    assert all(getattr(back, n) == getattr(h, n) for n, _ in h._fields_)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank owns a blob of a different size; sizes travel first, then the padded blobs
        n_local = 1000 + 777 * rank
        blob = torch.from_numpy(np.random.default_rng(rank).integers(0, 256, n_local, dtype=np.uint8))
        h = _lib.ShardHeader()
        h.magic, h.n_lines, h.n_reads, h.blob_bytes = 0x3144524853534D47, 10 + rank, 5 * rank, n_local
        heads = multi_gpu.gather_headers(multi_gpu.header_to_tensor(h, "cpu"), world, dist)
        headers = multi_gpu.tensor_to_headers(heads)
        stride = (max(int(x.blob_bytes) for x in headers) + 255) & ~255
        mine = torch.zeros(stride, dtype=torch.uint8)
        mine[:n_local] = blob
        allb = multi_gpu.gather_blobs(mine, world, dist)
        ok = [int(x.n_lines) for x in headers] == [10 + r for r in range(world)]
        for r in range(world):
            want = np.random.default_rng(r).integers(0, 256, 1000 + 777 * r, dtype=np.uint8)
            ok = ok and np.array_equal(allb[r, :int(headers[r].blob_bytes)].numpy(), want)
        # the one-collective form: [header | blob] padded to a stride every rank agrees on
        def write(dst):
            dst[:n_local] = blob
        stride = (max(int(x.blob_bytes) for x in headers) + multi_gpu.HEADER_BYTES + 255) & ~255
        hs, out, fits = multi_gpu.gather_one_shot(h, write, stride, world, dist, "cpu")
        ok = ok and fits and [int(x.n_reads) for x in hs] == [5 * r for r in range(world)]
        for r in range(world):
            want = np.random.default_rng(r).integers(0, 256, 1000 + 777 * r, dtype=np.uint8)
            got = out[r, multi_gpu.HEADER_BYTES:multi_gpu.HEADER_BYTES + int(hs[r].blob_bytes)]
            ok = ok and np.array_equal(got.numpy(), want)
        # a stride that is too small for the last rank: every rank must see that
        _, _, fits = multi_gpu.gather_one_shot(h, write, 1024 + 256, world, dist, "cpu")
        ok = ok and not fits
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_variable_size_all_gather_over_gloo_world_2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
    assert res == [(0, True), (1, True)]


class _FakeFilterBuilder:
    """Stands in for GraphBuilder in the host logic of ShardedBuilder.edge_filters: a view of rank-dependent size whose
    bytes are a known pattern; records what the sharded entry point is handed."""

    def __init__(self, rank):
        self.rank, self.seen = rank, None

    def set_stream(self, s):
        pass

    def _bytes(self, r):
        return np.random.default_rng(100 + r).integers(0, 256, 500 + 333 * r, dtype=np.uint8)

    def filter_view_header(self):
        h = _lib.FilterViewHeader()
        h.n_edges, h.n_orders, h.blob_bytes = 10 + self.rank, 20 + self.rank, self._bytes(self.rank).size
        return h

    def filter_view_export(self, ptr, cap):
        import ctypes
        b = self._bytes(self.rank)
        assert cap >= b.size
        ctypes.memmove(ptr, b.ctypes.data, b.size)

    def edge_filters_sharded(self, headers, ptrs, part, n_parts, wiggle):
        import ctypes
        got = [np.ctypeslib.as_array((ctypes.c_uint8 * int(h.blob_bytes)).from_address(int(p))).copy()
               for h, p in zip(headers, ptrs)]
        self.seen = (part, n_parts, wiggle, [int(h.n_edges) for h in headers], [int(h.n_orders) for h in headers], got)


def _filter_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fake = _FakeFilterBuilder(rank)
        sb = multi_gpu.ShardedBuilder(fake, rank, world, dist, "cpu")
        sb.edge_filters(123)
        part, n_parts, wiggle, ne, no, blobs = fake.seen
        ok = (part, n_parts, wiggle) == (rank, world, 123) and ne == [10 + r for r in range(world)]
        ok = ok and no == [20 + r for r in range(world)]
        for r in range(world):
            ok = ok and np.array_equal(blobs[r], fake._bytes(r))
        ok = ok and sb.filter_exchange_bytes == sum(fake._bytes(r).size for r in range(world) if r != rank)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_filter_views_exchange_over_gloo_world_2():
    """ShardedBuilder.edge_filters: sizes first, then ONE all-gather of the padded views; every rank hands the entry
    point all views in rank order."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_filter_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
    assert res == [(0, True), (1, True)]


class _FakeVerdictBuilder:
    """One rank's slice of the reference's filter verdicts (small x2), as GraphBuilder.counts() / .filters() give them."""

    def __init__(self, rank, world, broken=False):
        from oracle.pyoracle import FILTER_FIELDS, Reference
        paf = b"".join(synth.rank_chunk("small", r, 2) for r in range(2))
        f = Reference().filters(paf, 300)
        E = int(f["deletable"].shape[0])
        a, b = E * rank // world + (7 if 0 < rank else 0), E * (rank + 1) // world + (7 if rank + 1 < world else 0)
        self.f, self.broken = {k: f[k][a:b] for k in FILTER_FIELDS}, broken

    def counts(self):
        return {"n_edges": int(self.f["deletable"].shape[0])}

    def filters(self):
        if self.broken:
            raise RuntimeError("no verdicts on this rank")
        return self.f


def _parity_worker(rank, world, port, q):
    import sys
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        good = bench.check_filters_parity(_FakeVerdictBuilder(rank, world), "small", rank, world, dist)
        # a rank that cannot produce its verdicts: every rank still leaves both collectives, nobody claims parity
        bad = bench.check_filters_parity(_FakeVerdictBuilder(rank, world, broken=(rank == 1)), "small", rank, world, dist)
        q.put((rank, good["ok"] is True and good["digest"]["n_edges"] == 2941, bad["ok"] is None and "error" in bad))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_filter_verdict_parity_check_over_gloo_world_2():
    """bench.check_filters_parity: the digests of the ranks' slices (uneven cut) add up to the golden recorded from
    the reference; a failing rank degrades the answer to 'unknown' on every rank instead of hanging the others."""
    if not os.path.exists(os.path.join(os.path.dirname(os.path.abspath(_lib.__file__)), "..", "oracle", "_ref", "libmsref.so")):
        pytest.skip("oracle/_ref/libmsref.so not built")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_parity_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=150) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
    assert res == [(0, True, True), (1, True, True)]
