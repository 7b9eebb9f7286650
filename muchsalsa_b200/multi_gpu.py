"""Multi-GPU graph construction: one process per GPU, `torch.distributed` for the two collectives between the stages
of the C ABI (include/msgpu.h, "multi-GPU construction").

    rank r:  chunk r of the PAF --msgpu_shard_ingest--> chunk-local tables --msgpu_shard_export--> flat blob
    all ranks: all-gather(headers), all-gather(blobs)            <- the only bytes that cross NVLink (match records)
    rank r:  msgpu_build_from_shards(all blobs, part=r, n_parts=N)  -> global read/unitig/match tables (replicated)
                                                                     + rank r's contiguous slice of the edge tables

The reference is a single process with a thread pool (src/main.cpp:143-157); SURVEY.md §8e derives the sharding: the
unit of pairing is one unitig's group, the reduction key is the read pair.  Here the rows of the edge table (all edge
matches whose `first` is one read) are the unit that is partitioned, so every rank needs the whole match table (small)
and no candidate pair is ever exchanged.

The collectives work on CPU tensors with the `gloo` backend too (tests/test_multi_cpu.py exercises them with two
processes); everything that touches the GPU goes through muchsalsa_b200.construct.GraphBuilder.
"""
from __future__ import annotations

import time

import torch

from . import _lib

HEADER_WORDS = 8  # msgpu_shard_header = 8 x uint64


def split_paf(paf: bytes, n: int) -> list[bytes]:
    """n contiguous chunks of about equal size, cut after a newline.  Concatenated they are the input again.

    The LAST chunk always holds the last line of the file (earlier chunks may be empty for tiny inputs): the rank that
    ingests it with is_last_shard=1 is the one that drops that line, as BlastFileReader::read() does
    (BlastFileReader.cpp:76) — a last line parsed by an earlier rank would be a line the reference never sees."""
    last_line = paf.rfind(b"\n", 0, max(len(paf) - 1, 0)) + 1  # where the last line starts
    cuts = [0]
    for r in range(1, n):
        want = max(cuts[-1], len(paf) * r // n)
        nl = paf.find(b"\n", want)
        cuts.append(min(len(paf) if nl < 0 else nl + 1, last_line))
    cuts.append(len(paf))
    return [paf[cuts[i]:cuts[i + 1]] for i in range(n)]


def header_to_tensor(h: "_lib.ShardHeader", device) -> torch.Tensor:
    vals = [int(getattr(h, name)) for name, _ in h._fields_]
    # uint64 magic does not fit int64 as a positive number: carry the bit pattern
    vals = [v - (1 << 64) if v >= (1 << 63) else v for v in vals]
    return torch.tensor(vals, dtype=torch.int64, device=device)


def tensor_to_headers(t: torch.Tensor) -> list:
    out = []
    for row in t.cpu().tolist():
        h = _lib.ShardHeader()
        for (name, _), v in zip(h._fields_, row):
            setattr(h, name, v + (1 << 64) if v < 0 else v)
        out.append(h)
    return out


def gather_headers(local: torch.Tensor, world: int, dist) -> torch.Tensor:
    """[HEADER_WORDS] int64 on every rank -> [world, HEADER_WORDS] (rank order), on the same device."""
    out = torch.empty(world * HEADER_WORDS, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(out, local)
    return out.view(world, HEADER_WORDS)


def gather_blobs(mine: torch.Tensor, world: int, dist) -> torch.Tensor:
    """Every rank contributes `mine` (uint8, the same padded length on all ranks) -> [world, len] in rank order."""
    out = torch.empty(world * mine.numel(), dtype=torch.uint8, device=mine.device)
    dist.all_gather_into_tensor(out, mine)
    return out.view(world, mine.numel())


HEADER_BYTES = HEADER_WORDS * 8


def gather_one_shot(h: "_lib.ShardHeader", write_blob, stride: int, world: int, dist, device):
    """The exchange as ONE collective: every rank contributes [header | blob] padded to `stride` bytes (the same
    value on every rank, a multiple of 256).  `write_blob(dst_tensor)` fills the blob part; it is not called when this
    rank's blob does not fit.  Returns (headers, gathered [world, stride] tensor, ok): ok is False iff some rank's
    blob did not fit — every rank sees the same headers, so every rank reaches the same verdict and can fall back to
    the two-collective path together."""
    buf = torch.empty(stride, dtype=torch.uint8, device=device)
    buf[:HEADER_BYTES] = header_to_tensor(h, device).view(torch.uint8)
    if int(h.blob_bytes) + HEADER_BYTES <= stride:
        write_blob(buf[HEADER_BYTES:])
    out = gather_blobs(buf, world, dist) if world > 1 else buf.view(1, stride)
    heads = out[:, :HEADER_BYTES].contiguous().view(torch.int64).view(world, HEADER_WORDS)
    headers = tensor_to_headers(heads)  # device -> host: the one synchronisation of the exchange
    ok = all(int(x.blob_bytes) + HEADER_BYTES <= stride for x in headers)
    return headers, out, ok


class ShardedBuilder:
    """The N-GPU construction step of one rank.  `dist` is torch.distributed (initialised) or None for world == 1."""

    def __init__(self, builder, rank: int, world: int, dist, device, stream=None):
        """`stream`: the torch CUDA stream everything of this builder runs on — the library's kernels, the tensors
        handed to the collectives and (through torch's stream bookkeeping) NCCL's reads and writes of them.  It must
        not be the default stream: its handle is 0, which the C ABI reads as "the context's own stream", and work on
        two unrelated streams is not ordered.  None = a new stream."""
        self.gb, self.rank, self.world, self.dist, self.device = builder, rank, world, dist, device
        self.exchange_bytes = 0   # bytes this rank received from the others in the last step
        self._keep = None
        self.stream = None
        if torch.device(device).type == "cuda":
            self.stream = stream if stream is not None else torch.cuda.Stream(device)
            if self.stream.cuda_stream == 0:
                raise ValueError("ShardedBuilder needs a non-default CUDA stream")
            builder.set_stream(self.stream.cuda_stream)

    def _on_stream(self):
        import contextlib
        return torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()

    def _exchange(self):
        with self._on_stream():
            return self._exchange_on_stream()

    def _exchange_on_stream(self):
        gb, world = self.gb, self.world
        h = gb.shard_header()
        stride = getattr(self, "_stride", None)
        peer = self._peer_buffers(stride) if stride is not None else None
        if peer is not None:
            # steady state on NVLink-connected GPUs: the export kernel stores the blob straight into every rank's
            # receive buffer (msgpu_shard_export_peers), then ONE barrier across the ranks.  Two buffer sets alternate:
            # a rank that is one step ahead writes the set its peers finished reading a step ago (its own barrier of
            # the step in between completed only after every peer had queued that step's export behind those reads).
            ev = self._timer()
            tensor, hdl, ptrs = peer
            self._peer_set = 1 - getattr(self, "_peer_set", 1)
            base = self._peer_set * world * stride
            gb.shard_export_peers([p + base + self.rank * stride for p in ptrs], stride)
            hdl.barrier(channel=0)
            local = tensor[base:base + world * stride].view(world, stride)
            heads = local[:, :HEADER_BYTES].contiguous().view(torch.int64).view(world, HEADER_WORDS)
            headers = tensor_to_headers(heads)  # device -> host: the one synchronisation of the exchange
            if all(int(x.blob_bytes) <= stride for x in headers):
                self._stop(ev)
                self.exchange_mode = "peer stores (export kernel writes every rank's buffer over NVLink) + 1 barrier"
                self.exchange_bytes = sum(int(x.blob_bytes) for i, x in enumerate(headers) if i != self.rank)
                self._keep = tensor
                return headers, [local.data_ptr() + s * stride for s in range(world)]
            self._peer = None  # a blob outgrew its slot (every rank sees that): sizes first, below; new buffers next step
            stride = None
        if stride is not None:
            # steady state: the stride that fitted every rank last time (plus slack) — ONE collective, one host sync
            ev = self._timer()
            headers, blobs, ok = gather_one_shot(
                h, lambda dst: gb.shard_export(dst.data_ptr(), dst.numel()), stride, world, self.dist, self.device)
            if ok:
                self._stop(ev)
                self.exchange_mode = "all-gather of [header | blob] slabs (NCCL over NVLink), one collective"
                self.exchange_bytes = sum(int(x.blob_bytes) for i, x in enumerate(headers) if i != self.rank)
                self._keep = blobs
                return headers, [blobs.data_ptr() + s * stride + HEADER_BYTES for s in range(world)]
        # first step, or a blob outgrew the stride: sizes first, then the padded blobs
        heads = gather_headers(header_to_tensor(h, self.device), world, self.dist) if world > 1 else \
            header_to_tensor(h, self.device).view(1, HEADER_WORDS)
        headers = tensor_to_headers(heads)  # device -> host: the sizes decide the allocations below
        biggest = max(int(x.blob_bytes) for x in headers)
        stride = (biggest + 255) & ~255
        self._stride = (biggest + biggest // 8 + HEADER_BYTES + 255) & ~255  # what the next steps try
        mine = torch.empty(stride, dtype=torch.uint8, device=self.device)
        gb.shard_export(mine.data_ptr(), stride)
        if world > 1:
            ev = self._timer()
            blobs = gather_blobs(mine, world, self.dist)
            self._stop(ev)
        else:
            blobs = mine.view(1, stride)
        self.exchange_mode = "all-gather of headers, then of blobs (NCCL over NVLink)"
        self.exchange_bytes = sum(int(x.blob_bytes) for i, x in enumerate(headers) if i != self.rank)
        self._keep = (mine, blobs)
        return headers, [blobs.data_ptr() + s * stride for s in range(world)]

    PEER_BUFFER_LIMIT = 24 << 30  # bytes of peer-mapped receive buffers (two sets) a rank is willing to hold

    def _peer_buffers(self, stride):
        """(tensor, handle, peer pointers) of the symmetric receive buffers [2 sets][world][stride], allocated and
        exchanged on first use (collective: every rank calls it in the same step with the same stride), or None when
        peer memory is not available / disabled (MSGPU_PEER_EXCHANGE=0) — then the NCCL all-gather carries the blobs."""
        import os
        if self.world == 1 or self.stream is None or getattr(self, "_peer_off", False):
            return None
        cur = getattr(self, "_peer", None)
        if cur is not None and cur[3] == stride:
            return cur[:3]
        if os.environ.get("MSGPU_PEER_EXCHANGE", "1") == "0" or 2 * self.world * stride > self.PEER_BUFFER_LIMIT:
            self._peer_off = True
            return None
        ok, made = 1, None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            t = symm_mem.empty(2 * self.world * stride, dtype=torch.uint8, device=self.device)
            hdl = symm_mem.rendezvous(t, self.dist.group.WORLD)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            if len(ptrs) != self.world or any(p % 16 for p in ptrs):
                raise RuntimeError("unexpected symmetric-memory layout")
            made = (t, hdl, ptrs, stride)
        except Exception as e:  # no peer access on this box: every rank must take the same decision (below)
            ok = 0
            self.peer_error = repr(e)
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            self._peer_off = True
            return None
        self._peer = made
        return made[:3]

    def _timer(self):
        if self.stream is None or self.world == 1:
            return None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        return (a, b)

    def _stop(self, ev):
        if ev is not None:
            ev[1].record()
            self._events = ev

    def build_device(self, dev_ptr: int, n_bytes: int):
        self.gb.shard_ingest_device(dev_ptr, n_bytes, self.rank == self.world - 1)
        headers, ptrs = self._exchange()
        self.gb.build_from_shards(headers, ptrs, self.rank, self.world)
        return self.gb

    def build_buffer(self, host_buf):
        self.gb.shard_ingest_buffer(host_buf, self.rank == self.world - 1)
        headers, ptrs = self._exchange()
        self.gb.build_from_shards(headers, ptrs, self.rank, self.world)
        return self.gb

    def edge_filters(self, wiggle_room: int = 300):
        """findContractionEdges / findDeletableEdges / computeBitweight for this rank's slice of the edges (after
        `chain_overlaps` on the builder).  The contraction test reads edges that live on other ranks, so every rank
        publishes a compact view of its slice (msgpu_filter_view_*), ONE all-gather carries the views, and the passes
        run against their concatenation (msgpu_edge_filters_sharded).  Returns the builder (verdicts: .filters())."""
        gb, world = self.gb, self.world
        with self._on_stream():
            h = gb.filter_view_header()
            if world == 1:
                blob = torch.empty(max(int(h.blob_bytes), 16), dtype=torch.uint8, device=self.device)
                gb.filter_view_export(blob.data_ptr(), blob.numel())
                gb.edge_filters_sharded([h], [blob.data_ptr()], 0, 1, wiggle_room)
                return gb
            mine = torch.tensor([int(h.n_edges), int(h.n_orders), int(h.blob_bytes), 0], dtype=torch.int64,
                                device=self.device)
            heads = torch.empty(world * 4, dtype=torch.int64, device=self.device)
            self.dist.all_gather_into_tensor(heads, mine)
            rows = heads.view(world, 4).cpu().tolist()  # device -> host: the sizes decide the allocations below
            headers = []
            for row in rows:
                x = _lib.FilterViewHeader()
                x.n_edges, x.n_orders, x.blob_bytes, x.reserved = row
                headers.append(x)
            stride = (max(int(x.blob_bytes) for x in headers) + 255) & ~255
            blob = torch.empty(max(stride, 256), dtype=torch.uint8, device=self.device)
            gb.filter_view_export(blob.data_ptr(), blob.numel())
            views = gather_blobs(blob, world, self.dist)
            self.filter_exchange_bytes = sum(int(x.blob_bytes) for i, x in enumerate(headers) if i != self.rank)
            gb.edge_filters_sharded(headers, [views.data_ptr() + r * blob.numel() for r in range(world)], self.rank,
                                    world, wiggle_room)
            self._keep_views = views
        return gb

    def last_exchange_ms(self):
        ev = getattr(self, "_events", None)
        if not ev:
            return None
        ev[1].synchronize()
        return ev[0].elapsed_time(ev[1])
