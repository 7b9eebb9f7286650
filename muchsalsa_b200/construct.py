"""Thin Python host layer over the C ABI (include/msgpu.h) — used by tests, bench.py and the multi-GPU driver.

The product is libmsgpu.so; this module only moves pointers.  Names follow the reference: a `GraphBuilder` does what
`BlastFileReader::read()` + `MatchMap::calculateEdges()` do (src/main.cpp:156-157), `SequenceStore` what
`SequenceAccessor` + `get{Illumina,Nanopore}Sequence` do (SequenceUtils.cpp:63-85).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _lib

STATUS = {0: "OK", 1: "INVALID_ARG", 2: "CUDA", 3: "IO", 4: "INVALID_PAF", 5: "NOMEM", 6: "LIMIT", 7: "NO_RESULT",
          8: "OUT_OF_RANGE", 9: "INTERNAL"}


class MsGpuError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"msgpu status {status} ({STATUS.get(status, '?')}): {message}")
        self.status = status
        self.message = message


@dataclass
class Construction:
    """Host copy of a construction result; row orders as documented in include/msgpu.h."""
    counts: dict
    read_names: list
    read_nlen: np.ndarray
    read_line: np.ndarray
    unitig_names: list
    matches: dict  # unitig read nr0 nr1 ir0 ir1 rratio score line flags
    edge_first: np.ndarray
    edge_second: np.ndarray
    edge_em_begin: np.ndarray
    edge_matches: dict  # unitig ov0 ov1 score line flags
    timings: dict


def _np(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dtype, copy=True)


def _names(off_ptr, blob_ptr, n):
    if n == 0 or not off_ptr:
        return []
    off = np.ctypeslib.as_array(off_ptr, shape=(int(n) + 1,))
    blob = C.string_at(blob_ptr, int(off[-1])) if off[-1] else b""
    return [blob[int(off[i]):int(off[i + 1])].decode("latin-1") for i in range(int(n))]


class GraphBuilder:
    """One msgpu context = one GPU.  `build_*` replaces read() + calculateEdges()."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = self._lib.msgpu_create(device, C.byref(h))
        if rc != 0:
            raise MsGpuError(rc, "msgpu_create failed: no usable CUDA device (there is no CPU fallback)")
        self._h = h
        self._keep = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.msgpu_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise MsGpuError(rc, self._lib.msgpu_last_error(self._h).decode("utf-8", "replace"))

    @property
    def handle(self):
        return self._h

    def set_stream(self, cuda_stream: int | None):
        """A cudaStream_t handle to run on; None / 0 = the context's own (non-blocking) stream.  Note that torch's
        default stream has handle 0: to share a stream with torch, use an explicit torch.cuda.Stream."""
        self._check(self._lib.msgpu_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    # ---- construction
    def build_from_paf(self, path):
        self._check(self._lib.msgpu_build_from_paf(self._h, os.fspath(path).encode()))
        return self

    def build_from_buffer(self, buf):
        """buf: bytes / bytearray / numpy uint8 array / anything with (data_ptr(), numel()) in HOST memory."""
        ptr, n, keep = _host_ptr(buf)
        self._keep = keep
        self._check(self._lib.msgpu_build_from_paf_buffer(self._h, C.c_void_p(ptr), n))
        return self

    def build_from_device(self, dev_ptr: int, n_bytes: int):
        self._check(self._lib.msgpu_build_from_paf_device(self._h, C.c_void_p(dev_ptr), n_bytes))
        return self

    def anchor_graph_from_buffer(self, buf):
        """The scrubber's anchor graph (pipeline/scrubber_bfs.py:57-114) of a PAF in host memory; see include/msgpu.h."""
        ptr, n, keep = _host_ptr(buf)
        self._keep = keep
        self._check(self._lib.msgpu_anchor_graph_from_paf_buffer(self._h, C.c_void_p(ptr), n))
        return self

    def anchor_graph_from_device(self, dev_ptr: int, n_bytes: int):
        self._check(self._lib.msgpu_anchor_graph_from_paf_device(self._h, C.c_void_p(dev_ptr), n_bytes))
        return self

    def unitig_coverage_from_buffer(self, buf) -> dict:
        """Pass 1 of the unitig filter (pipeline/unitig_filter.py:34-75): per block of consecutive lines naming one
        unitig its unitig id, distinct reads, length and maximum coverage (numpy copies); see include/msgpu.h."""
        ptr, n, keep = _host_ptr(buf)
        self._keep = keep
        c = _lib.Coverage()
        self._check(self._lib.msgpu_unitig_coverage_from_paf_buffer(self._h, C.c_void_p(ptr), n, C.byref(c)))
        B = int(c.n_blocks)
        return dict(block_unitig=_np(c.block_unitig, B, np.uint32), block_n_reads=_np(c.block_n_reads, B, np.uint32),
                    block_len=_np(c.block_len, B, np.uint32), block_max_cov=_np(c.block_max_cov, B, np.uint32))

    def unitig_coverage_cut(self, cutoff: float, q3: float) -> dict:
        """Pass 2 (unitig_filter.py:102-176): outlier flag per block and the fragments of the outlier blocks."""
        f = _lib.Fragments()
        self._check(self._lib.msgpu_unitig_coverage_cut(self._h, float(cutoff), float(q3), C.byref(f)))
        B, F = int(f.n_blocks), int(f.n_fragments)
        return dict(block_outlier=_np(f.block_outlier, B, np.uint8), frag_block=_np(f.frag_block, F, np.uint32),
                    frag_start=_np(f.frag_start, F, np.int32), frag_end=_np(f.frag_end, F, np.int32))

    def build_from_records(self, rec: dict, n_reads: int, n_unitigs: int):
        """rec: dict of torch CUDA tensors (unitig, read, ir0, ir1, nr0, nr1, score, line, flags)."""
        soa = _lib.MatchSoA()
        soa.n = int(rec["unitig"].numel())
        for k in ("unitig", "read", "ir0", "ir1", "nr0", "nr1", "score", "line", "flags"):
            setattr(soa, k, rec[k].data_ptr())
        self._keep = rec
        self._check(self._lib.msgpu_build_from_records(self._h, C.byref(soa), n_reads, n_unitigs))
        return self

    # ---- multi-GPU construction (see muchsalsa_b200/multi_gpu.py for the collectives between the steps)
    def shard_ingest_buffer(self, buf, is_last_shard: bool):
        ptr, n, keep = _host_ptr(buf)
        self._keep = keep
        self._check(self._lib.msgpu_shard_ingest_buffer(self._h, C.c_void_p(ptr), n, 1 if is_last_shard else 0))
        return self

    def shard_ingest_device(self, dev_ptr: int, n_bytes: int, is_last_shard: bool):
        self._check(self._lib.msgpu_shard_ingest_device(self._h, C.c_void_p(dev_ptr), n_bytes,
                                                        1 if is_last_shard else 0))
        return self

    def shard_header(self) -> "_lib.ShardHeader":
        h = _lib.ShardHeader()
        self._check(self._lib.msgpu_shard_describe(self._h, C.byref(h)))
        return h

    def shard_export(self, dev_ptr: int, capacity: int):
        self._check(self._lib.msgpu_shard_export(self._h, C.c_void_p(dev_ptr), capacity))
        return self

    def shard_export_peers(self, dev_slots, capacity: int):
        """The blob into this rank's slot of every rank's receive buffer (peer-mapped device pointers) in one kernel."""
        arr = (C.c_void_p * len(dev_slots))(*[int(p) for p in dev_slots])
        self._check(self._lib.msgpu_shard_export_peers(self._h, arr, len(dev_slots), capacity))
        return self

    def build_from_shards(self, headers, blob_ptrs, part: int, n_parts: int):
        """headers: list of _lib.ShardHeader (one per rank, rank order); blob_ptrs: device pointers of the blobs."""
        n = len(headers)
        hs = (_lib.ShardHeader * n)(*headers)
        ps = (C.c_void_p * n)(*[C.c_void_p(int(p)) for p in blob_ptrs])
        self._check(self._lib.msgpu_build_from_shards(self._h, n, hs, ps, part, n_parts))
        return self

    # ---- results
    def counts(self) -> dict:
        c = _lib.Counts()
        self._check(self._lib.msgpu_result_counts(self._h, C.byref(c)))
        return {k: int(getattr(c, k)) for k, _ in c._fields_}

    def timings(self) -> dict:
        t = _lib.Timings()
        self._check(self._lib.msgpu_result_timings(self._h, C.byref(t)))
        return {k: float(getattr(t, k)) for k, _ in t._fields_}

    def profile(self, enable: bool):
        self._check(self._lib.msgpu_profile_enable(self._h, 1 if enable else 0))

    def kernel_stats(self) -> dict:
        """{kernel name: (launches, total ms)} of the last build (needs profile(True) before it)."""
        arr = (_lib.KernelStat * 128)()
        n = C.c_size_t()
        self._check(self._lib.msgpu_profile_read(self._h, arr, 128, C.byref(n)))
        return {arr[i].name.decode(): (int(arr[i].launches), float(arr[i].total_ms)) for i in range(min(n.value, 128))}

    def launch_count(self) -> int:
        return int(self._lib.msgpu_launch_count(self._h))

    def device_tables(self) -> _lib.Tables:
        t = _lib.Tables()
        self._check(self._lib.msgpu_result_device(self._h, C.byref(t)))
        return t

    def host_tables(self) -> _lib.Tables:
        """Pointers into the library's pinned host slab (valid until the next build)."""
        t = _lib.Tables()
        self._check(self._lib.msgpu_result_host(self._h, C.byref(t)))
        return t

    def result(self) -> Construction:
        t = self.host_tables()  # (a chunked build learns its row counts while the chunks are downloaded)
        c = self.counts()
        R, U, M, E, EM = c["n_reads"], c["n_unitigs"], c["n_matches"], c["n_edges"], c["n_edge_matches"]
        m = dict(unitig=_np(t.m_unitig, M, np.uint32), read=_np(t.m_read, M, np.uint32),
                 nr0=_np(t.m_nr0, M, np.int32), nr1=_np(t.m_nr1, M, np.int32), ir0=_np(t.m_ir0, M, np.int32),
                 ir1=_np(t.m_ir1, M, np.int32), rratio=_np(t.m_rratio, M, np.float64),
                 score=_np(t.m_score, M, np.int32), line=_np(t.m_line, M, np.uint32),
                 flags=_np(t.m_flags, M, np.uint8))
        em = dict(unitig=_np(t.em_unitig, EM, np.uint32), ov0=_np(t.em_ov0, EM, np.int32),
                  ov1=_np(t.em_ov1, EM, np.int32), score=_np(t.em_score, EM, np.float64),
                  line=_np(t.em_line, EM, np.uint32), flags=_np(t.em_flags, EM, np.uint8))
        have_names = bool(t.read_name_off)
        return Construction(
            counts=c,
            read_names=_names(t.read_name_off, t.read_names, R) if have_names else [],
            read_nlen=_np(t.read_nlen, R if have_names else 0, np.int32),
            read_line=_np(t.read_line, R if have_names else 0, np.uint32),
            unitig_names=_names(t.unitig_name_off, t.unitig_names, U) if have_names else [],
            matches=m, edge_first=_np(t.edge_first, E, np.uint32), edge_second=_np(t.edge_second, E, np.uint32),
            edge_em_begin=_np(t.edge_em_begin, E + 1, np.uint64), edge_matches=em, timings=self.timings())

    # ---- chunked construction (results larger than HBM; downloads overlapped with the pairing)
    def set_chunk_tiles(self, tiles_per_chunk: int):
        """> 0: the builds that follow stop after planning; chunk_next() / host_tables() produce the edge tables
        piece by piece.  0: resident (default)."""
        self._check(self._lib.msgpu_set_chunk_tiles(self._h, int(tiles_per_chunk)))
        return self

    def chunk_next(self):
        """The next chunk (a _lib.Chunk with device pointers), or None when every chunk has been delivered."""
        c = _lib.Chunk()
        rc = self._lib.msgpu_chunk_next(self._h, C.byref(c))
        if rc == 7:
            return None
        self._check(rc)
        return c

    def chunk_digest(self) -> dict:
        d = _lib.Digest()
        self._check(self._lib.msgpu_chunk_digest(self._h, C.byref(d)))
        return {k: int(getattr(d, k)) for k, _ in d._fields_}

    def digest(self) -> dict:
        """Order-independent checksum + invariant violations of the edge tables, computed on the device
        (msgpu_result_digest); the digests of slices / chunks add up mod 2^64."""
        d = _lib.Digest()
        self._check(self._lib.msgpu_result_digest(self._h, C.byref(d)))
        return {k: int(getattr(d, k)) for k, _ in d._fields_}

    # ---- chaining + overlaps (the stage after construction)
    def chain_overlaps(self, wiggle_room: int = 300):
        """chainingAndOverlaps for every edge of the last build (src/main.cpp:328-414)."""
        self._check(self._lib.msgpu_chain_overlaps(self._h, wiggle_room))
        return self

    def chain_ms(self) -> float:
        ms = C.c_float()
        self._check(self._lib.msgpu_chain_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def orders(self) -> dict:
        """Host copy of the shadow flags and EdgeOrder rows in canonical (compacted) form: edges in table order,
        orders in the order the reference appends them."""
        o = _lib.Orders()
        self._check(self._lib.msgpu_orders_host(self._h, C.byref(o)))
        E, S = int(o.n_edges), int(o.n_slots)
        t = self.host_tables()
        begin = _np(t.edge_em_begin, E + 1, np.uint64).astype(np.int64)
        n_ord = _np(o.edge_n_orders, E, np.uint32).astype(np.int64)
        order_begin = np.zeros(E + 1, np.int64)
        np.cumsum(n_ord, out=order_begin[1:])
        # slot of every order: edge_em_begin[e] + i
        edge_of = np.repeat(np.arange(E), n_ord)
        slot = begin[edge_of] + (np.arange(int(order_begin[-1])) - order_begin[edge_of])
        flags = _np(o.ord_flags, S, np.uint8)[slot]
        n_ids = _np(o.ord_n_ids, S, np.uint32)[slot].astype(np.int64)
        ids_begin = np.zeros(slot.size + 1, np.int64)
        np.cumsum(n_ids, out=ids_begin[1:])
        # ids of edge e: ord_ids[edge_em_begin[e] : + sum of its orders' lengths]
        per_edge = np.zeros(E, np.int64)
        np.add.at(per_edge, edge_of, n_ids)
        id_edge = np.repeat(np.arange(E), per_edge)
        id_first = np.zeros(E + 1, np.int64)
        np.cumsum(per_edge, out=id_first[1:])
        id_slot = begin[id_edge] + (np.arange(int(id_first[-1])) - id_first[id_edge])
        return dict(edge_shadow=_np(o.edge_shadow, E, np.uint8), edge_order_begin=order_begin.astype(np.uint64),
                    start_is_first=(flags & 1).astype(np.uint8), contained=((flags >> 1) & 1).astype(np.uint8),
                    direction=((flags >> 2) & 1).astype(np.uint8), primary=((flags >> 3) & 1).astype(np.uint8),
                    left=_np(o.ord_left, S, np.float64)[slot], right=_np(o.ord_right, S, np.float64)[slot],
                    score=_np(o.ord_score, S, np.uint64)[slot], ids_begin=ids_begin.astype(np.uint64),
                    ids=_np(o.ord_ids, S, np.uint32)[id_slot])

    # ---- per-edge passes after chaining
    def edge_filters(self, wiggle_room: int = 300):
        """findContractionEdges, findDeletableEdges and computeBitweight for every edge of the last chaining
        (src/main.cpp:416-460, 534-573)."""
        self._check(self._lib.msgpu_edge_filters(self._h, wiggle_room))
        return self

    def filter_view_header(self) -> "_lib.FilterViewHeader":
        """Sizes of this rank's view of its edges and orders (multi-GPU edge filters; after chain_overlaps)."""
        h = _lib.FilterViewHeader()
        self._check(self._lib.msgpu_filter_view_describe(self._h, C.byref(h)))
        return h

    def filter_view_export(self, dev_ptr: int, capacity: int):
        self._check(self._lib.msgpu_filter_view_export(self._h, C.c_void_p(dev_ptr), capacity))
        return self

    def edge_filters_sharded(self, headers, blob_ptrs, part: int, n_parts: int, wiggle_room: int = 300):
        """The edge filters on one slice of a multi-GPU build: headers[r] / blob_ptrs[r] = the view of rank r (device
        pointers, rank order).  Verdicts for this rank's edges (filters())."""
        hs = (_lib.FilterViewHeader * n_parts)(*headers)
        ps = (C.c_void_p * n_parts)(*[C.c_void_p(int(p)) for p in blob_ptrs])
        self._check(self._lib.msgpu_edge_filters_sharded(self._h, n_parts, part, hs, ps, wiggle_room))
        return self

    def edge_filters_ms(self) -> float:
        ms = C.c_float()
        self._check(self._lib.msgpu_edge_filters_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def filters(self) -> dict:
        """Host copy of the per-edge verdicts, edges in table order."""
        f = _lib.EdgeFilters()
        self._check(self._lib.msgpu_edge_filters_host(self._h, C.byref(f)))
        E = int(f.n_edges)
        return dict(contraction_order=_np(f.contraction_order, E, np.int32), deletable=_np(f.deletable, E, np.uint8),
                    n_kept_orders=_np(f.n_kept_orders, E, np.uint32), weight=_np(f.weight, E, np.uint64),
                    consensus=_np(f.consensus, E, np.uint8))

    # ---- sequences
    def seq_load(self, unitigs_path, nanopore_path):
        self._check(self._lib.msgpu_seq_load(self._h, os.fspath(unitigs_path).encode(),
                                             os.fspath(nanopore_path).encode()))
        return self

    def seq_load_buffers(self, unitigs: bytes, nanopore: bytes, nanopore_is_fastq: bool):
        self._check(self._lib.msgpu_seq_load_buffers(self._h, unitigs, len(unitigs), nanopore, len(nanopore),
                                                     1 if nanopore_is_fastq else 0))
        return self

    def seq_id(self, which: int, name: str):
        out = C.c_uint32()
        rc = self._lib.msgpu_seq_id(self._h, which, name.encode("latin-1"), C.byref(out))
        if rc == 8:
            return None
        self._check(rc)
        return int(out.value)

    def seq_fetch_batch(self, reqs):
        """reqs: iterable of (which, id, left, right, direction).  Returns a list of bytes, None where the reference
        would throw."""
        reqs = list(reqs)
        n = len(reqs)
        arr = (_lib.SeqReq * max(1, n))()
        for i, (w, id_, l, r, d) in enumerate(reqs):
            arr[i] = _lib.SeqReq(w, id_, l, r, 1 if d else 0)
        bound = C.c_uint64()
        self._check(self._lib.msgpu_seq_batch_bound(self._h, arr, n, C.byref(bound)))
        out = C.create_string_buffer(int(bound.value) + 1)
        off = (C.c_uint64 * (n + 1))()
        st = (C.c_uint8 * max(1, n))()
        rc = self._lib.msgpu_seq_fetch_batch(self._h, arr, n, out, int(bound.value), off, st)
        if rc != 0:
            raise MsGpuError(rc, "msgpu_seq_fetch_batch failed")
        raw = out.raw
        return [None if st[i] else raw[off[i]:off[i + 1]] for i in range(n)]


    def seq_fetch_batch_device(self, dev_reqs: int, n: int, dev_out: int, capacity: int, dev_out_off: int,
                               dev_status: int):
        """msgpu_seq_fetch_batch_device: every buffer resident in HBM (device pointers as ints)."""
        self._check(self._lib.msgpu_seq_fetch_batch_device(self._h, C.c_void_p(dev_reqs), n, C.c_void_p(dev_out),
                                                           capacity, C.c_void_p(dev_out_off), C.c_void_p(dev_status)))
        return self

    def seq_fetch_batch_host(self, reqs: np.ndarray, out: np.ndarray, off: np.ndarray, status: np.ndarray):
        """msgpu_seq_fetch_batch on caller-owned numpy buffers: reqs [n, 5] uint32 (which, id, left, right, direction),
        out uint8 [capacity], off uint64 [n + 1], status uint8 [n]."""
        n = int(reqs.shape[0])
        rc = self._lib.msgpu_seq_fetch_batch(self._h, C.cast(reqs.ctypes.data, C.POINTER(_lib.SeqReq)), n,
                                             C.c_void_p(out.ctypes.data), int(out.size),
                                             C.cast(off.ctypes.data, _lib.u64p), C.cast(status.ctypes.data, _lib.u8p))
        if rc != 0:
            raise MsGpuError(rc, "msgpu_seq_fetch_batch failed")
        return self


class _DeviceSpan:
    """A device pointer + length as a __cuda_array_interface__ object (v2), so torch / cupy can wrap it without a
    copy.  The memory belongs to the library: valid until the next build on the context."""

    def __init__(self, ptr: int, n: int, typestr: str, keep=None):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}
        self._keep = keep


_TYPESTR = {"u1": "|u1", "i4": "<i4", "u4": "<u4", "i8": "<i8", "u8": "<u8", "f8": "<f8"}


def device_view(ptr, n: int, kind: str):
    """torch tensor over `n` elements of library-owned device memory (kind: u1 i4 u4 i8 u8 f8; unsigned 32/64-bit
    columns are viewed as the signed type of the same width, which torch can index with).  Host plumbing for tests
    and bench.py — nothing is copied."""
    import torch
    addr = C.cast(ptr, C.c_void_p).value if not isinstance(ptr, int) else ptr
    if not addr or n == 0:
        return torch.zeros(0, dtype={"u1": torch.uint8, "i4": torch.int32, "u4": torch.int32, "i8": torch.int64,
                                     "u8": torch.int64, "f8": torch.float64}[kind], device="cuda")
    signed = {"u4": "i4", "u8": "i8"}.get(kind, kind)
    return torch.as_tensor(_DeviceSpan(addr, n, _TYPESTR[signed]), device="cuda")


def _host_ptr(buf):
    if hasattr(buf, "data_ptr"):  # torch tensor (pinned or pageable host memory)
        return int(buf.data_ptr()), int(buf.numel() * buf.element_size()), buf
    if isinstance(buf, np.ndarray):
        a = np.ascontiguousarray(buf)
        return int(a.ctypes.data), int(a.nbytes), a
    if isinstance(buf, (bytes, bytearray)):
        a = np.frombuffer(buf, np.uint8)
        return int(a.ctypes.data), int(a.nbytes), (a, buf)
    raise TypeError(f"unsupported buffer type {type(buf)}")
