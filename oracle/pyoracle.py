"""ctypes bindings for the two CHECKERS under oracle/ — TEST INFRASTRUCTURE ONLY.

    Oracle()     oracle/libms_oracle.so   plain-C restatement of the reference hot path (ms_oracle.c)
    Reference()  oracle/_ref/libmsref.so  the reference's own sources compiled with g++ (ref_driver.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module; nothing under
muchsalsa_b200/ does (tests/test_layout.py enforces it).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libms_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libmsref.so")
REF_CLI = os.path.join(HERE, "_ref", "msref_cli")


def build(ref: bool = True) -> None:
    """(Re)build the checkers with oracle/Makefile.  `ref` needs /root/reference and is skipped without it."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"] + (["ref"] if ref else []), check=True)


class _MsTables(C.Structure):
    _P = C.POINTER
    _fields_ = [
        ("n_reads", C.c_uint64), ("read_name_off", _P(C.c_uint64)), ("read_names", _P(C.c_char)),
        ("read_nlen", _P(C.c_uint64)), ("read_line", _P(C.c_uint64)),
        ("n_unitigs", C.c_uint64), ("unitig_name_off", _P(C.c_uint64)), ("unitig_names", _P(C.c_char)),
        ("n_vm", C.c_uint64), ("vm_read", _P(C.c_uint32)), ("vm_unitig", _P(C.c_uint32)),
        ("vm_nr0", _P(C.c_int32)), ("vm_nr1", _P(C.c_int32)), ("vm_ir0", _P(C.c_int32)), ("vm_ir1", _P(C.c_int32)),
        ("vm_rratio", _P(C.c_double)), ("vm_dir", _P(C.c_uint8)), ("vm_prim", _P(C.c_uint8)),
        ("vm_score", _P(C.c_uint64)), ("vm_line", _P(C.c_uint64)),
        ("n_edges", C.c_uint64), ("edge_first", _P(C.c_uint32)), ("edge_second", _P(C.c_uint32)),
        ("edge_em_begin", _P(C.c_uint64)),
        ("n_em", C.c_uint64), ("em_unitig", _P(C.c_uint32)), ("em_ov0", _P(C.c_int32)), ("em_ov1", _P(C.c_int32)),
        ("em_score", _P(C.c_double)), ("em_dir", _P(C.c_uint8)), ("em_prim", _P(C.c_uint8)),
        ("em_line", _P(C.c_uint64)),
        ("t_index", C.c_double), ("t_read", C.c_double), ("t_edges", C.c_double),
    ]


class _MsOrders(C.Structure):
    _P = C.POINTER
    _fields_ = [
        ("n_edges", C.c_uint64), ("edge_shadow", _P(C.c_uint8)), ("edge_order_begin", _P(C.c_uint64)),
        ("n_orders", C.c_uint64), ("ord_start_is_first", _P(C.c_uint8)), ("ord_left", _P(C.c_double)),
        ("ord_right", _P(C.c_double)), ("ord_contained", _P(C.c_uint8)), ("ord_direction", _P(C.c_uint8)),
        ("ord_primary", _P(C.c_uint8)), ("ord_score", _P(C.c_uint64)), ("ord_ids_begin", _P(C.c_uint64)),
        ("n_ids", C.c_uint64), ("ord_ids", _P(C.c_uint32)), ("t_chain", C.c_double),
    ]


class _MsFilters(C.Structure):
    _P = C.POINTER
    _fields_ = [
        ("n_edges", C.c_uint64), ("contraction_order", _P(C.c_int32)), ("deletable", _P(C.c_uint8)),
        ("n_kept_orders", _P(C.c_uint32)), ("weight", _P(C.c_uint64)), ("consensus", _P(C.c_uint8)),
        ("t_filters", C.c_double), ("t_chain", C.c_double),
    ]


FILTER_FIELDS = ("contraction_order", "deletable", "n_kept_orders", "weight", "consensus")


@dataclass
class Orders:
    """Canonical result of chainingAndOverlaps (see oracle/ms_tables.h, ms_orders): per edge a shadow flag and its
    EdgeOrder list; edges in (first, second) order."""
    edge_shadow: np.ndarray
    edge_order_begin: np.ndarray
    start_is_first: np.ndarray
    left: np.ndarray
    right: np.ndarray
    contained: np.ndarray
    direction: np.ndarray
    primary: np.ndarray
    score: np.ndarray
    ids_begin: np.ndarray
    ids: np.ndarray
    seconds: float = 0.0


def orders_equal(a: Orders, b: Orders) -> list:
    """Names of the fields in which two dumps differ (doubles compared bit for bit)."""
    bad = []
    for f in ("edge_shadow", "edge_order_begin", "start_is_first", "left", "right", "contained", "direction",
              "primary", "score", "ids_begin", "ids"):
        x, y = getattr(a, f), getattr(b, f)
        if x.dtype == np.float64:
            x, y = x.view(np.uint64), np.asarray(y, np.float64).view(np.uint64)
        if x.shape != y.shape or not np.array_equal(x, y):
            bad.append(f)
    return bad


@dataclass
class Tables:
    """Canonical construction result (see oracle/ms_tables.h for the sort orders)."""
    read_names: list
    read_nlen: np.ndarray
    read_line: np.ndarray
    unitig_names: list
    vm: dict  # read, unitig, nr0, nr1, ir0, ir1, rratio, dir, prim, score, line
    edge_first: np.ndarray
    edge_second: np.ndarray
    edge_em_begin: np.ndarray
    em: dict  # unitig, ov0, ov1, score, dir, prim, line
    times: dict

    @property
    def counts(self):
        return dict(reads=len(self.read_names), unitigs=len(self.unitig_names), vm=int(self.vm["read"].size),
                    edges=int(self.edge_first.size), em=int(self.em["unitig"].size))


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dtype, copy=True)


def _names(off_ptr, blob_ptr, n):
    if n == 0:
        return []
    off = np.ctypeslib.as_array(off_ptr, shape=(int(n) + 1,))
    blob = C.string_at(blob_ptr, int(off[-1]))
    return [blob[int(off[i]):int(off[i + 1])].decode("latin-1") for i in range(int(n))]


def _to_tables(t: _MsTables) -> Tables:
    vm = dict(read=_arr(t.vm_read, t.n_vm, np.uint32), unitig=_arr(t.vm_unitig, t.n_vm, np.uint32),
              nr0=_arr(t.vm_nr0, t.n_vm, np.int32), nr1=_arr(t.vm_nr1, t.n_vm, np.int32),
              ir0=_arr(t.vm_ir0, t.n_vm, np.int32), ir1=_arr(t.vm_ir1, t.n_vm, np.int32),
              rratio=_arr(t.vm_rratio, t.n_vm, np.float64), dir=_arr(t.vm_dir, t.n_vm, np.uint8),
              prim=_arr(t.vm_prim, t.n_vm, np.uint8), score=_arr(t.vm_score, t.n_vm, np.uint64),
              line=_arr(t.vm_line, t.n_vm, np.uint64))
    em = dict(unitig=_arr(t.em_unitig, t.n_em, np.uint32), ov0=_arr(t.em_ov0, t.n_em, np.int32),
              ov1=_arr(t.em_ov1, t.n_em, np.int32), score=_arr(t.em_score, t.n_em, np.float64),
              dir=_arr(t.em_dir, t.n_em, np.uint8), prim=_arr(t.em_prim, t.n_em, np.uint8),
              line=_arr(t.em_line, t.n_em, np.uint64))
    return Tables(read_names=_names(t.read_name_off, t.read_names, t.n_reads),
                  read_nlen=_arr(t.read_nlen, t.n_reads, np.uint64), read_line=_arr(t.read_line, t.n_reads, np.uint64),
                  unitig_names=_names(t.unitig_name_off, t.unitig_names, t.n_unitigs), vm=vm,
                  edge_first=_arr(t.edge_first, t.n_edges, np.uint32),
                  edge_second=_arr(t.edge_second, t.n_edges, np.uint32),
                  edge_em_begin=_arr(t.edge_em_begin, t.n_edges + 1, np.uint64), em=em,
                  times=dict(index=t.t_index, read=t.t_read, edges=t.t_edges))


def tables_equal(a: Tables, b: Tables, names: bool = True) -> list:
    """Bit-exact comparison (doubles compared as uint64 bit patterns).  Returns a list of differing fields."""
    bad = []
    if names and a.read_names != b.read_names:
        bad.append("read_names")
    if names and a.unitig_names != b.unitig_names:
        bad.append("unitig_names")
    for f in ("read_nlen", "read_line", "edge_first", "edge_second", "edge_em_begin"):
        if not np.array_equal(getattr(a, f), getattr(b, f)):
            bad.append(f)
    for grp in ("vm", "em"):
        da, db = getattr(a, grp), getattr(b, grp)
        for k in da:
            x, y = da[k], db[k]
            if x.dtype == np.float64:
                x, y = x.view(np.uint64), np.asarray(y, np.float64).view(np.uint64)
            if x.shape != y.shape or not np.array_equal(x, y):
                bad.append(f"{grp}.{k}")
    return bad


def tables_digest(t: Tables) -> str:
    """sha256 over every table in canonical order — the golden fingerprint committed for the big configs."""
    import hashlib
    h = hashlib.sha256()
    for nm in (t.read_names, t.unitig_names):
        h.update(("\n".join(nm) + "\0").encode("latin-1"))
    for a in (t.read_nlen, t.read_line, t.edge_first, t.edge_second, t.edge_em_begin):
        h.update(np.ascontiguousarray(a).tobytes())
    for grp in (t.vm, t.em):
        for k in sorted(grp):
            h.update(k.encode())
            h.update(np.ascontiguousarray(grp[k]).tobytes())
    return h.hexdigest()


def _mix64(x: np.ndarray) -> np.ndarray:
    """splitmix64's finaliser on uint64 arrays (wrap-around arithmetic) — the mix of include/msgpu.h, msgpu_digest."""
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xbf58476d1ce4e5b9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94d049bb133111eb)
        x ^= x >> np.uint64(31)
    return x


def tables_sum_digest(t: Tables) -> dict:
    """The order-independent digest of include/msgpu.h (msgpu_digest), restated in numpy on the checker's tables:
    what msgpu_result_digest must return for the same construction, and what the digests of the slices of a multi-GPU
    build (or the chunks of a chunked one) must add up to mod 2^64."""
    u64 = np.uint64
    E = int(t.edge_first.size)
    b = t.edge_em_begin.astype(np.int64)
    n_per = np.diff(b)
    first = np.repeat(t.edge_first.astype(u64), n_per)
    second = np.repeat(t.edge_second.astype(u64), n_per)
    em = t.em
    flags = (em["dir"].astype(u64) & u64(1)) | ((em["prim"].astype(u64) & u64(1)) << u64(1))
    with np.errstate(over="ignore"):
        h = _mix64(flags + u64(0x9e3779b97f4a7c15))
        h = _mix64(h ^ np.ascontiguousarray(em["score"], np.float64).view(u64))
        h = _mix64(h ^ ((em["ov1"].astype(np.int64).astype(u64) & u64(0xffffffff)) << u64(32)
                        | (em["line"].astype(u64) & u64(0xffffffff))))
        h = _mix64(h ^ (em["unitig"].astype(u64) << u64(32) | (em["ov0"].astype(np.int64).astype(u64) & u64(0xffffffff))))
        h = _mix64(h ^ (first << u64(32) | second))
        em_sum = int(h.sum(dtype=u64)) if h.size else 0
        key = t.edge_first.astype(u64) << u64(32) | t.edge_second.astype(u64)
        he = _mix64(_mix64(n_per.astype(u64) + u64(0xd1b54a32d192ed03)) ^ key)
        edge_sum = int(he.sum(dtype=u64)) if E else 0
    return dict(n_edges=E, n_edge_matches=int(b[-1]) if b.size else 0, edge_sum=edge_sum, em_sum=em_sum,
                first_key=int(key[0]) if E else (1 << 64) - 1, last_key=int(key[-1]) if E else 0, violations=0)


class _Checker:
    so_path = ""
    prefix = ""

    def __init__(self):
        if not os.path.exists(self.so_path):
            raise FileNotFoundError(f"{self.so_path} missing: run `make -C oracle` (or __graft_entry__.build())")
        self.lib = C.CDLL(self.so_path)
        self.lib.ms_tables_free.argtypes = [C.POINTER(_MsTables)]
        self.lib.ms_tables_free.restype = None

    @staticmethod
    def _as_path(paf):
        if isinstance(paf, (bytes, bytearray, memoryview)):
            f = tempfile.NamedTemporaryFile(suffix=".paf", delete=False)
            f.write(paf)
            f.close()
            return f.name, True
        return os.fspath(paf), False


class Reference(_Checker):
    """The reference's own BlastFileReader::read() + MatchMap::calculateEdges(), compiled from /root/reference."""
    so_path = REF_SO

    def __init__(self):
        super().__init__()
        L = self.lib
        L.msref_build.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(_MsTables)]
        L.msref_build.restype = C.c_int
        L.msref_strslice.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int, C.c_char_p]
        L.msref_strslice.restype = C.c_int64
        L.msref_revcomp.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p]
        L.msref_revcomp.restype = None
        L.msref_seq_open.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p]
        L.msref_seq_open.restype = C.c_void_p
        L.msref_seq_close.argtypes = [C.c_void_p]
        L.msref_seq_close.restype = None
        L.msref_seq_id.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        L.msref_seq_id.restype = C.c_uint32
        L.msref_seq_full.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_char_p, C.c_uint64]
        L.msref_seq_full.restype = C.c_int64
        L.msref_seq_fetch.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_char_p,
                                      C.c_uint64]
        L.msref_seq_fetch.restype = C.c_int64

    def build(self, paf, threads: int = 1, dump: bool = True) -> Tables:
        path, tmp = self._as_path(paf)
        try:
            t = _MsTables()
            rc = self.lib.msref_build(path.encode(), threads, 1 if dump else 0, C.byref(t))
            if rc:
                raise RuntimeError("reference threw (see stderr)")
            if not dump:
                return Tables([], np.zeros(0), np.zeros(0), [], {"read": np.zeros(int(t.n_vm), np.uint8)},
                              np.zeros(int(t.n_edges), np.uint32), np.zeros(0), np.zeros(0),
                              {"unitig": np.zeros(int(t.n_em), np.uint8)},
                              dict(index=t.t_index, read=t.t_read, edges=t.t_edges, n_reads=int(t.n_reads)))
            out = _to_tables(t)
            self.lib.ms_tables_free(C.byref(t))
            return out
        finally:
            if tmp:
                os.unlink(path)

    def chain(self, paf, wiggle_room: int = 300) -> Orders:
        """Construction with one thread, then the reference's chainingAndOverlaps on every edge."""
        L = self.lib
        L.msref_chain.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(_MsOrders)]
        L.msref_chain.restype = C.c_int
        L.ms_orders_free.argtypes = [C.POINTER(_MsOrders)]
        L.ms_orders_free.restype = None
        path, tmp = self._as_path(paf)
        try:
            o = _MsOrders()
            if L.msref_chain(path.encode(), wiggle_room, C.byref(o)):
                raise RuntimeError("reference threw (see stderr)")
            out = self._orders(o)
            L.ms_orders_free(C.byref(o))
            return out
        finally:
            if tmp:
                os.unlink(path)

    def filters(self, paf, wiggle_room: int = 300) -> dict:
        """Construction + chaining as in `chain`, then the reference's per-edge jobs findContractionEdges,
        findDeletableEdges and computeBitweight (src/main.cpp:416-460, 534-573; sanityCheck is the reference's own
        sc.cpp).  Returns the FILTER_FIELDS arrays (edges in (first, second) order) and `seconds`."""
        L = self.lib
        L.msref_filters.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(_MsFilters)]
        L.msref_filters.restype = C.c_int
        L.ms_filters_free.argtypes = [C.POINTER(_MsFilters)]
        L.ms_filters_free.restype = None
        path, tmp = self._as_path(paf)
        try:
            f = _MsFilters()
            if L.msref_filters(path.encode(), wiggle_room, C.byref(f)):
                raise RuntimeError("reference threw (see stderr)")
            E = int(f.n_edges)
            out = dict(contraction_order=_arr(f.contraction_order, E, np.int32), deletable=_arr(f.deletable, E, np.uint8),
                       n_kept_orders=_arr(f.n_kept_orders, E, np.uint32), weight=_arr(f.weight, E, np.uint64),
                       consensus=_arr(f.consensus, E, np.uint8), seconds=float(f.t_filters))
            L.ms_filters_free(C.byref(f))
            return out
        finally:
            if tmp:
                os.unlink(path)

    def filters_gpu(self, paf, wiggle_room: int = 300) -> dict:
        """Like `filters`, but construction, chaining and the per-edge jobs run on the GPU and are applied to the
        reference's own Graph / Edge objects through the host shim (findContractionEdges, findDeletableEdges,
        computeBitweights of muchsalsa_b200/host/ms_gpu_shim.cpp); dumped from those objects."""
        L = self.lib
        L.msref_filters_gpu.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(_MsFilters), C.c_char_p, C.c_uint64]
        L.msref_filters_gpu.restype = C.c_int
        L.ms_filters_free.argtypes = [C.POINTER(_MsFilters)]
        L.ms_filters_free.restype = None
        path, tmp = self._as_path(paf)
        try:
            f = _MsFilters()
            err = C.create_string_buffer(512)
            rc = L.msref_filters_gpu(path.encode(), wiggle_room, C.byref(f), err, 512)
            if rc:
                raise RuntimeError(f"rc={rc}: {err.value.decode(errors='replace')}")
            E = int(f.n_edges)
            out = dict(contraction_order=_arr(f.contraction_order, E, np.int32), deletable=_arr(f.deletable, E, np.uint8),
                       n_kept_orders=_arr(f.n_kept_orders, E, np.uint32), weight=_arr(f.weight, E, np.uint64),
                       consensus=_arr(f.consensus, E, np.uint8))
            L.ms_filters_free(C.byref(f))
            return out
        finally:
            if tmp:
                os.unlink(path)

    def chain_gpu(self, paf, wiggle_room: int = 300) -> Orders:
        """Construction AND chaining by the GPU path through the host shim (constructGraph + chainAndOverlaps of
        muchsalsa_b200/host/ms_gpu_shim.cpp) into the reference's own Graph / Edge objects, dumped like `chain`."""
        L = self.lib
        L.msref_chain_gpu.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(_MsOrders), C.c_char_p, C.c_uint64]
        L.msref_chain_gpu.restype = C.c_int
        L.ms_orders_free.argtypes = [C.POINTER(_MsOrders)]
        L.ms_orders_free.restype = None
        path, tmp = self._as_path(paf)
        try:
            o = _MsOrders()
            err = C.create_string_buffer(512)
            rc = L.msref_chain_gpu(path.encode(), wiggle_room, C.byref(o), err, 512)
            if rc:
                raise RuntimeError(f"rc={rc}: {err.value.decode(errors='replace')}")
            out = self._orders(o)
            L.ms_orders_free(C.byref(o))
            return out
        finally:
            if tmp:
                os.unlink(path)

    @staticmethod
    def _orders(o) -> Orders:
        E, N, I = int(o.n_edges), int(o.n_orders), int(o.n_ids)
        return Orders(_arr(o.edge_shadow, E, np.uint8), _arr(o.edge_order_begin, E + 1, np.uint64),
                      _arr(o.ord_start_is_first, N, np.uint8), _arr(o.ord_left, N, np.float64),
                      _arr(o.ord_right, N, np.float64), _arr(o.ord_contained, N, np.uint8),
                      _arr(o.ord_direction, N, np.uint8), _arr(o.ord_primary, N, np.uint8),
                      _arr(o.ord_score, N, np.uint64), _arr(o.ord_ids_begin, N + 1, np.uint64),
                      _arr(o.ord_ids, I, np.uint32), float(o.t_chain))

    @property
    def has_gpu_shim(self) -> bool:
        return hasattr(self.lib, "msref_build_gpu")

    def build_gpu(self, paf) -> Tables:
        """The drop-in check: the reference's own Graph / MatchMap / Registry objects filled by the GPU path through
        muchsalsa_b200/host/ms_gpu_shim.cpp, dumped by the same code as `build`."""
        L = self.lib
        L.msref_build_gpu.argtypes = [C.c_char_p, C.POINTER(_MsTables), C.c_char_p, C.c_uint64]
        L.msref_build_gpu.restype = C.c_int
        path, tmp = self._as_path(paf)
        try:
            t = _MsTables()
            err = C.create_string_buffer(512)
            rc = L.msref_build_gpu(path.encode(), C.byref(t), err, 512)
            if rc:
                raise RuntimeError(f"rc={rc}: {err.value.decode(errors='replace')}")
            out = _to_tables(t)
            out.times["shim_total"] = t.t_edges
            self.lib.ms_tables_free(C.byref(t))
            return out
        finally:
            if tmp:
                os.unlink(path)

    def seq_fetch_gpu(self, unitigs_path, nanopore_path, which: int, name: str, l: int, r: int, direction: bool,
                      cap: int = 1 << 22):
        L = self.lib
        L.msref_seq_fetch_gpu.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int,
                                          C.c_char_p, C.c_uint64]
        L.msref_seq_fetch_gpu.restype = C.c_int64
        out = C.create_string_buffer(cap)
        n = L.msref_seq_fetch_gpu(os.fspath(unitigs_path).encode(), os.fspath(nanopore_path).encode(), which,
                                  name.encode("latin-1"), l, r, 1 if direction else 0, out, cap)
        if n == -2:
            raise RuntimeError("GPU sequence store could not be set up")
        return None if n < 0 else out.raw[:n]

    def strslice(self, s: bytes, i: int, j: int):
        out = C.create_string_buffer(len(s) + 2)
        n = self.lib.msref_strslice(s, len(s), i, j, out)
        return None if n < 0 else out.raw[:n]

    def revcomp(self, s: bytes) -> bytes:
        out = C.create_string_buffer(len(s) + 1)
        self.lib.msref_revcomp(s, len(s), out)
        return out.raw[:len(s)]

    def seq_open(self, nanopore_path, unitigs_path, prime_reads=None, prime_unitigs=None):
        enc = lambda names: None if names is None else ("\n".join(names)).encode("latin-1")
        h = self.lib.msref_seq_open(os.fspath(nanopore_path).encode(), os.fspath(unitigs_path).encode(),
                                    enc(prime_reads), enc(prime_unitigs))
        if not h:
            raise RuntimeError("reference could not open the sequence files")
        return _RefSeq(self.lib, h)


class _RefSeq:
    def __init__(self, lib, h):
        self.lib, self.h = lib, h

    def close(self):
        if self.h:
            self.lib.msref_seq_close(self.h)
            self.h = None

    def id(self, which: int, name: str) -> int:
        return self.lib.msref_seq_id(self.h, which, name.encode("latin-1"))

    def full(self, which: int, id_: int, cap: int = 1 << 22):
        out = C.create_string_buffer(cap)
        n = self.lib.msref_seq_full(self.h, which, id_, out, cap)
        return None if n < 0 else out.raw[:n]

    def fetch(self, which: int, id_: int, l: int, r: int, direction: bool, cap: int = 1 << 22):
        out = C.create_string_buffer(cap)
        n = self.lib.msref_seq_fetch(self.h, which, id_, l, r, 1 if direction else 0, out, cap)
        return None if n < 0 else out.raw[:n]


def reference_cli(paf_path: str, threads: int, timeout: float = 3600.0) -> dict:
    """Runs the reference as a PROCESS; returns its timing line, or {'crashed': rc} (it races for threads >= 2)."""
    import json
    p = subprocess.run([REF_CLI, paf_path, str(threads)], capture_output=True, text=True, timeout=timeout)
    if p.returncode != 0:
        return {"crashed": p.returncode, "threads": threads}
    return json.loads(p.stdout.strip().splitlines()[-1])


class Oracle(_Checker):
    """Plain-C restatement (oracle/ms_oracle.c)."""
    so_path = ORACLE_SO

    def __init__(self):
        super().__init__()
        L = self.lib
        L.mso_build.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(_MsTables)]
        L.mso_build.restype = C.c_int
        L.mso_strslice.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int, C.c_char_p]
        L.mso_strslice.restype = C.c_int64
        L.mso_revcomp.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p]
        L.mso_revcomp.restype = None
        L.mso_seq_index.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.POINTER(C.POINTER(C.c_uint64)),
                                    C.POINTER(C.c_uint64)]
        L.mso_seq_index.restype = C.c_int
        L.mso_free.argtypes = [C.c_void_p]
        L.mso_free.restype = None
        L.mso_seq_extract.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_char_p]
        L.mso_seq_extract.restype = C.c_int64

    def build(self, paf) -> Tables:
        if not isinstance(paf, (bytes, bytearray)):
            with open(paf, "rb") as f:
                paf = f.read()
        t = _MsTables()
        rc = self.lib.mso_build(bytes(paf), len(paf), C.byref(t))
        if rc:
            raise RuntimeError(f"oracle: invalid PAF (rc={rc})")
        out = _to_tables(t)
        self.lib.ms_tables_free(C.byref(t))
        return out

    def strslice(self, s: bytes, i: int, j: int):
        out = C.create_string_buffer(len(s) + 2)
        n = self.lib.mso_strslice(s, len(s), i, j, out)
        return None if n < 0 else out.raw[:n]

    def revcomp(self, s: bytes) -> bytes:
        out = C.create_string_buffer(len(s) + 1)
        self.lib.mso_revcomp(s, len(s), out)
        return out.raw[:len(s)]

    def seq_index(self, data: bytes, fastq: bool):
        """[(name, byte offset, byte length incl. newlines)] in file order (SequenceAccessor.cpp:143-231)."""
        p = C.POINTER(C.c_uint64)()
        n = C.c_uint64()
        rc = self.lib.mso_seq_index(data, len(data), 1 if fastq else 0, C.byref(p), C.byref(n))
        if rc:
            raise RuntimeError("oracle: sequence index failed")
        a = np.ctypeslib.as_array(p, shape=(int(n.value), 4)).copy() if n.value else np.zeros((0, 4), np.uint64)
        self.lib.mso_free(p)
        return [(data[int(r[0]):int(r[0] + r[1])].decode("latin-1"), int(r[2]), int(r[3])) for r in a]

    def seq_extract(self, data: bytes, off: int, length: int) -> bytes:
        """Whole record with isspace bytes removed (getSequenceFromFile, SequenceAccessor.cpp:54-69)."""
        out = C.create_string_buffer(length + 1)
        n = self.lib.mso_seq_extract(data, len(data), off, length, out)
        if n < 0:
            raise RuntimeError("Failed to read sequence.")
        return out.raw[:n]

    def seq_fetch(self, full: bytes, l: int, r: int, direction: bool):
        """get{Illumina,Nanopore}Sequence on an already extracted record (SequenceUtils.cpp:63-85)."""
        s = self.strslice(full, l, r + 1)
        if s is None:
            return None
        return s if direction else self.revcomp(s)
