"""ctypes view of libmsgpu.so — exactly the symbols include/msgpu.h declares.

There is no CPU fallback: if the CUDA library is missing or no B200 is visible, loading / msgpu_create fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# MSGPU_LIB: development only — another build of the same library (A/B of compile-time constants on the GPU box)
LIB_PATH = os.environ.get("MSGPU_LIB") or os.path.join(HERE, "libmsgpu.so")

u8p, u32p, i32p, u64p, f64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_int32, C.c_uint64, C.c_double))


class Counts(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("n_lines", "n_reads", "n_unitigs", "n_matches", "n_pairs", "n_edges", "n_edge_matches")]


class Tables(C.Structure):
    _fields_ = [
        ("read_nlen", i32p), ("read_line", u32p), ("read_name_off", u64p), ("read_names", C.c_void_p),
        ("unitig_name_off", u64p), ("unitig_names", C.c_void_p),
        ("m_unitig", u32p), ("m_read", u32p), ("m_nr0", i32p), ("m_nr1", i32p), ("m_ir0", i32p), ("m_ir1", i32p),
        ("m_rratio", f64p), ("m_score", i32p), ("m_line", u32p), ("m_flags", u8p),
        ("edge_first", u32p), ("edge_second", u32p), ("edge_em_begin", u64p),
        ("em_unitig", u32p), ("em_ov0", i32p), ("em_ov1", i32p), ("em_score", f64p), ("em_line", u32p),
        ("em_flags", u8p),
    ]


class Digest(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("n_edges", "n_edge_matches", "edge_sum", "em_sum", "first_key", "last_key", "violations")]


class Chunk(C.Structure):
    _fields_ = [("index", C.c_uint64), ("n_edges", C.c_uint64), ("n_edge_matches", C.c_uint64),
                ("edge_base", C.c_uint64), ("em_base", C.c_uint64), ("last", C.c_int),
                ("edge_first", u32p), ("edge_second", u32p), ("edge_em_begin", u64p), ("em_unitig", u32p),
                ("em_ov0", i32p), ("em_ov1", i32p), ("em_score", f64p), ("em_line", u32p), ("em_flags", u8p)]


class Coverage(C.Structure):
    _fields_ = [("n_blocks", C.c_uint64), ("block_unitig", u32p), ("block_n_reads", u32p), ("block_len", u32p),
                ("block_max_cov", u32p)]


class Fragments(C.Structure):
    _fields_ = [("n_blocks", C.c_uint64), ("n_fragments", C.c_uint64), ("block_outlier", u8p), ("frag_block", u32p),
                ("frag_start", i32p), ("frag_end", i32p)]


class Timings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d_ms", "ingest_ms", "pair_ms", "merge_ms", "d2h_ms", "total_ms")]


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_uint32), ("total_ms", C.c_float)]


class MatchSoA(C.Structure):
    _fields_ = [("n", C.c_uint64), ("unitig", C.c_void_p), ("read", C.c_void_p), ("ir0", C.c_void_p),
                ("ir1", C.c_void_p), ("nr0", C.c_void_p), ("nr1", C.c_void_p), ("score", C.c_void_p),
                ("line", C.c_void_p), ("flags", C.c_void_p)]


class ShardHeader(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("magic", "n_lines", "n_reads", "n_unitigs", "n_matches", "read_name_bytes",
                                          "unitig_name_bytes", "blob_bytes")]


class Orders(C.Structure):
    _fields_ = [("n_edges", C.c_uint64), ("n_slots", C.c_uint64), ("edge_shadow", u8p), ("edge_n_orders", u32p),
                ("ord_left", f64p), ("ord_right", f64p), ("ord_score", u64p), ("ord_n_ids", u32p), ("ord_flags", u8p),
                ("ord_ids", u32p)]


class EdgeFilters(C.Structure):
    _fields_ = [("n_edges", C.c_uint64), ("contraction_order", C.POINTER(C.c_int32)), ("deletable", u8p),
                ("n_kept_orders", u32p), ("weight", u64p), ("consensus", u8p)]


class FilterViewHeader(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_edges", "n_orders", "blob_bytes", "reserved")]


class SeqReq(C.Structure):
    _fields_ = [("which", C.c_uint32), ("id", C.c_uint32), ("left", C.c_int32), ("right", C.c_int32),
                ("direction", C.c_uint32)]


# name -> (restype, argtypes); the single source of truth the symbol test checks against include/msgpu.h
SIGNATURES = {
    "msgpu_abi_version": (C.c_int, []),
    "msgpu_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "msgpu_destroy": (None, [C.c_void_p]),
    "msgpu_last_error": (C.c_char_p, [C.c_void_p]),
    "msgpu_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "msgpu_build_from_paf": (C.c_int, [C.c_void_p, C.c_char_p]),
    "msgpu_build_from_paf_buffer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "msgpu_build_from_paf_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "msgpu_anchor_graph_from_paf_buffer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "msgpu_anchor_graph_from_paf_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "msgpu_unitig_coverage_from_paf_buffer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(Coverage)]),
    "msgpu_unitig_coverage_from_paf_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(Coverage)]),
    "msgpu_unitig_coverage_cut": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.POINTER(Fragments)]),
    "msgpu_build_from_records": (C.c_int, [C.c_void_p, C.POINTER(MatchSoA), C.c_uint32, C.c_uint32]),
    "msgpu_shard_ingest_buffer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
    "msgpu_shard_ingest_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
    "msgpu_shard_describe": (C.c_int, [C.c_void_p, C.POINTER(ShardHeader)]),
    "msgpu_shard_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "msgpu_shard_export_peers": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_uint32, C.c_uint64]),
    "msgpu_build_from_shards": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(ShardHeader), C.POINTER(C.c_void_p),
                                          C.c_uint32, C.c_uint32]),
    "msgpu_result_counts": (C.c_int, [C.c_void_p, C.POINTER(Counts)]),
    "msgpu_result_device": (C.c_int, [C.c_void_p, C.POINTER(Tables)]),
    "msgpu_result_host": (C.c_int, [C.c_void_p, C.POINTER(Tables)]),
    "msgpu_set_chunk_tiles": (C.c_int, [C.c_void_p, C.c_uint32]),
    "msgpu_chunk_next": (C.c_int, [C.c_void_p, C.POINTER(Chunk)]),
    "msgpu_chunk_digest": (C.c_int, [C.c_void_p, C.POINTER(Digest)]),
    "msgpu_result_digest": (C.c_int, [C.c_void_p, C.POINTER(Digest)]),
    "msgpu_chain_overlaps": (C.c_int, [C.c_void_p, C.c_uint64]),
    "msgpu_orders_device": (C.c_int, [C.c_void_p, C.POINTER(Orders)]),
    "msgpu_orders_host": (C.c_int, [C.c_void_p, C.POINTER(Orders)]),
    "msgpu_chain_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "msgpu_edge_filters": (C.c_int, [C.c_void_p, C.c_uint64]),
    "msgpu_edge_filters_device": (C.c_int, [C.c_void_p, C.POINTER(EdgeFilters)]),
    "msgpu_edge_filters_host": (C.c_int, [C.c_void_p, C.POINTER(EdgeFilters)]),
    "msgpu_edge_filters_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "msgpu_filter_view_describe": (C.c_int, [C.c_void_p, C.POINTER(FilterViewHeader)]),
    "msgpu_filter_view_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "msgpu_edge_filters_sharded": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(FilterViewHeader),
                                             C.POINTER(C.c_void_p), C.c_uint64]),
    "msgpu_result_timings": (C.c_int, [C.c_void_p, C.POINTER(Timings)]),
    "msgpu_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "msgpu_profile_read": (C.c_int, [C.c_void_p, C.POINTER(KernelStat), C.c_size_t, C.POINTER(C.c_size_t)]),
    "msgpu_launch_count": (C.c_uint64, [C.c_void_p]),
    "msgpu_seq_load": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "msgpu_seq_load_buffers": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]),
    "msgpu_seq_id": (C.c_int, [C.c_void_p, C.c_uint32, C.c_char_p, u32p]),
    "msgpu_seq_batch_bound": (C.c_int, [C.c_void_p, C.POINTER(SeqReq), C.c_size_t, u64p]),
    "msgpu_seq_fetch_batch": (C.c_int, [C.c_void_p, C.POINTER(SeqReq), C.c_size_t, C.c_void_p, C.c_uint64, u64p,
                                        u8p]),
    "msgpu_seq_fetch_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64,
                                               C.c_void_p, C.c_void_p]),
    "msgpu_slice_revcomp_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_size_t, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen libmsgpu.so (RTLD_GLOBAL not needed: cudart is linked statically) and bind every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing — build it with `python -m muchsalsa_b200.build` "
                          "(__graft_entry__.build()). There is no CPU fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError = the library does not export what the header declares
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib
