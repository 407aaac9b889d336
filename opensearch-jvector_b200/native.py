"""ctypes binding of libjvgpu.so — the same symbols the Java codec binds through Panama FFM
(INTEGRATION.md).  Loading fails loudly when the CUDA library is missing: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "lib" / "libjvgpu.so"

JV_OK = 0
ERR_INVALID_ARGUMENT, ERR_CUDA, ERR_OUT_OF_MEMORY, ERR_UNSUPPORTED, ERR_INTERNAL, ERR_CORRUPT = -1, -2, -3, -4, -5, -6
SIM_EUCLIDEAN, SIM_DOT, SIM_COSINE, SIM_MIP = 0, 1, 2, 3
FLAG_FUSED_LAYOUT, FLAG_LUT_F16, FLAG_NO_VECTORS_ON_DEVICE, FLAG_LUT_U8 = 1, 2, 4, 8


class IndexDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("similarity", C.c_int32), ("dim", C.c_int32), ("max_degree", C.c_int32),
        ("n", C.c_int64), ("entry_node", C.c_int32), ("max_doc", C.c_int32),
        ("adjacency", C.c_void_p), ("vectors", C.c_void_p), ("ord_to_doc", C.c_void_p),
        ("pq_m", C.c_int32), ("pq_k", C.c_int32),
        ("pq_codebooks", C.c_void_p), ("pq_global_centroid", C.c_void_p), ("pq_codes", C.c_void_p),
        ("device", C.c_int32), ("flags", C.c_uint32),
        ("nvq_m", C.c_int32), ("nvq_reserved", C.c_int32),
        ("nvq_bytes", C.c_void_p), ("nvq_params", C.c_void_p), ("nvq_global_mean", C.c_void_p),
    ]


class FieldMeta(C.Structure):
    """jv_field_meta: one VectorIndexFieldMetadata record of the meta file (JVectorWriter.java:528-540)."""
    _fields_ = [
        ("struct_size", C.c_int32), ("field_number", C.c_int32), ("vector_encoding", C.c_int32), ("similarity", C.c_int32),
        ("dim", C.c_int32), ("quantization_type", C.c_int32),
        ("index_offset", C.c_int64), ("index_length", C.c_int64), ("pq_offset", C.c_int64), ("pq_length", C.c_int64),
        ("degree_overflow", C.c_float), ("graph_nodes", C.c_int32), ("max_doc", C.c_int32), ("format_version", C.c_int32),
    ]


class QueryStats(C.Structure):
    _fields_ = [("visited", C.c_int32), ("expanded", C.c_int32), ("expanded_base", C.c_int32), ("reranked", C.c_int32)]


class BatchTiming(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("search_ms", C.c_float), ("rerank_ms", C.c_float), ("d2h_ms", C.c_float),
                ("total_ms", C.c_float), ("launches", C.c_int32), ("lut_ms", C.c_float), ("expand_width_used", C.c_int32),
                ("traversal_kernel", C.c_int32)]


class SearchParams(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("k", C.c_int32), ("rerank_k", C.c_int32), ("threshold", C.c_float),
                ("rerank_floor", C.c_float), ("expand_width", C.c_int32), ("accept_bits", C.c_void_p),
                ("accept_stride_words", C.c_int64)]


# every symbol include/jvgpu.h declares: (name, restype, argtypes)
_P, _I32, _I64, _F, _U64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint64
SYMBOLS = {
    "jv_version": (_I32, []),
    "jv_last_error": (C.c_char_p, []),
    "jv_device_count": (_I32, [_P]),
    "jv_host_alloc": (_I32, [_I64, _P]),
    "jv_host_free": (_I32, [_P]),
    "jv_host_register": (_I32, [_P, _I64]),
    "jv_host_unregister": (_I32, [_P]),
    "jv_index_create": (_I32, [_P, _P]),
    "jv_index_destroy": (_I32, [_P]),
    "jv_index_device_bytes": (_I32, [_P, _P]),
    "jv_index_debug_counter": (_I32, [_P, _I32, _P]),
    "jv_search_batch": (_I32, [_P, _P, _I32, _P, _P, _P, _P, _P, _P]),
    "jv_search_batch_dev": (_I32, [_P, _P, _I32, _P, _P, _P, _P, _P, _P]),
    "jv_exact_topk": (_I32, [_P, _P, _I32, _I32, _P, _I64, _P, _P, _P]),
    "jv_exact_topk_dev": (_I32, [_P, _P, _I32, _I32, _P, _I64, _P, _P, _P]),
    "jv_pq_encode": (_I32, [_I32, _P, _I64, _I32, _I32, _I32, _P, _P, _P, _P]),
    "jv_pq_encode_dev": (_I32, [_I32, _P, _I64, _I32, _I32, _I32, _P, _P, _P, _P]),
    "jv_pq_lut": (_I32, [_P, _P, _I32, _P]),
    "jv_pq_lut_q8": (_I32, [_P, _P, _I32, _P, _P]),
    "jv_pq_adc_scores": (_I32, [_P, _P, _I32, _P, _I32, _P]),
    "jv_merge_topk": (_I32, [_I32, _I32, _I32, _I32, _P, _P, _P, _P, _P]),
    "jv_merge_topk_dev": (_I32, [_I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    "jv_merge_topk_stream": (_I32, [_I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    "jv_pq_train": (_I32, [_I32, _P, _I64, _I32, _I32, _I32, _I32, _I32, _U64, _P, _P]),
    "jv_pq_train_dev": (_I32, [_I32, _P, _I64, _I32, _I32, _I32, _I32, _I32, _U64, _P, _P]),
    "jv_pq_decode": (_I32, [_I32, _P, _I64, _I32, _I32, _I32, _P, _P, _P]),
    "jv_pq_decode_dev": (_I32, [_I32, _P, _I64, _I32, _I32, _I32, _P, _P, _P]),
    "jv_graph_build_pq": (_I32, [_I32, _P, _I64, _I32, _I32, _I32, _I32, _P, _P, _I32, _I32, _F, _F, _P, _P]),
    "jv_graph_build_pq_dev": (_I32, [_I32, _P, _I64, _I32, _I32, _I32, _I32, _P, _P, _I32, _I32, _F, _F, _P, _P]),
    "jv_graph_build": (_I32, [_I32, _P, _I64, _I32, _I32, _I32, _I32, _F, _F, _P, _P]),
    "jv_graph_build_dev": (_I32, [_I32, _P, _I64, _I32, _I32, _I32, _I32, _F, _F, _P, _P]),
    "jv_graph_extend": (_I32, [_I32, _P, _I64, _I64, _P, _I32, _I32, _I32, _I32, _I32, _F, _F, _P]),
    "jv_graph_extend_dev": (_I32, [_I32, _P, _I64, _I64, _P, _I32, _I32, _I32, _I32, _I32, _F, _F, _P]),
    "jv_graph_remove_deleted": (_I32, [_I32, _P, _I64, _I32, _I32, _I32, _F, _P, _P, _I32, _P, _P]),
    "jv_graph_remove_deleted_dev": (_I32, [_I32, _P, _I64, _I32, _I32, _I32, _F, _P, _P, _I32, _P, _P]),
    "jv_segment_open": (_I32, [C.c_char_p, C.c_uint32, _P]),
    "jv_segment_close": (_I32, [_P]),
    "jv_segment_field_count": (_I32, [_P, _P]),
    "jv_segment_field_meta": (_I32, [_P, _I32, _P]),
    "jv_segment_field_doc_map": (_I32, [_P, _I32, _P, _I32]),
    "jv_segment_set_lucene_similarity": (_I32, [_P, _I32, _I32]),
    "jv_segment_load_field": (_I32, [_P, _I32, C.c_char_p, C.c_uint32, _P]),
    "jv_field_data_desc": (_I32, [_P, _P]),
    "jv_field_data_free": (_I32, [_P]),
    "jv_segment_index_create": (_I32, [_P, _I32, C.c_char_p, _I32, C.c_uint32, C.c_uint32, _P]),
    "jv_file_check_integrity": (_I32, [C.c_char_p]),
}


class JVectorNativeError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libjvgpu status {status}: {message}")
        self.status = status


_lib = None


def load() -> C.CDLL:
    """dlopen libjvgpu.so and bind every declared symbol; raises if the library or a symbol is missing."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python opensearch-jvector_b200/build.py` "
                "(the product path has no CPU fallback)")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def host_alloc(shape, dtype):
    """numpy array over a page-locked buffer from jv_host_alloc (freed with host_free(arr))."""
    import numpy as np
    dt = np.dtype(dtype)
    n = int(np.prod(shape))
    p = C.c_void_p()
    check(load().jv_host_alloc(max(n * dt.itemsize, 1), C.byref(p)))
    buf = (C.c_byte * max(n * dt.itemsize, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dt, count=n).reshape(shape)
    return arr, p.value


def host_free(ptr: int) -> None:
    check(load().jv_host_free(C.c_void_p(ptr)))


def check(status: int) -> None:
    if status == JV_OK:
        return
    msg = load().jv_last_error().decode("utf-8", "replace")
    if status == ERR_INVALID_ARGUMENT:
        raise ValueError(msg)                      # Java: IllegalArgumentException
    if status == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)             # Java: UnsupportedOperationException
    if status == ERR_OUT_OF_MEMORY:
        raise MemoryError(msg)
    raise JVectorNativeError(status, msg)          # Java: IOException
