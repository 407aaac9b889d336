"""Persisted segment layout ("next" row 8f-1).

`write_segment` is the host-side mirror of what JVectorWriter puts on disk (JVectorWriter.java:134-165,
299-300, 383-433, 469-510, 528-540, 573-577): Lucene CodecUtil framing, the meta records with the doc
map, and one data file per field holding the OnDiskGraphIndex bytes followed by the PQVectors blob.
`SegmentFiles` reads them back through libjvgpu's loader (csrc/jv_segment.cu) — the same C-ABI calls the
Java FieldEntry constructor would make — into decoded arrays or straight into a device index.

Lucene framing and the meta record are verified from the reference tree.  The byte layout INSIDE the two
jVector blobs belongs to the un-vendored jar (jvector 4.0.0-rc.9) and is restated from its published
format; writer and loader agree with each other, neither has been checked against a file
produced by the real plugin (no JVM here) — DESIGN.md section 8.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import zlib
from pathlib import Path
from typing import Dict, Optional

import numpy as np

from . import native as N

CODEC_MAGIC = 0x3FD76C17                      # CodecUtil.CODEC_MAGIC
FOOTER_MAGIC = ~CODEC_MAGIC & 0xFFFFFFFF
META_CODEC_NAME = "JVectorVectorsFormatMeta"          # JVectorFormat.java:23
VECTOR_INDEX_CODEC_NAME = "JVectorVectorsFormatIndex"  # JVectorFormat.java:24
META_EXTENSION = "meta-jvector"                        # JVectorFormat.java:27
VECTOR_INDEX_EXTENSION = "data-jvector"                # JVectorFormat.java:28
VERSION_START, VERSION_WITH_QUANTIZATION_TYPE, VERSION_CURRENT = 0, 1, 1   # JVectorFormat.java:31-33
QUANTIZATION_TYPE_NONE, QUANTIZATION_TYPE_PQ, QUANTIZATION_TYPE_NVQ_INLINE = 0, 1, 2
DOC_MAP_VERSION = 1
# jVector blobs [restated, SURVEY B.2]
GRAPH_MAGIC, GRAPH_FOOTER_MAGIC, PQ_MAGIC = 0xFFFF0D61, 0x4A564244, 0x75EC4012
FEATURE_INLINE_VECTORS = 1
PQ_UNWEIGHTED = -1.0

FLAG_FLOATS_BIG_ENDIAN, FLAG_VERIFY_DATA_CRC, FLAG_LENIENT_MAGIC = 1, 2, 4


# ---- Lucene DataOutput subset ------------------------------------------------------------------
def _vint(v: int) -> bytes:
    v &= 0xFFFFFFFF                      # writeVInt of a negative int takes 5 bytes
    out = bytearray()
    while v & ~0x7F:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _vlong(v: int) -> bytes:
    if v < 0:
        raise ValueError("writeVLong of a negative value")
    out = bytearray()
    while v & ~0x7F:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _vints(a: np.ndarray) -> bytes:
    """vectorised writeVInt of an int32 array (the doc map can hold millions of entries)."""
    v = a.astype(np.int64) & 0xFFFFFFFF
    groups = np.stack([(v >> (7 * i)) & 0x7F for i in range(5)], axis=1).astype(np.uint8)
    nbytes = np.ones(v.shape[0], dtype=np.int64)
    for i in range(1, 5):
        nbytes[v >= (1 << (7 * i))] = i + 1
    cont = np.arange(5)[None, :] < (nbytes[:, None] - 1)
    groups |= (cont.astype(np.uint8) << 7)
    keep = np.arange(5)[None, :] < nbytes[:, None]
    return groups[keep].tobytes()


def index_header(codec: str, version: int, segment_id: bytes, suffix: str) -> bytes:
    """CodecUtil.writeIndexHeader: BE magic, codec string, BE version, 16-byte id, suffix."""
    if len(segment_id) != 16:
        raise ValueError("segment id must be 16 bytes")
    name, sfx = codec.encode("utf-8"), suffix.encode("utf-8")
    return struct.pack(">I", CODEC_MAGIC) + _vint(len(name)) + name + struct.pack(">i", version) + segment_id + bytes([len(sfx)]) + sfx


def footer(crc_so_far: int) -> bytes:
    """CodecUtil.writeFooter: ~magic, algorithm 0, CRC-32 (as a long) of everything before the CRC."""
    head = struct.pack(">II", FOOTER_MAGIC, 0)
    return head + struct.pack(">Q", zlib.crc32(head, crc_so_far) & 0xFFFFFFFF)


class _Out:
    """IndexOutput: sequential writes with a running CRC-32 and a file pointer."""

    def __init__(self, path):
        self.f = open(path, "wb")
        self.crc = 0
        self.pos = 0

    def write(self, b) -> None:
        b = memoryview(b).cast("B") if not isinstance(b, (bytes, bytearray)) else b
        self.f.write(b)
        self.crc = zlib.crc32(b, self.crc)
        self.pos += len(b)

    def close_with_footer(self) -> None:
        self.f.write(footer(self.crc))
        self.f.close()


def segment_file_name(segment_name: str, suffix: str, ext: str) -> str:
    """IndexFileNames.segmentFileName."""
    return f"{segment_name}_{suffix}.{ext}" if suffix else f"{segment_name}.{ext}"


def field_data_file_name(segment_name: str, suffix: str, field: str) -> str:
    """baseDataFileName + "_" + field + "." + VECTOR_INDEX_EXTENSION, JVectorWriter.java:383-384."""
    return f"{segment_name}_{suffix}_{field}.{VECTOR_INDEX_EXTENSION}"


def _graph_header(version: int, n: int, dim: int, entry: int, degree: int) -> bytes:
    h = struct.pack("<Ii", GRAPH_MAGIC, version) + struct.pack("<iiii", n, dim, entry, degree)
    if version >= 4:
        h += struct.pack("<ii", n, 1) + struct.pack("<ii", n, degree)      # idUpperBound, one layer (hierarchy off)
    h += struct.pack("<I", FEATURE_INLINE_VECTORS) + struct.pack("<i", dim)  # feature set, InlineVectors header
    return h


def write_field_data(path, fd, segment_id: bytes, suffix: str, graph_version: int = 5, float_order: str = "<",
                     chunk_nodes: int = 65536):
    """writeGraph / writeFullPrecisionGraph, JVectorWriter.java:374-433,469-510.  Returns
    (index_offset, index_length, pq_offset, pq_length)."""
    out = _Out(path)
    out.write(index_header(VECTOR_INDEX_CODEC_NAME, VERSION_CURRENT, segment_id, suffix))
    start = out.pos
    vec = np.ascontiguousarray(fd.vectors, dtype=np.float32)
    n, dim = vec.shape
    adj = np.ascontiguousarray(fd.adjacency, dtype=np.int32).reshape(n, -1)
    degree = adj.shape[1]
    header = _graph_header(graph_version, n, dim, int(fd.entry_node), degree)
    out.write(header)
    rec_dt = np.dtype([("ord", "<i4"), ("vec", float_order + "f4", (dim,)), ("deg", "<i4"), ("nb", "<i4", (degree,))])
    for lo in range(0, n, chunk_nodes):                                    # layer-0 records, ordinal order
        hi = min(n, lo + chunk_nodes)
        rec = np.empty(hi - lo, dtype=rec_dt)
        rec["ord"] = np.arange(lo, hi, dtype=np.int32)
        rec["vec"] = vec[lo:hi]
        a = adj[lo:hi]
        live = a >= 0
        order = np.argsort(~live, axis=1, kind="stable")                   # neighbours first, -1 padding after
        rec["nb"] = np.take_along_axis(a, order, axis=1)
        rec["deg"] = live.sum(axis=1)
        out.write(rec.tobytes())
    if graph_version >= 5:                                                 # header repeated as a footer
        header_offset = out.pos
        out.write(header)
        out.write(struct.pack("<qI", header_offset, GRAPH_FOOTER_MAGIC))
    end_graph = out.pos
    pq_offset = pq_length = 0
    if fd.pq_codes is not None and fd.pq_m > 0:
        pq_offset = end_graph
        m, k = int(fd.pq_m), int(fd.pq_k)
        base, rem = divmod(dim, m)
        sizes = [base + (1 if i < rem else 0) for i in range(m)]
        offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(int)
        out.write(struct.pack("<Ii", PQ_MAGIC, graph_version))             # ProductQuantization.write
        g = fd.pq_global_centroid
        if g is None:
            out.write(struct.pack("<i", 0))
        else:
            out.write(struct.pack("<i", dim))
            out.write(np.ascontiguousarray(g, dtype=np.float32).astype(float_order + "f4").tobytes())
        out.write(struct.pack("<f", PQ_UNWEIGHTED))
        out.write(struct.pack("<i", m))
        for s, o in zip(sizes, offs):
            out.write(struct.pack("<ii", int(s), int(o)))
        out.write(struct.pack("<i", k))
        out.write(np.ascontiguousarray(fd.pq_codebooks, dtype=np.float32).reshape(-1).astype(float_order + "f4").tobytes())
        codes = np.ascontiguousarray(fd.pq_codes, dtype=np.uint8).reshape(n, m)
        out.write(struct.pack("<ii", n, m))                                # PQVectors.write
        out.write(codes.tobytes())
        pq_length = out.pos - pq_offset
    out.close_with_footer()
    return start, end_graph - start, pq_offset, pq_length


def write_segment(segment, directory, segment_name: str = "_0", suffix: str = "JVector_0",
                  segment_id: Optional[bytes] = None, field_numbers: Optional[Dict[str, int]] = None,
                  version: int = VERSION_CURRENT, graph_version: int = 5, float_order: str = "<") -> Dict[str, Path]:
    """JVectorWriter ctor + writeField* + finish.  Returns {"meta": path, "<field>": data path, ...}.
    The neighbours-score-cache file (merge-only, JVectorWriter.java:339-363) is not written."""
    directory = Path(directory)
    directory.mkdir(parents=True, exist_ok=True)
    segment_id = segment_id or os.urandom(16)
    field_numbers = field_numbers or {name: i for i, name in enumerate(segment.fields)}
    paths: Dict[str, Path] = {}
    meta_path = directory / segment_file_name(segment_name, suffix, META_EXTENSION)
    meta = _Out(meta_path)
    meta.write(index_header(META_CODEC_NAME, version, segment_id, suffix))
    vestigial = _Out(directory / segment_file_name(segment_name, suffix, VECTOR_INDEX_EXTENSION))   # :140-165: header + footer only
    vestigial.write(index_header(VECTOR_INDEX_CODEC_NAME, version, segment_id, suffix))
    vestigial.close_with_footer()
    for name, fd in segment.fields.items():
        number = field_numbers[name]
        path = directory / field_data_file_name(segment_name, suffix, name)
        io, il, po, pl = write_field_data(path, fd, segment_id, suffix, graph_version, float_order)
        paths[name] = path
        qtype = QUANTIZATION_TYPE_PQ if pl else QUANTIZATION_TYPE_NONE
        meta.write(struct.pack("<i", number))                              # writeField, :299-300
        rec = struct.pack("<iii", number, 1, fd.similarity.meta_ord)       # toOutput, :528-540 (VectorEncoding.FLOAT32 = 1; MIP -> 1)
        rec += _vint(fd.vectors.shape[1]) + _vlong(io) + _vlong(il) + _vlong(po) + _vlong(pl)
        if version >= VERSION_WITH_QUANTIZATION_TYPE:
            rec += bytes([qtype])
        rec += struct.pack("<f", 0.0)                                       # degreeOverflow: never set by the builder
        ords = fd.doc_map.graph_node_ids_to_doc_ids
        rec += struct.pack("<i", DOC_MAP_VERSION) + _vint(len(ords)) + _vint(fd.doc_map.max_docs)
        meta.write(rec)
        meta.write(_vints(ords))                                            # GraphNodeIdToDocMap.toOutput, :169-176
    meta.write(struct.pack("<i", -1))                                       # finish(), :573-577
    meta.close_with_footer()
    paths["meta"] = meta_path
    return paths


# ---- reading through the C-ABI loader -----------------------------------------------------------
class SegmentFiles:
    """jv_segment_open / jv_segment_load_field / jv_segment_index_create (include/jvgpu.h)."""

    def __init__(self, meta_path, flags: int = 0):
        self._h = C.c_void_p()
        self.flags = flags
        N.check(N.load().jv_segment_open(str(meta_path).encode(), flags, C.addressof(self._h)))
        cnt = C.c_int32(0)
        N.check(N.load().jv_segment_field_count(self._h, C.addressof(cnt)))
        self.metas = []
        for i in range(cnt.value):
            m = N.FieldMeta()
            m.struct_size = C.sizeof(N.FieldMeta)
            N.check(N.load().jv_segment_field_meta(self._h, i, C.addressof(m)))
            self.metas.append(m)

    def field_index_of(self, field_number: int) -> int:
        for i, m in enumerate(self.metas):
            if m.field_number == field_number:
                return i
        raise KeyError(field_number)

    def set_lucene_similarity(self, i: int, lucene_similarity: int) -> None:
        """FieldInfo.getVectorSimilarityFunction() of field i (needed for MAXIMUM_INNER_PRODUCT, stored as DOT_PRODUCT)."""
        N.check(N.load().jv_segment_set_lucene_similarity(self._h, i, int(lucene_similarity)))
        self.metas[i].similarity = int(lucene_similarity)

    def doc_map(self, i: int) -> np.ndarray:
        out = np.empty(self.metas[i].graph_nodes, dtype=np.int32)
        N.check(N.load().jv_segment_field_doc_map(self._h, i, out.ctypes.data if out.size else None, out.size))
        return out

    def load_field(self, i: int, field_data_path, flags: int = 0) -> dict:
        """Decoded host arrays of one field (copies; the native buffers are released before returning)."""
        lib = N.load()
        h = C.c_void_p()
        N.check(lib.jv_segment_load_field(self._h, i, str(field_data_path).encode(), flags, C.addressof(h)))
        try:
            d = N.IndexDesc()
            N.check(lib.jv_field_data_desc(h, C.addressof(d)))
            n, dim, r = int(d.n), int(d.dim), int(d.max_degree)

            def arr(ptr, ctype, count, shape):
                if not ptr or count == 0:
                    return None
                return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(count,)).reshape(shape).copy()

            out = {
                "similarity": int(d.similarity), "n": n, "dim": dim, "max_degree": r, "entry_node": int(d.entry_node),
                "max_doc": int(d.max_doc),
                "adjacency": arr(d.adjacency, C.c_int32, n * r, (n, r)),
                "vectors": arr(d.vectors, C.c_float, n * dim, (n, dim)),
                "ord_to_doc": arr(d.ord_to_doc, C.c_int32, n, (n,)),
                "pq_m": int(d.pq_m), "pq_k": int(d.pq_k),
                "pq_codebooks": arr(d.pq_codebooks, C.c_float, int(d.pq_k) * dim, (-1,)),
                "pq_global_centroid": arr(d.pq_global_centroid, C.c_float, dim, (dim,)),
                "pq_codes": arr(d.pq_codes, C.c_uint8, n * int(d.pq_m), (n, max(int(d.pq_m), 1))),
            }
            return out
        finally:
            lib.jv_field_data_free(h)

    def index_create(self, i: int, field_data_path, device: int = 0, index_flags: int = 0, load_flags: int = 0):
        """FieldEntry constructor in one native call; returns a GpuIndex owning the device index."""
        from .index import GpuIndex
        h = C.c_void_p()
        N.check(N.load().jv_segment_index_create(self._h, i, str(field_data_path).encode(), device, index_flags, load_flags,
                                                 C.addressof(h)))
        m = self.metas[i]
        return GpuIndex.from_handle(h, similarity=m.similarity, n=m.graph_nodes, dim=m.dim, max_doc=m.max_doc, device=device,
                                    has_pq=m.pq_length > 0)

    def close(self):
        if self._h:
            N.load().jv_segment_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def check_integrity(path) -> None:
    """CodecUtil.checksumEntireFile (JVectorReader.checkIntegrity, JVectorReader.java:87-99)."""
    N.check(N.load().jv_file_check_integrity(str(path).encode()))
