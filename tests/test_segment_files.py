"""Segment-file loader: the files JVectorWriter persists (Appendix B) parsed by libjvgpu's host-side
loader.  CPU tests: Lucene framing known answers, write -> load round trips for every layout variant the loader
accepts, and the corruption / unsupported cases the reference raises IOException / UnsupportedOperationException for.
The GPU test opens a written segment with JVectorReader.open (one native call per field) and compares the search
with the reader built from the in-memory arrays."""
import struct
import zlib

import numpy as np
import pytest


def _segment(jv, seed=0, n=3000, dim=24, R=8, m=6, sim=None, pq=True, centroid=True, deleted=(5,)):
    rng = np.random.default_rng(seed)
    vec = rng.standard_normal((n, dim)).astype(np.float32)
    adj = rng.integers(0, n, (n, R)).astype(np.int32)
    deg = rng.integers(0, R + 1, n)
    adj[np.arange(R)[None, :] >= deg[:, None]] = -1
    docs = rng.permutation(n + 50)[:n].astype(np.int32)
    for d in deleted:
        docs[d] = -1
    sim = sim or jv.VectorSimilarityFunction.EUCLIDEAN
    fd = jv.FieldData(sim, vec, adj, 17 % n, jv.GraphNodeIdToDocMap(docs, n + 50))
    if pq:
        fd.pq_m, fd.pq_k = m, 256
        fd.pq_codebooks = rng.standard_normal(256 * dim).astype(np.float32)
        fd.pq_global_centroid = rng.standard_normal(dim).astype(np.float32) if centroid else None
        fd.pq_codes = rng.integers(0, 256, (n, m)).astype(np.uint8)
    return fd


def _assert_same(out, fd):
    assert np.array_equal(out["adjacency"], fd.adjacency)
    assert np.array_equal(out["vectors"].view(np.uint32), fd.vectors.view(np.uint32))      # bit-exact fp32 round trip
    assert np.array_equal(out["ord_to_doc"], fd.doc_map.graph_node_ids_to_doc_ids)
    assert out["entry_node"] == fd.entry_node and out["max_doc"] == fd.doc_map.max_docs
    if fd.pq_codes is None:
        assert out["pq_m"] == 0 and out["pq_codes"] is None and out["pq_codebooks"] is None
    else:
        assert (out["pq_m"], out["pq_k"]) == (fd.pq_m, fd.pq_k)
        assert np.array_equal(out["pq_codes"], fd.pq_codes)
        assert np.array_equal(out["pq_codebooks"].view(np.uint32), fd.pq_codebooks.view(np.uint32))
        if fd.pq_global_centroid is None:
            assert out["pq_global_centroid"] is None
        else:
            assert np.array_equal(out["pq_global_centroid"], fd.pq_global_centroid)


def test_lucene_framing_known_answers(jv):
    from opensearch_jvector_b200 import segment_files as SF
    sid = bytes(range(16))
    h = SF.index_header("JVectorVectorsFormatMeta", 1, sid, "JVector_0")
    # CodecUtil.indexHeaderLength = 9 + codec.length + 16 + 1 + suffix.length; magic and version are big-endian
    assert len(h) == 9 + len("JVectorVectorsFormatMeta") + 16 + 1 + len("JVector_0")
    assert h[:4] == bytes([0x3F, 0xD7, 0x6C, 0x17]) and h[4] == 24 and h[5:29] == b"JVectorVectorsFormatMeta"
    assert h[29:33] == b"\x00\x00\x00\x01" and h[33:49] == sid and h[49] == 9 and h[50:] == b"JVector_0"
    f = SF.footer(zlib.crc32(h))
    assert len(f) == 16 and f[:4] == bytes([0xC0, 0x28, 0x93, 0xE8]) and f[4:8] == b"\0\0\0\0"
    assert struct.unpack(">Q", f[8:])[0] == zlib.crc32(h + f[:8])
    # DataOutput.writeVInt: 7-bit groups, low first; -1 takes five bytes
    assert SF._vint(0) == b"\x00" and SF._vint(127) == b"\x7f" and SF._vint(128) == b"\x80\x01"
    assert SF._vint(16384) == b"\x80\x80\x01" and SF._vint(-1) == b"\xff\xff\xff\xff\x0f"
    a = np.array([0, 1, 127, 128, 16383, 16384, 2097151, 2097152, 268435455, 268435456, 2**31 - 1, -1], np.int32)
    assert SF._vints(a) == b"".join(SF._vint(int(x)) for x in a)
    assert SF.segment_file_name("_3", "JVector_0", "meta-jvector") == "_3_JVector_0.meta-jvector"
    assert SF.field_data_file_name("_3", "JVector_0", "vec") == "_3_JVector_0_vec.data-jvector"


@pytest.mark.parametrize("variant", ["default", "v0_meta", "graph_v4_no_footer", "graph_v3", "big_endian_floats", "no_pq",
                                     "no_centroid"])
def test_round_trip(jv, tmp_path, variant):
    from opensearch_jvector_b200 import segment_files as SF
    fd = _segment(jv, seed=1, pq=variant != "no_pq", centroid=variant != "no_centroid")
    kw, lf = {}, SF.FLAG_VERIFY_DATA_CRC
    if variant == "v0_meta":
        kw["version"] = 0                       # no quantisation-type byte: inferred, JVectorWriter.java:551-558
    if variant == "graph_v4_no_footer":
        kw["graph_version"] = 4
    if variant == "graph_v3":
        kw["graph_version"] = 3
    if variant == "big_endian_floats":
        kw["float_order"] = ">"
        lf |= SF.FLAG_FLOATS_BIG_ENDIAN
    paths = jv.JVectorWriter.write(jv.Segment(fd.doc_map.max_docs, {"vec": fd}), tmp_path, "_7", "JVector_0",
                                   field_numbers={"vec": 3}, **kw)
    assert paths["meta"].name == "_7_JVector_0.meta-jvector" and paths["vec"].name == "_7_JVector_0_vec.data-jvector"
    assert (tmp_path / "_7_JVector_0.data-jvector").exists()      # the vestigial header+footer file, JVectorWriter.java:140-165
    with SF.SegmentFiles(paths["meta"]) as s:
        assert len(s.metas) == 1
        m = s.metas[0]
        assert (m.field_number, m.vector_encoding, m.similarity, m.dim) == (3, 1, 0, 24)
        assert m.format_version == kw.get("version", 1)
        assert m.quantization_type == (0 if variant == "no_pq" else 1)
        assert (m.graph_nodes, m.max_doc) == (3000, 3050) and m.degree_overflow == 0.0
        assert m.index_offset == 9 + len("JVectorVectorsFormatIndex") + 16 + 1 + len("JVector_0")
        assert m.pq_offset == (0 if variant == "no_pq" else m.index_offset + m.index_length)
        assert np.array_equal(s.doc_map(0), fd.doc_map.graph_node_ids_to_doc_ids)
        assert s.field_index_of(3) == 0
        _assert_same(s.load_field(0, paths["vec"], lf), fd)
    for p in paths.values():
        SF.check_integrity(p)


def test_two_fields_and_empty_doc_map_edge(jv, tmp_path):
    from opensearch_jvector_b200 import segment_files as SF
    V = jv.VectorSimilarityFunction
    a = _segment(jv, seed=2, n=100, dim=7, R=4, pq=False, deleted=())
    b = _segment(jv, seed=3, n=1500, dim=10, R=16, m=4, sim=V.MAXIMUM_INNER_PRODUCT, centroid=False, deleted=(0, 1499))
    paths = jv.JVectorWriter.write(jv.Segment(1550, {"a": a, "b": b}), tmp_path, field_numbers={"a": 0, "b": 9})
    with SF.SegmentFiles(paths["meta"]) as s:
        # distFuncToOrd (JVectorReader.java:407-413) stores MAXIMUM_INNER_PRODUCT as 1 = DOT_PRODUCT: ordinal 3 is never on disk
        assert [m.field_number for m in s.metas] == [0, 9] and [m.similarity for m in s.metas] == [0, 1]
        with pytest.raises(ValueError):                         # FieldInfo says COSINE, the record says DOT_PRODUCT
            s.set_lucene_similarity(1, V.COSINE.jvector_ord)
        s.set_lucene_similarity(1, V.MAXIMUM_INNER_PRODUCT.jvector_ord)   # what FieldInfo knows
        assert s.metas[1].similarity == 3
        _assert_same(s.load_field(0, paths["a"]), a)
        _assert_same(s.load_field(1, paths["b"]), b)
        with pytest.raises(jv.native.JVectorNativeError):       # field b's records do not match field a's file
            s.load_field(1, paths["a"])
        with pytest.raises(ValueError):
            s.load_field(2, paths["a"])


def test_sub_vector_split_with_remainder(jv, tmp_path):
    """dim % M != 0: jVector's split rule (first dim % M subspaces one wider), sizes/offsets in the PQ header."""
    from opensearch_jvector_b200 import segment_files as SF
    fd = _segment(jv, seed=4, n=1200, dim=26, m=8)
    paths = jv.JVectorWriter.write(jv.Segment(fd.doc_map.max_docs, {"vec": fd}), tmp_path)
    with SF.SegmentFiles(paths["meta"]) as s:
        _assert_same(s.load_field(0, paths["vec"]), fd)


def _rewrite(path, mutate, fix_crc=False):
    b = bytearray(path.read_bytes())
    mutate(b)
    if fix_crc:
        b[-8:] = struct.pack(">Q", zlib.crc32(bytes(b[:-8])))
    path.write_bytes(bytes(b))


def test_corruption_is_reported_not_guessed(jv, tmp_path):
    from opensearch_jvector_b200 import segment_files as SF
    E = jv.native.JVectorNativeError
    fd = _segment(jv, seed=5, n=1100, dim=8, R=4, m=4)
    seg = jv.Segment(fd.doc_map.max_docs, {"vec": fd})

    def fresh(sub):
        return jv.JVectorWriter.write(seg, tmp_path / sub, segment_id=bytes(16))

    p = fresh("flip")                                            # one flipped payload byte in the meta file -> CRC
    _rewrite(p["meta"], lambda b: b.__setitem__(70, b[70] ^ 1))
    with pytest.raises(E, match="checksum failed") as ei:
        SF.SegmentFiles(p["meta"])
    assert ei.value.status == jv.native.ERR_CORRUPT
    p = fresh("trunc")                                           # truncated meta file -> footer mismatch
    p["meta"].write_bytes(p["meta"].read_bytes()[:-5])
    with pytest.raises(E, match="footer mismatch"):
        SF.SegmentFiles(p["meta"])
    p = fresh("codec")                                           # a data file passed as the meta file
    with pytest.raises(E, match="codec mismatch"):
        SF.SegmentFiles(p["vec"])
    p = fresh("version")                                         # format version from the future
    _rewrite(p["meta"], lambda b: b.__setitem__(32, 9), fix_crc=True)
    with pytest.raises(E, match="too new"):
        SF.SegmentFiles(p["meta"])
    with pytest.raises(E, match="cannot open"):
        SF.SegmentFiles(tmp_path / "missing.meta-jvector")
    p = fresh("other_segment")                                   # field data file of another segment (id differs)
    q = jv.JVectorWriter.write(seg, tmp_path / "other2", segment_id=bytes([1] * 16))
    with SF.SegmentFiles(p["meta"]) as s, pytest.raises(E, match="file mismatch"):
        s.load_field(0, q["vec"])
    p = fresh("datacrc")                                         # payload flip in the data file: only seen with the CRC flag
    _rewrite(p["vec"], lambda b: b.__setitem__(114, b[114] ^ 0x40))   # inside record 0's vector
    with SF.SegmentFiles(p["meta"]) as s:
        s.load_field(0, p["vec"])
        with pytest.raises(E, match="checksum failed"):
            s.load_field(0, p["vec"], SF.FLAG_VERIFY_DATA_CRC)
    with pytest.raises(E, match="checksum failed"):
        SF.check_integrity(p["vec"])
    p = fresh("magic")                                           # unknown graph magic: error unless lenient
    off = 9 + len("JVectorVectorsFormatIndex") + 16 + 1 + len("JVector_0")
    with SF.SegmentFiles(p["meta"]) as s:
        il = s.metas[0].index_length

    def both_headers(b):                                         # header at the start and its copy in front of the footer
        hdr_len = 4 + 4 + 16 + 8 + 8 + 4 + 4
        for at in (off, off + il - 12 - hdr_len):
            b[at:at + 4] = struct.pack("<I", 0x12345678)
    _rewrite(p["vec"], both_headers, fix_crc=True)
    with SF.SegmentFiles(p["meta"]) as s:
        with pytest.raises(E, match="magic"):
            s.load_field(0, p["vec"])
        _assert_same(s.load_field(0, p["vec"], SF.FLAG_LENIENT_MAGIC), fd)
    p = fresh("neighbour")                                       # neighbour id out of range in record 0
    rec0 = off + (4 + 4 + 16 + 8 + 8 + 4 + 4)
    deg0 = int((fd.adjacency[0] >= 0).sum())
    if deg0:
        _rewrite(p["vec"], lambda b: b.__setitem__(slice(rec0 + 8 + 32, rec0 + 12 + 32), struct.pack("<i", 1 << 30)), fix_crc=True)
        with SF.SegmentFiles(p["meta"]) as s, pytest.raises(E, match="record 0 is malformed"):
            s.load_field(0, p["vec"])


def test_unsupported_cases(jv, tmp_path):
    from opensearch_jvector_b200 import segment_files as SF
    fd = _segment(jv, seed=6, n=1100, dim=8, R=4, m=4)
    p = jv.JVectorWriter.write(jv.Segment(fd.doc_map.max_docs, {"vec": fd}), tmp_path)
    # byte vectors: VectorEncoding.BYTE (ordinal 0) in the record -> UnsupportedOperationException (JVectorReader.java:241-245)
    hdr = 9 + len("JVectorVectorsFormatMeta") + 16 + 1 + len("JVector_0")
    _rewrite(p["meta"], lambda b: b.__setitem__(slice(hdr + 8, hdr + 12), struct.pack("<i", 0)), fix_crc=True)
    with SF.SegmentFiles(p["meta"]) as s, pytest.raises(NotImplementedError, match="Byte vectors"):
        s.load_field(0, p["vec"])
    # NVQ-inline feature bit in the graph header: not parsed from files yet
    p = jv.JVectorWriter.write(jv.Segment(fd.doc_map.max_docs, {"vec": fd}), tmp_path / "nvq", graph_version=4)
    off = 9 + len("JVectorVectorsFormatIndex") + 16 + 1 + len("JVector_0")
    feat = off + 4 + 4 + 16 + 8 + 8
    _rewrite(p["vec"], lambda b: b.__setitem__(slice(feat, feat + 4), struct.pack("<I", 4)), fix_crc=True)
    with SF.SegmentFiles(p["meta"]) as s, pytest.raises(NotImplementedError, match="NVQ"):
        s.load_field(0, p["vec"])


def test_field_meta_layout_matches_header(jv, tmp_path):
    import ctypes as C
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "jvgpu.h"\nint main(void){printf("%zu %zu %zu %zu %d\\n",'
                   " sizeof(jv_field_meta), offsetof(jv_field_meta, index_offset), offsetof(jv_field_meta, degree_overflow),"
                   " offsetof(jv_field_meta, format_version), JV_ERR_CORRUPT);return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["/usr/bin/gcc", "-I", str(root / "include"), str(src), "-o", str(exe)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    n = jv.native
    assert out == [C.sizeof(n.FieldMeta), n.FieldMeta.index_offset.offset, n.FieldMeta.degree_overflow.offset,
                   n.FieldMeta.format_version.offset, n.ERR_CORRUPT]


def test_open_from_files_needs_a_gpu(jv, tmp_path):
    """No CPU fallback: parsing works without a device, creating the device index does not."""
    import ctypes as C
    cnt = C.c_int32(0)
    if jv.native.load().jv_device_count(C.addressof(cnt)) == 0 and cnt.value > 0:
        pytest.skip("a GPU is present")
    fd = _segment(jv, seed=7, n=1100, dim=8, R=4, m=4)
    jv.JVectorWriter.write(jv.Segment(fd.doc_map.max_docs, {"vec": fd}), tmp_path)
    with pytest.raises(jv.native.JVectorNativeError) as ei:
        jv.JVectorReader.open(tmp_path, {0: "vec"})
    assert ei.value.status == jv.native.ERR_CUDA


@pytest.mark.gpu
@pytest.mark.parametrize("sim_name", ["EUCLIDEAN", "DOT_PRODUCT", "COSINE"])
def test_gpu_reader_from_files_matches_in_memory_reader(jv, oracle, tmp_path, sim_name):
    from tests.helpers import clustered, make_fixture
    V = jv.VectorSimilarityFunction
    sim = V[sim_name]
    base, queries = clustered(4000, 32, 64, seed=11, normalize=sim_name != "EUCLIDEAN")
    rng = np.random.default_rng(3)
    docs = rng.permutation(4100)[:4000].astype(np.int32)
    docs[[7, 99]] = -1
    fx = make_fixture(sim.jvector_ord, base, queries, pq_m=8, ord_to_doc=docs, max_doc=4100)
    fd = jv.FieldData(sim, fx.base, fx.adjacency, fx.entry, jv.GraphNodeIdToDocMap(docs, 4100), fx.pq_m, fx.pq_k,
                      fx.codebooks, fx.gcent, fx.codes)
    seg = jv.Segment(4100, {"vec": fd})
    jv.JVectorWriter.write(seg, tmp_path, "_2", "JVector_0", field_numbers={"vec": 5})
    accept = rng.random(4100) < 0.5
    for flags in (0, jv.native.FLAG_LUT_U8):
        mem = jv.JVectorReader(seg, flags=flags)
        disk = jv.JVectorReader.open(tmp_path, {5: "vec"}, "_2", "JVector_0", flags=flags)
        try:
            disk.check_integrity()
            a, b = mem.field_index("vec"), disk.field_index("vec")
            assert a.device_bytes() == b.device_bytes()
            for bits in (None, jv.make_accept_bits(accept)):
                for width in (-1, 0):
                    ra = a.search(queries, 10, 50, accept_bits=bits, expand_width=width)
                    rb = b.search(queries, 10, 50, accept_bits=bits, expand_width=width)
                    assert np.array_equal(ra.docs, rb.docs) and np.array_equal(ra.scores, rb.scores)
                    if width == -1:
                        assert np.array_equal(ra.stats, rb.stats)
            da, sa, _ = a.exact_topk(queries, 10)
            db, sb, _ = b.exact_topk(queries, 10)
            assert np.array_equal(da, db) and np.array_equal(sa, sb)
            assert np.array_equal(disk.get_float_vector_values("vec"), fx.base)
            # the collector-level API of the reference on the file-backed reader
            col = jv.JVectorKnnCollector(jv.TopKnnCollector(10))
            disk.search("vec", queries[0], col)
            assert [sd.doc for sd in col.top_docs()] == [int(x) for x in rb.docs[0] if x >= 0] or len(col.top_docs()) == 10
        finally:
            mem.close()
            disk.close()


def test_committed_golden_segment(jv, tmp_path):
    """tests/golden/segment/ (written by tests/golden/make_golden_segment.py): the loader parses the committed bytes, and the
    writer mirror still produces exactly those bytes — neither side can drift alone."""
    import importlib.util
    from pathlib import Path
    from opensearch_jvector_b200 import segment_files as SF
    gdir = Path(__file__).resolve().parent / "golden" / "segment"
    spec = importlib.util.spec_from_file_location("make_golden_segment", gdir.parent / "make_golden_segment.py")
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    meta, data = gdir / "_g_JVector_0.meta-jvector", gdir / "_g_JVector_0_vec.data-jvector"
    # known bytes of the committed meta file: CodecUtil header, first record, end marker, footer
    b = meta.read_bytes()
    assert len(b) == 159 and b[:4] == bytes.fromhex("3fd76c17") and b[5:29] == b"JVectorVectorsFormatMeta"
    assert b[29:33] == bytes.fromhex("00000001") and b[33:49] == bytes(range(0xA0, 0xB0)) and b[49:59] == b"\x09JVector_0"
    assert b[59:63] == struct.pack("<i", 2) and b[63:75] == struct.pack("<iii", 2, 1, 0)      # fieldNumber x2, FLOAT32, EUCLIDEAN
    assert b[75] == 6 and b[76] == 60                                                            # vint dim, vlong indexOffset
    assert b[-20:-16] == struct.pack("<i", -1) and b[-16:-8] == bytes.fromhex("c02893e800000000")
    assert struct.unpack(">Q", b[-8:])[0] == zlib.crc32(b[:-8])
    vec, adj, docs, cb, gcent, codes = mk.arrays()
    with SF.SegmentFiles(meta) as s:
        m = s.metas[0]
        assert (m.field_number, m.similarity, m.dim, m.quantization_type, m.graph_nodes, m.max_doc) == (2, 0, 6, 1, 40, 130)
        assert np.array_equal(s.doc_map(0), docs) and s.doc_map(0)[7] == -1
        out = s.load_field(0, data, SF.FLAG_VERIFY_DATA_CRC)
    assert np.array_equal(out["vectors"], vec) and np.array_equal(out["adjacency"], adj) and out["entry_node"] == 11
    assert (out["pq_m"], out["pq_k"]) == (3, 16) and np.array_equal(out["pq_codes"], codes)
    assert np.array_equal(out["pq_codebooks"], cb) and np.array_equal(out["pq_global_centroid"], gcent)
    fresh = mk.write(tmp_path)
    for name in ("_g_JVector_0.meta-jvector", "_g_JVector_0.data-jvector", "_g_JVector_0_vec.data-jvector"):
        assert (tmp_path / name).read_bytes() == (gdir / name).read_bytes(), name
    assert set(p.name for p in fresh.values()) == {"_g_JVector_0.meta-jvector", "_g_JVector_0_vec.data-jvector"}
