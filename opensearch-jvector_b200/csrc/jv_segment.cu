// jv_segment.cu — SURVEY 8f-1: loader for the files JVectorWriter persists (host code only; no kernels).
//
// What is read, and where the reference writes / reads it (paths relative to .../codec/jvector/):
//   meta file        JVectorWriter.java:134-157 (header), :299-300 + :528-540 (records), :573-577 (end marker, footer);
//                    read at JVectorReader.java:52-81,255-262; doc map GraphNodeIdToDocMap.java:39-59,169-176
//   field data file  JVectorWriter.java:383-433 (header, graph), :469-510 (PQ blob, footer); read at
//                    JVectorReader.java:306-331 (slice [0, header + indexLength), graph header at indexOffset)
//   Lucene framing   CodecUtil.writeIndexHeader / writeFooter: big-endian magic 0x3fd76c17, codec name, version,
//                    16-byte segment id, suffix; footer = ~magic, algorithm 0, CRC-32 of everything before the CRC
//   ints / longs     little-endian (Lucene >= 9 IndexOutput; JVectorIndexWriter.java:72-83)
// The layout inside the OnDiskGraphIndex and PQVectors blobs is jVector's (un-vendored jar, 4.0.0-rc.9); it is restated
// here from the published format: magic + version + common header (+ layer table from version 4, + the
// header repeated as a footer from version 5), a feature bit set, dense layer-0 records
// [int ordinal][fp32 vector][int degree][int neighbours[R], -1 padded], then ProductQuantization.write + PQVectors.write.
// The open questions of SURVEY B.3 are flags (JV_SEGMENT_FLAG_*), never silent guesses: a mismatch fails with
// JV_ERR_CORRUPT / JV_ERR_UNSUPPORTED and says what was expected.
#include <errno.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <memory>
#include <thread>

#include "jv_internal.h"

namespace {

using jv::set_error;

constexpr uint32_t CODEC_MAGIC = 0x3fd76c17u;         // CodecUtil.CODEC_MAGIC
constexpr uint32_t FOOTER_MAGIC = ~CODEC_MAGIC;       // CodecUtil.FOOTER_MAGIC
constexpr int FOOTER_LENGTH = 16;                     // CodecUtil.footerLength()
constexpr const char *META_CODEC_NAME = "JVectorVectorsFormatMeta";   // JVectorFormat.java:23
constexpr const char *INDEX_CODEC_NAME = "JVectorVectorsFormatIndex"; // JVectorFormat.java:24
constexpr int VERSION_START = 0, VERSION_WITH_QUANTIZATION_TYPE = 1, VERSION_CURRENT = 1; // JVectorFormat.java:31-33
constexpr int DOC_MAP_VERSION = 1;                    // GraphNodeIdToDocMap.VERSION
// jVector blobs [restated, SURVEY B.2]
constexpr uint32_t GRAPH_MAGIC = 0xFFFF0D61u;         // OnDiskGraphIndex.MAGIC (versions >= 3)
constexpr uint32_t GRAPH_FOOTER_MAGIC = 0x4a564244u;  // header-as-footer marker (version >= 5)
constexpr int GRAPH_FOOTER_SIZE = 12;                 // long headerOffset + int magic
constexpr uint32_t PQ_MAGIC = 0x75EC4012u;            // ProductQuantization.MAGIC (versions >= 3)
constexpr uint32_t FEATURE_INLINE_VECTORS = 1u << 0, FEATURE_FUSED_ADC = 1u << 1, FEATURE_NVQ_VECTORS = 1u << 2;

#define JV_CORRUPT(...)              \
    do {                             \
        set_error(__VA_ARGS__);      \
        return JV_ERR_CORRUPT;       \
    } while (0)

// ---- CRC-32 (java.util.zip.CRC32 = zlib polynomial), slice-by-8 -------------------------------
uint32_t g_crc[8][256];
std::once_flag g_crc_once;
void crc_init() {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        g_crc[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
        for (int t = 1; t < 8; t++) g_crc[t][i] = (g_crc[t - 1][i] >> 8) ^ g_crc[0][g_crc[t - 1][i] & 0xFF];
}
uint32_t crc32_of(const uint8_t *p, size_t n) {
    std::call_once(g_crc_once, crc_init);
    uint32_t c = 0xFFFFFFFFu;
    while (n && ((uintptr_t)p & 7)) c = g_crc[0][(c ^ *p++) & 0xFF] ^ (c >> 8), n--;
    while (n >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        uint32_t lo = (uint32_t)w ^ c, hi = (uint32_t)(w >> 32);
        c = g_crc[7][lo & 0xFF] ^ g_crc[6][(lo >> 8) & 0xFF] ^ g_crc[5][(lo >> 16) & 0xFF] ^ g_crc[4][lo >> 24] ^
            g_crc[3][hi & 0xFF] ^ g_crc[2][(hi >> 8) & 0xFF] ^ g_crc[1][(hi >> 16) & 0xFF] ^ g_crc[0][hi >> 24];
        p += 8, n -= 8;
    }
    while (n--) c = g_crc[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
    return ~c;
}

// ---- read-only mapping of one file ---------------------------------------------------------------
struct Mapped {
    const uint8_t *p = nullptr;
    size_t len = 0;
    Mapped() = default;
    Mapped(const Mapped &) = delete;
    Mapped &operator=(const Mapped &) = delete;
    ~Mapped() {
        if (p && len) munmap((void *)p, len);
    }
    int32_t open(const char *path) {
        int fd = ::open(path, O_RDONLY);
        if (fd < 0) JV_CORRUPT("cannot open %s: %s", path, strerror(errno));
        struct stat st;
        if (fstat(fd, &st) != 0 || st.st_size <= 0) {
            ::close(fd);
            JV_CORRUPT("cannot stat %s (or it is empty)", path);
        }
        len = (size_t)st.st_size;
        void *m = mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0);
        ::close(fd);
        if (m == MAP_FAILED) {
            len = 0;
            JV_CORRUPT("mmap of %s failed: %s", path, strerror(errno));
        }
        p = (const uint8_t *)m;
        return JV_OK;
    }
};

// ---- bounds-checked cursor (Lucene DataInput subset) -----------------------------------------
struct Cur {
    const uint8_t *b;
    size_t end, pos;
    bool ok = true;
    Cur(const uint8_t *base, size_t begin, size_t end_) : b(base), end(end_), pos(begin) {}
    bool need(size_t n) {
        if (!ok || n > end - pos) ok = false;
        return ok;
    }
    uint8_t u8() { return need(1) ? b[pos++] : 0; }
    uint32_t be32() {
        if (!need(4)) return 0;
        uint32_t v = (uint32_t)b[pos] << 24 | (uint32_t)b[pos + 1] << 16 | (uint32_t)b[pos + 2] << 8 | b[pos + 3];
        pos += 4;
        return v;
    }
    uint64_t be64() {
        uint64_t hi = be32();
        return hi << 32 | be32();
    }
    int32_t le32() {
        if (!need(4)) return 0;
        uint32_t v;
        memcpy(&v, b + pos, 4);  // the hosts this library runs on are little-endian
        pos += 4;
        return (int32_t)v;
    }
    int64_t le64() {
        if (!need(8)) return 0;
        uint64_t v;
        memcpy(&v, b + pos, 8);
        pos += 8;
        return (int64_t)v;
    }
    int32_t vint() {  // DataInput.readVInt: 7 bits per byte, low group first; negative values take 5 bytes
        uint32_t v = 0;
        for (int shift = 0; shift < 35; shift += 7) {
            uint8_t x = u8();
            v |= (uint32_t)(x & 0x7F) << shift;
            if (!(x & 0x80)) return (int32_t)v;
        }
        ok = false;
        return 0;
    }
    int64_t vlong() {
        uint64_t v = 0;
        for (int shift = 0; shift < 63; shift += 7) {
            uint8_t x = u8();
            v |= (uint64_t)(x & 0x7F) << shift;
            if (!(x & 0x80)) return (int64_t)v;
        }
        ok = false;
        return 0;
    }
    const uint8_t *bytes(size_t n) {
        if (!need(n)) return nullptr;
        const uint8_t *r = b + pos;
        pos += n;
        return r;
    }
};

struct LuceneHeader {
    int version = 0;
    uint8_t id[16] = {0};
    std::string suffix;
    size_t length = 0;
};

// CodecUtil.checkIndexHeader without the expected id / suffix (the caller compares the two files' instead)
int32_t read_index_header(Cur &c, const char *codec, const char *what, LuceneHeader *h) {
    size_t start = c.pos;
    uint32_t magic = c.be32();
    if (!c.ok || magic != CODEC_MAGIC) JV_CORRUPT("%s: codec header mismatch: actual header=0x%08x vs expected header=0x%08x", what, magic, CODEC_MAGIC);
    int32_t len = c.vint();
    const uint8_t *name = (len >= 0 && len < 128) ? c.bytes((size_t)len) : nullptr;
    if (!name || (size_t)len != strlen(codec) || memcmp(name, codec, (size_t)len) != 0)
        JV_CORRUPT("%s: codec mismatch: expected codec=%s", what, codec);
    h->version = (int32_t)c.be32();
    if (!c.ok || h->version < VERSION_START) JV_CORRUPT("%s: format version %d is too old (minimum %d)", what, h->version, VERSION_START);
    if (h->version > VERSION_CURRENT) JV_CORRUPT("%s: format version %d is too new (maximum %d)", what, h->version, VERSION_CURRENT);
    const uint8_t *id = c.bytes(16);
    if (!id) JV_CORRUPT("%s: truncated index header", what);
    memcpy(h->id, id, 16);
    int sl = c.u8();
    const uint8_t *sfx = c.bytes((size_t)sl);
    if (!c.ok) JV_CORRUPT("%s: truncated index header", what);
    h->suffix.assign((const char *)sfx, (size_t)sl);
    h->length = c.pos - start;
    return JV_OK;
}

// CodecUtil.checkFooter (structure) + optional full CRC; the footer occupies the last 16 bytes of the file
int32_t check_footer(const Mapped &f, const char *what, bool verify_crc) {
    if (f.len < (size_t)FOOTER_LENGTH) JV_CORRUPT("%s: misplaced codec footer (file truncated?)", what);
    Cur c(f.p, f.len - FOOTER_LENGTH, f.len);
    uint32_t magic = c.be32();
    if (magic != FOOTER_MAGIC) JV_CORRUPT("%s: codec footer mismatch (file truncated?): actual footer=0x%08x vs expected footer=0x%08x", what, magic, FOOTER_MAGIC);
    uint32_t algo = c.be32();
    if (algo != 0) JV_CORRUPT("%s: codec footer mismatch: unknown algorithmID: %u", what, algo);
    uint64_t stored = c.be64();
    if (stored >> 32) JV_CORRUPT("%s: illegal CRC-32 checksum: %llu", what, (unsigned long long)stored);
    if (verify_crc) {
        uint32_t actual = crc32_of(f.p, f.len - 8);
        if (actual != (uint32_t)stored) JV_CORRUPT("%s: checksum failed (hardware problem?) : expected=%x actual=%x", what, (uint32_t)stored, actual);
    }
    return JV_OK;
}

struct FieldRecord {
    jv_field_meta meta;
    std::vector<int32_t> ord_to_doc;
};

inline float load_float(const uint8_t *p, bool big_endian) {
    uint32_t v;
    memcpy(&v, p, 4);
    if (big_endian) v = __builtin_bswap32(v);
    float f;
    memcpy(&f, &v, 4);
    return f;
}

void copy_floats(float *dst, const uint8_t *src, size_t count, bool big_endian) {
    if (!big_endian) {
        memcpy(dst, src, count * 4);
        return;
    }
    for (size_t i = 0; i < count; i++) dst[i] = load_float(src + 4 * i, true);
}

}  // namespace

struct jv_segment {
    std::string meta_path;
    uint32_t flags = 0;
    LuceneHeader header;
    std::vector<FieldRecord> fields;
};

struct jv_field_data {
    jv_field_meta meta;
    int graph_version = 0, num_layers = 1;
    int64_t n = 0;
    int dim = 0, R = 0, entry = 0;
    std::vector<int32_t> adjacency, ord_to_doc;
    float *vectors = nullptr;  // n * dim, malloc'd (can be several GB)
    int pq_m = 0, pq_k = 0;
    std::vector<float> codebooks, gcent;
    std::vector<uint8_t> codes;
    ~jv_field_data() { free(vectors); }
};

namespace {

// OnDiskGraphIndex bytes -> adjacency / vectors  [layout restated, SURVEY B.2]
int32_t decode_graph(const Mapped &f, const jv_field_meta &m, uint32_t flags, jv_field_data *out) {
    const bool be = flags & JV_SEGMENT_FLAG_FLOATS_BIG_ENDIAN, lenient = flags & JV_SEGMENT_FLAG_LENIENT_MAGIC;
    const size_t begin = (size_t)m.index_offset, end = begin + (size_t)m.index_length;
    if (m.index_offset < 0 || m.index_length < 16 || end > f.len - FOOTER_LENGTH || end < begin)
        JV_CORRUPT("graph index range [%lld, +%lld) lies outside the field data file (%zu bytes)", (long long)m.index_offset,
                   (long long)m.index_length, f.len);
    // version >= 5 repeats the header as a footer; JVectorReader.java:306-316 slices the file so that it ends there
    size_t header_at = begin;
    {
        Cur t(f.p, end - GRAPH_FOOTER_SIZE, end);
        int64_t off = t.le64();
        uint32_t magic = (uint32_t)t.le32();
        if (magic == GRAPH_FOOTER_MAGIC) {
            if (off < (int64_t)begin || off >= (int64_t)end) JV_CORRUPT("graph footer points outside the graph (%lld)", (long long)off);
            header_at = (size_t)off;
        }
    }
    Cur c(f.p, header_at, end);
    uint32_t magic = (uint32_t)c.le32();
    if (magic != GRAPH_MAGIC && !lenient)
        JV_CORRUPT("OnDiskGraphIndex magic 0x%08x, expected 0x%08x (a version-2 graph without magic is not supported; "
                   "JV_SEGMENT_FLAG_LENIENT_MAGIC skips this check)", magic, GRAPH_MAGIC);
    int version = c.le32();
    if (version < 3 || version > 6) {
        set_error("OnDiskGraphIndex version %d is not supported (3..6)", version);
        return JV_ERR_UNSUPPORTED;
    }
    int size0 = c.le32(), dim = c.le32(), entry = c.le32(), degree0 = c.le32();
    int id_upper = size0, layers = 1;
    if (version >= 4) {
        id_upper = c.le32();
        layers = c.le32();
        if (!c.ok || layers < 1 || layers > 64) JV_CORRUPT("bad layer count %d in the graph header", layers);
        for (int l = 0; l < layers; l++) {
            int ls = c.le32(), ld = c.le32();
            if (l == 0 && (ls != size0 || ld != degree0)) JV_CORRUPT("layer table disagrees with the common header (%d/%d vs %d/%d)", ls, ld, size0, degree0);
        }
    }
    uint32_t features = (uint32_t)c.le32();
    if (features & FEATURE_NVQ_VECTORS) {
        set_error("NVQ-inline graph records are not parsed from files yet: pass the decoded NVQ arrays through jv_index_desc.nvq_*");
        return JV_ERR_UNSUPPORTED;
    }
    if (features != FEATURE_INLINE_VECTORS) {
        set_error("graph feature set 0x%x is not supported (the plugin writes INLINE_VECTORS only, JVectorWriter.java:482-485)", features);
        return JV_ERR_UNSUPPORTED;
    }
    int fdim = c.le32();  // InlineVectors.writeHeader
    if (!c.ok) JV_CORRUPT("truncated graph header");
    if (dim != m.dim || fdim != dim) JV_CORRUPT("graph dimension %d / inline-vector dimension %d differ from the field's %d", dim, fdim, m.dim);
    if (degree0 < 1 || degree0 > 128) JV_CORRUPT("graph max degree %d outside [1,128]", degree0);
    if (id_upper < size0 || id_upper < 0) JV_CORRUPT("idUpperBound %d < size %d", id_upper, size0);
    const size_t header_size = c.pos - header_at;
    const int64_t n = id_upper;
    const size_t rec = 4 + (size_t)dim * 4 + 4 + (size_t)degree0 * 4;
    const size_t rec_begin = begin + header_size;
    if (rec_begin + (size_t)n * rec > end) JV_CORRUPT("%lld layer-0 records of %zu bytes do not fit the graph blob", (long long)n, rec);
    if (n > 0 && (entry < 0 || entry >= n)) JV_CORRUPT("entry node %d outside [0,%lld)", entry, (long long)n);
    if (n != m.graph_nodes) JV_CORRUPT("graph has %lld nodes, the doc map %d (ordinals must be dense, SURVEY B.3-3)", (long long)n, m.graph_nodes);

    out->graph_version = version, out->num_layers = layers;
    out->n = n, out->dim = dim, out->R = degree0, out->entry = entry;
    out->adjacency.assign((size_t)n * degree0, -1);
    out->vectors = (float *)malloc(std::max<size_t>((size_t)n * dim * 4, 4));
    if (!out->vectors) {
        set_error("out of host memory for %lld x %d vectors", (long long)n, dim);
        return JV_ERR_OUT_OF_MEMORY;
    }
    // split the DiskANN-style records into structure-of-arrays, a few threads over disjoint ordinal ranges
    int nt = (int)std::min<int64_t>(8, std::max<int64_t>(1, n / 16384));
    std::vector<int64_t> bad(nt, -1);
    auto work = [&](int t) {
        int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
        for (int64_t i = lo; i < hi; i++) {
            const uint8_t *r = f.p + rec_begin + (size_t)i * rec;
            int32_t ord, deg;
            memcpy(&ord, r, 4);
            copy_floats(out->vectors + (size_t)i * dim, r + 4, (size_t)dim, be);
            memcpy(&deg, r + 4 + (size_t)dim * 4, 4);
            if (ord != (int32_t)i || deg < 0 || deg > degree0) {
                if (bad[t] < 0) bad[t] = i;
                continue;
            }
            const uint8_t *nb = r + 8 + (size_t)dim * 4;
            int32_t *row = out->adjacency.data() + (size_t)i * degree0;
            for (int j = 0; j < deg; j++) {
                int32_t v;
                memcpy(&v, nb + 4 * j, 4);
                if (v < 0 || v >= n) {
                    if (bad[t] < 0) bad[t] = i;
                    break;
                }
                row[j] = v;
            }
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    for (int t = 0; t < nt; t++)
        if (bad[t] >= 0) JV_CORRUPT("layer-0 record %lld is malformed (ordinal, degree or a neighbour id out of range)", (long long)bad[t]);
    return JV_OK;
}

// ProductQuantization.write + PQVectors.write  [layout restated, SURVEY B.2]
int32_t decode_pq(const Mapped &f, const jv_field_meta &m, uint32_t flags, jv_field_data *out) {
    const bool be = flags & JV_SEGMENT_FLAG_FLOATS_BIG_ENDIAN, lenient = flags & JV_SEGMENT_FLAG_LENIENT_MAGIC;
    const size_t begin = (size_t)m.pq_offset, end = begin + (size_t)m.pq_length;
    if (m.pq_offset < 0 || end > f.len - FOOTER_LENGTH || end < begin)
        JV_CORRUPT("PQ blob range [%lld, +%lld) lies outside the field data file", (long long)m.pq_offset, (long long)m.pq_length);
    Cur c(f.p, begin, end);
    uint32_t magic = (uint32_t)c.le32();
    if (magic != PQ_MAGIC && !lenient) JV_CORRUPT("ProductQuantization magic 0x%08x, expected 0x%08x", magic, PQ_MAGIC);
    int version = c.le32();
    if (version < 3 || version > 6) {
        set_error("ProductQuantization version %d is not supported (3..6)", version);
        return JV_ERR_UNSUPPORTED;
    }
    int gl = c.le32();
    if (!c.ok || (gl != 0 && gl != m.dim)) JV_CORRUPT("global centroid length %d (expected 0 or %d)", gl, m.dim);
    if (gl) {
        const uint8_t *p = c.bytes((size_t)gl * 4);
        if (!p) JV_CORRUPT("truncated global centroid");
        out->gcent.resize((size_t)gl);
        copy_floats(out->gcent.data(), p, (size_t)gl, be);
    }
    int32_t aniso_bits = c.le32();  // DataOutput.writeFloat -> JVectorIndexWriter.writeFloat -> little-endian int
    float aniso;
    memcpy(&aniso, &aniso_bits, 4);
    if (c.ok && aniso > 0.f) {  // UNWEIGHTED = -1; the plugin never trains anisotropic codebooks (JVectorIndexQuantization.java:123-131)
        set_error("anisotropic PQ (threshold %g) is not supported", aniso);
        return JV_ERR_UNSUPPORTED;
    }
    int M = c.le32();
    if (!c.ok || M < 1 || M > m.dim) JV_CORRUPT("bad subspace count %d", M);
    jv::PqShape shape;
    int K_placeholder = 1;
    shape.init(m.dim, M, K_placeholder);
    for (int s = 0; s < M; s++) {
        int sz = c.le32(), off = c.le32();
        if (!c.ok || sz != shape.size[s] || off != shape.off[s]) {
            set_error("subspace %d has size/offset %d/%d; only jVector's default split (%d/%d) is supported", s, sz, off, shape.size[s], shape.off[s]);
            return c.ok ? JV_ERR_UNSUPPORTED : JV_ERR_CORRUPT;
        }
    }
    int K = c.le32();
    if (!c.ok || K < 1 || K > 256) JV_CORRUPT("bad cluster count %d", K);
    size_t cb_floats = (size_t)K * m.dim;
    const uint8_t *cb = c.bytes(cb_floats * 4);
    if (!cb) JV_CORRUPT("truncated codebooks");
    out->codebooks.resize(cb_floats);
    copy_floats(out->codebooks.data(), cb, cb_floats, be);
    int N = c.le32(), M2 = c.le32();
    if (!c.ok || M2 != M) JV_CORRUPT("PQVectors subspace count %d differs from the codebook's %d", M2, M);
    if (N != out->n) JV_CORRUPT("PQVectors holds %d vectors, the graph %lld", N, (long long)out->n);
    const uint8_t *codes = c.bytes((size_t)N * M);
    if (!codes) JV_CORRUPT("truncated PQ codes");
    if (c.pos != end) JV_CORRUPT("%zu trailing bytes after the PQ codes", end - c.pos);
    out->codes.assign(codes, codes + (size_t)N * M);
    out->pq_m = M, out->pq_k = K;
    return JV_OK;
}

}  // namespace

extern "C" {

int32_t jv_segment_open(const char *meta_path, uint32_t flags, jv_segment **out) {
    JV_REQUIRE(meta_path != nullptr && out != nullptr, "meta_path/out is NULL");
    *out = nullptr;
    Mapped f;
    JV_TRY(f.open(meta_path));
    std::unique_ptr<jv_segment> seg(new jv_segment());
    seg->meta_path = meta_path, seg->flags = flags;
    Cur c(f.p, 0, f.len);
    JV_TRY(read_index_header(c, META_CODEC_NAME, meta_path, &seg->header));
    JV_TRY(check_footer(f, meta_path, true));  // the reference reads the meta file through a ChecksumIndexInput
    c.end = f.len - FOOTER_LENGTH;
    const int version = seg->header.version;
    for (;;) {  // JVectorReader.readFields, JVectorReader.java:255-262
        int32_t number = c.le32();
        if (!c.ok) JV_CORRUPT("%s: truncated before the end-of-fields marker", meta_path);
        if (number == -1) break;
        FieldRecord r;
        memset(&r.meta, 0, sizeof(r.meta));
        jv_field_meta &m = r.meta;
        m.struct_size = (int32_t)sizeof(jv_field_meta);
        m.format_version = version;
        m.field_number = c.le32();  // VectorIndexFieldMetadata(IndexInput, version), JVectorWriter.java:542-562
        m.vector_encoding = c.le32();
        m.similarity = c.le32();
        m.dim = c.vint();
        m.index_offset = c.vlong(), m.index_length = c.vlong();
        m.pq_offset = c.vlong(), m.pq_length = c.vlong();
        if (version >= VERSION_WITH_QUANTIZATION_TYPE) m.quantization_type = c.u8();
        else m.quantization_type = m.pq_length > 0 ? 1 : 0;  // v0: PQ iff a compressed blob is present
        int32_t bits = c.le32();
        memcpy(&m.degree_overflow, &bits, 4);
        int32_t map_version = c.le32();  // GraphNodeIdToDocMap(IndexInput), GraphNodeIdToDocMap.java:39-59
        if (!c.ok) JV_CORRUPT("%s: truncated field record", meta_path);
        if (map_version != DOC_MAP_VERSION) JV_CORRUPT("Unsupported version: %d", map_version);
        if (m.field_number != number) JV_CORRUPT("%s: field number %d repeated as %d", meta_path, number, m.field_number);
        if (m.vector_encoding < 0 || m.vector_encoding > 1) JV_CORRUPT("Invalid vector encoding id: %d", m.vector_encoding);
        // simOrd indexes JVECTOR_SUPPORTED_SIMILARITY_FUNCTIONS = [EUCLIDEAN, DOT_PRODUCT, COSINE, DOT_PRODUCT] (JVectorReader.java:389-394):
        // distFuncToOrd writes indexOf(...) so MAXIMUM_INNER_PRODUCT fields are stored as 1 and ordinal 3 never appears on disk; a
        // 3 would still resolve to DOT_PRODUCT (ordToDistFunc).  Whether the field is MAXIMUM_INNER_PRODUCT is FieldInfo's knowledge:
        // jv_segment_set_lucene_similarity().
        if (m.similarity < 0 || m.similarity > 3) JV_CORRUPT("invalid distance function: %d", m.similarity);
        if (m.similarity == 3) m.similarity = JV_SIM_DOT;
        if (m.quantization_type < 0 || m.quantization_type > 2) JV_CORRUPT("unknown quantization type %d", m.quantization_type);
        m.graph_nodes = c.vint();
        m.max_doc = c.vint();
        if (!c.ok || m.graph_nodes < 0 || m.max_doc < 0 || m.dim < 1) JV_CORRUPT("%s: bad doc map header / dimension", meta_path);
        if ((size_t)m.graph_nodes > c.end - c.pos) JV_CORRUPT("%s: doc map of %d entries does not fit the file", meta_path, m.graph_nodes);
        r.ord_to_doc.assign((size_t)m.graph_nodes, -1);
        for (int32_t o = 0; o < m.graph_nodes; o++) {
            int32_t doc = c.vint();
            if (doc != -1) {  // deleted documents stay -1
                if (doc < 0 || doc >= m.max_doc) JV_CORRUPT("%s: docId %d of ordinal %d outside [0,%d)", meta_path, doc, o, m.max_doc);
                r.ord_to_doc[(size_t)o] = doc;
            }
        }
        if (!c.ok) JV_CORRUPT("%s: truncated doc map", meta_path);
        seg->fields.push_back(std::move(r));
    }
    if (c.pos != c.end) JV_CORRUPT("%s: %zu unread bytes before the footer", meta_path, c.end - c.pos);
    *out = seg.release();
    return JV_OK;
}

int32_t jv_segment_close(jv_segment *segment) {
    delete segment;
    return JV_OK;
}

int32_t jv_segment_field_count(const jv_segment *segment, int32_t *out_count) {
    JV_REQUIRE(segment != nullptr && out_count != nullptr, "segment/out_count is NULL");
    *out_count = (int32_t)segment->fields.size();
    return JV_OK;
}

int32_t jv_segment_field_meta(const jv_segment *segment, int32_t i, jv_field_meta *out_meta) {
    JV_REQUIRE(segment != nullptr && out_meta != nullptr, "segment/out_meta is NULL");
    JV_REQUIRE(i >= 0 && i < (int32_t)segment->fields.size(), "field index %d out of range", i);
    JV_REQUIRE(out_meta->struct_size == (int32_t)sizeof(jv_field_meta), "jv_field_meta.struct_size mismatch");
    *out_meta = segment->fields[(size_t)i].meta;
    return JV_OK;
}

int32_t jv_segment_set_lucene_similarity(jv_segment *segment, int32_t i, int32_t lucene_similarity) {
    JV_REQUIRE(segment != nullptr, "segment is NULL");
    JV_REQUIRE(i >= 0 && i < (int32_t)segment->fields.size(), "field index %d out of range", i);
    JV_REQUIRE(lucene_similarity >= JV_SIM_EUCLIDEAN && lucene_similarity <= JV_SIM_MIP, "invalid distance function: %d", lucene_similarity);
    jv_field_meta &m = segment->fields[(size_t)i].meta;
    const int32_t stored = m.similarity == JV_SIM_MIP ? JV_SIM_DOT : m.similarity; // what the meta record said
    const int32_t expect = lucene_similarity == JV_SIM_MIP ? JV_SIM_DOT : lucene_similarity;
    JV_REQUIRE(stored == expect, "field %d: FieldInfo similarity %d does not match the meta record's distance function %d", m.field_number,
               lucene_similarity, stored);
    m.similarity = lucene_similarity;
    return JV_OK;
}

int32_t jv_segment_field_doc_map(const jv_segment *segment, int32_t i, int32_t *out_ord_to_doc, int32_t capacity) {
    JV_REQUIRE(segment != nullptr, "segment is NULL");
    JV_REQUIRE(i >= 0 && i < (int32_t)segment->fields.size(), "field index %d out of range", i);
    const FieldRecord &r = segment->fields[(size_t)i];
    JV_REQUIRE(capacity >= r.meta.graph_nodes && (out_ord_to_doc != nullptr || r.meta.graph_nodes == 0), "doc map needs %d entries", r.meta.graph_nodes);
    if (r.meta.graph_nodes) memcpy(out_ord_to_doc, r.ord_to_doc.data(), (size_t)r.meta.graph_nodes * 4);
    return JV_OK;
}

int32_t jv_segment_load_field(const jv_segment *segment, int32_t i, const char *field_data_path, uint32_t flags, jv_field_data **out) {
    JV_REQUIRE(segment != nullptr && field_data_path != nullptr && out != nullptr, "segment/field_data_path/out is NULL");
    *out = nullptr;
    JV_REQUIRE(i >= 0 && i < (int32_t)segment->fields.size(), "field index %d out of range", i);
    const FieldRecord &r = segment->fields[(size_t)i];
    flags |= segment->flags;
    if (r.meta.vector_encoding != 1) {  // JVectorReader.java:241-245, JVectorWriter.java:176-184
        set_error("Byte vectors are not supported by jVector");
        return JV_ERR_UNSUPPORTED;
    }
    Mapped f;
    JV_TRY(f.open(field_data_path));
    Cur c(f.p, 0, f.len);
    LuceneHeader h;
    JV_TRY(read_index_header(c, INDEX_CODEC_NAME, field_data_path, &h));
    if (memcmp(h.id, segment->header.id, 16) != 0 || h.suffix != segment->header.suffix)
        JV_CORRUPT("%s: file mismatch, segment id / suffix differ from the meta file's", field_data_path);
    JV_TRY(check_footer(f, field_data_path, flags & JV_SEGMENT_FLAG_VERIFY_DATA_CRC));
    if ((size_t)r.meta.index_offset != h.length)
        JV_CORRUPT("%s: graph offset %lld does not follow the %zu-byte index header", field_data_path, (long long)r.meta.index_offset, h.length);
    std::unique_ptr<jv_field_data> d(new jv_field_data());
    d->meta = r.meta;
    d->ord_to_doc = r.ord_to_doc;
    JV_TRY(decode_graph(f, r.meta, flags, d.get()));
    if (r.meta.pq_length > 0) {
        if ((size_t)r.meta.pq_offset != (size_t)r.meta.index_offset + (size_t)r.meta.index_length)
            JV_CORRUPT("%s: the PQ blob does not follow the graph (JVectorWriter.java:496-498)", field_data_path);
        JV_TRY(decode_pq(f, r.meta, flags, d.get()));
    } else if (r.meta.quantization_type == 1) {
        JV_CORRUPT("%s: quantization type PQ without a compressed-vectors blob", field_data_path);
    }
    *out = d.release();
    return JV_OK;
}

int32_t jv_field_data_desc(const jv_field_data *d, jv_index_desc *out) {
    JV_REQUIRE(d != nullptr && out != nullptr, "data/out_desc is NULL");
    memset(out, 0, sizeof(*out));
    out->struct_size = (int32_t)sizeof(jv_index_desc);
    out->similarity = d->meta.similarity;
    out->dim = d->dim;
    out->max_degree = d->R;
    out->n = d->n;
    out->entry_node = d->entry;
    out->max_doc = d->meta.max_doc;
    out->adjacency = d->adjacency.data();
    out->vectors = d->vectors;
    out->ord_to_doc = d->ord_to_doc.data();
    if (d->pq_m > 0) {
        out->pq_m = d->pq_m, out->pq_k = d->pq_k;
        out->pq_codebooks = d->codebooks.data();
        out->pq_global_centroid = d->gcent.empty() ? nullptr : d->gcent.data();
        out->pq_codes = d->codes.data();
    }
    return JV_OK;
}

int32_t jv_field_data_free(jv_field_data *data) {
    delete data;
    return JV_OK;
}

int32_t jv_segment_index_create(const jv_segment *segment, int32_t i, const char *field_data_path, int32_t device,
                                uint32_t index_flags, uint32_t load_flags, jv_index **out_index) {
    JV_REQUIRE(out_index != nullptr, "out_index is NULL");
    *out_index = nullptr;
    jv_field_data *d = nullptr;
    JV_TRY(jv_segment_load_field(segment, i, field_data_path, load_flags, &d));
    jv_index_desc desc;
    int32_t st = jv_field_data_desc(d, &desc);
    if (st == JV_OK) {
        desc.device = device, desc.flags = index_flags;
        st = jv_index_create(&desc, out_index);
    }
    jv_field_data_free(d);
    return st;
}

int32_t jv_file_check_integrity(const char *path) {
    JV_REQUIRE(path != nullptr, "path is NULL");
    Mapped f;
    JV_TRY(f.open(path));
    return check_footer(f, path, true);
}

}  // extern "C"
