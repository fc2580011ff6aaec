// npz.cu -- native reader for the `.npz` index shards (host code only; part of the same C-ABI library).
// Replaces, for the three big members, what upstream does with scipy on one Python thread
// (src/ir/retriever/index.py:172-176: load_npz -> vstack -> astype): every member of every shard is inflated by
// its own caller thread straight into its slice of the final arrays, with the int64 -> int32 / fp32 -> fp16
// conversion and the row-pointer offset of the shard applied on the fly, so peak host memory is the final arrays plus
// 1 MB per thread (scipy holds the inflated member, the matrix, the stacked matrix and the astype copy).
// File layout: a zip (stored or deflate, ZIP64 for members >= 4 GB) of `.npy` members (numpy format 1.0-3.0,
// little-endian, C order).  The writer side stays in numpy (npz_io.save_csr_npz): files remain scipy-loadable.
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <new>
#include <string>
#include <memory>
#include <vector>

#include "../../include/vsearch_b200.h"

namespace vs {
void set_error(const char *fmt, ...);
}

#define NPZ_REQUIRE(cond, code, ...)                 \
    do {                                             \
        if (!(cond)) {                               \
            vs::set_error(__VA_ARGS__);              \
            return (code);                           \
        }                                            \
    } while (0)

struct NpzMember {
    std::string name;        // without ".npy"
    uint16_t method = 0;     // 0 stored, 8 deflate
    uint64_t comp_size = 0, raw_size = 0, local_off = 0;
    uint32_t crc = 0;        // CRC-32 of the inflated member (central directory)
    // filled lazily by parse_npy_header
    bool parsed = false;
    int dtype = VS_NONE;     // VS_* code, or VS_NONE for dtypes the search path does not use
    int item = 0;            // bytes per element
    int ndim = 0;
    int64_t shape[4] = {0, 0, 0, 0};
    uint64_t header_len = 0; // bytes of the .npy header inside the member
    char descr[16] = "";
};

struct vs_npz {
    std::string path;
    std::vector<NpzMember> members;
};

namespace {

uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint64_t rd64(const uint8_t *p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

struct File {
    FILE *f = nullptr;
    ~File() { if (f) fclose(f); }
};

// Streams the raw (inflated) bytes of a member through `sink(ptr, n)`; stops early when sink returns false.
// A member that was streamed to its end is checked against the CRC-32 of the central directory.
template <typename Sink>
int stream_member(const vs_npz *z, const NpzMember &m, Sink &&sink) {
    uint32_t crc = (uint32_t)crc32(0L, Z_NULL, 0);
    uint64_t total = 0;
    File fh;
    fh.f = fopen(z->path.c_str(), "rb");
    NPZ_REQUIRE(fh.f, VS_ERR_INVALID, "%s: cannot open", z->path.c_str());
    uint8_t lh[30];
    NPZ_REQUIRE(fseeko(fh.f, (off_t)m.local_off, SEEK_SET) == 0 && fread(lh, 1, 30, fh.f) == 30 && rd32(lh) == 0x04034b50u,
                VS_ERR_INVALID, "%s: bad local header of member %s", z->path.c_str(), m.name.c_str());
    const uint64_t data_off = m.local_off + 30 + rd16(lh + 26) + rd16(lh + 28);
    NPZ_REQUIRE(fseeko(fh.f, (off_t)data_off, SEEK_SET) == 0, VS_ERR_INVALID, "%s: seek failed", z->path.c_str());
    constexpr size_t kBuf = 1 << 20;
    std::vector<uint8_t> in(kBuf), out(kBuf);
    uint64_t left = m.comp_size;
    if (m.method == 0) {
        while (left) {
            const size_t n = (size_t)(left < kBuf ? left : kBuf);
            NPZ_REQUIRE(fread(in.data(), 1, n, fh.f) == n, VS_ERR_INVALID, "%s: truncated member %s", z->path.c_str(), m.name.c_str());
            left -= n;
            crc = (uint32_t)crc32(crc, in.data(), (uInt)n);
            total += n;
            if (!sink(in.data(), n) && total < m.raw_size) return VS_OK;
        }
        NPZ_REQUIRE(total == m.raw_size && crc == m.crc, VS_ERR_INVALID, "%s: size / CRC mismatch in member %s", z->path.c_str(), m.name.c_str());
        return VS_OK;
    }
    NPZ_REQUIRE(m.method == 8, VS_ERR_UNSUPPORTED, "%s: member %s uses zip method %d (only stored / deflate)", z->path.c_str(),
                m.name.c_str(), (int)m.method);
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    NPZ_REQUIRE(inflateInit2(&zs, -MAX_WBITS) == Z_OK, VS_ERR_NOMEM, "zlib inflateInit2 failed");
    int zrc = Z_OK;
    bool stop = false;
    while (zrc != Z_STREAM_END && !stop) {
        if (zs.avail_in == 0) {
            const size_t n = (size_t)(left < kBuf ? left : kBuf);
            if (n == 0) break;
            if (fread(in.data(), 1, n, fh.f) != n) { inflateEnd(&zs); NPZ_REQUIRE(false, VS_ERR_INVALID, "%s: truncated member %s", z->path.c_str(), m.name.c_str()); }
            left -= n;
            zs.next_in = in.data();
            zs.avail_in = (uInt)n;
        }
        zs.next_out = out.data();
        zs.avail_out = (uInt)kBuf;
        zrc = inflate(&zs, Z_NO_FLUSH);
        if (zrc != Z_OK && zrc != Z_STREAM_END) { inflateEnd(&zs); NPZ_REQUIRE(false, VS_ERR_INVALID, "%s: corrupt deflate stream in member %s (zlib %d)", z->path.c_str(), m.name.c_str(), zrc); }
        const size_t produced = kBuf - zs.avail_out;
        crc = (uint32_t)crc32(crc, out.data(), (uInt)produced);
        total += produced;
        if (produced && !sink(out.data(), produced) && total < m.raw_size) stop = true;
    }
    inflateEnd(&zs);
    NPZ_REQUIRE(stop || (total == m.raw_size && crc == m.crc), VS_ERR_INVALID, "%s: size / CRC mismatch in member %s", z->path.c_str(),
                m.name.c_str());
    return VS_OK;
}

int parse_npy_header(const vs_npz *z, NpzMember &m) {
    if (m.parsed) return VS_OK;
    std::vector<uint8_t> head;
    int rc = stream_member(z, m, [&](const uint8_t *p, size_t n) {
        head.insert(head.end(), p, p + n);
        if (head.size() < 12) return true;
        const size_t hl = head[6] == 1 ? (size_t)rd16(&head[8]) + 10 : (size_t)rd32(&head[8]) + 12;
        return head.size() < hl;
    });
    if (rc) return rc;
    NPZ_REQUIRE(head.size() >= 12 && memcmp(head.data(), "\x93NUMPY", 6) == 0, VS_ERR_INVALID, "%s: member %s is not a .npy array",
                z->path.c_str(), m.name.c_str());
    const size_t hl = head[6] == 1 ? (size_t)rd16(&head[8]) + 10 : (size_t)rd32(&head[8]) + 12;
    NPZ_REQUIRE(head.size() >= hl, VS_ERR_INVALID, "%s: truncated .npy header in %s", z->path.c_str(), m.name.c_str());
    const std::string dict((const char *)&head[head[6] == 1 ? 10 : 12], (const char *)&head[hl]);
    auto value_after = [&](const char *key) -> std::string {
        size_t p = dict.find(key);
        if (p == std::string::npos) return "";
        p = dict.find(':', p);
        return p == std::string::npos ? "" : dict.substr(p + 1);
    };
    std::string d = value_after("'descr'");
    size_t q0 = d.find('\''), q1 = q0 == std::string::npos ? q0 : d.find('\'', q0 + 1);
    NPZ_REQUIRE(q1 != std::string::npos, VS_ERR_INVALID, "%s: no descr in the header of %s", z->path.c_str(), m.name.c_str());
    const std::string descr = d.substr(q0 + 1, q1 - q0 - 1);
    snprintf(m.descr, sizeof(m.descr), "%s", descr.c_str());
    NPZ_REQUIRE(value_after("'fortran_order'").find("True") > 8, VS_ERR_UNSUPPORTED, "%s: member %s is Fortran-ordered", z->path.c_str(),
                m.name.c_str());
    m.dtype = VS_NONE;
    m.item = descr.size() >= 3 ? atoi(descr.c_str() + 2) : 0;
    if (descr == "<i8") m.dtype = VS_I64;
    else if (descr == "<i4") m.dtype = VS_I32;
    else if (descr == "<u2") m.dtype = VS_U16;
    else if (descr == "<u4") m.dtype = VS_U32;
    else if (descr == "<f4") m.dtype = VS_F32;
    else if (descr == "<f2") m.dtype = VS_F16;
    else if (descr == "<f8") m.dtype = VS_F64;
    std::string sh = value_after("'shape'");
    size_t a = sh.find('('), b = sh.find(')');
    NPZ_REQUIRE(a != std::string::npos && b != std::string::npos, VS_ERR_INVALID, "%s: no shape in the header of %s", z->path.c_str(),
                m.name.c_str());
    m.ndim = 0;
    const char *c = sh.c_str() + a + 1, *end = sh.c_str() + b;
    while (c < end && m.ndim < 4) {
        while (c < end && (*c < '0' || *c > '9')) ++c;
        if (c >= end) break;
        m.shape[m.ndim++] = strtoll(c, (char **)&c, 10);
    }
    m.header_len = hl;
    m.parsed = true;
    return VS_OK;
}

NpzMember *find_member(vs_npz *z, const char *name) {
    for (auto &m : z->members)
        if (m.name == name) return &m;
    return nullptr;
}

// element-wise conversion of `n` source elements (src dtype S) into dst dtype code `dd`, adding `add` to integers
template <typename S>
bool convert_run(const S *src, size_t n, void *dst, int dd, size_t at, int64_t add) {
    switch (dd) {
        case VS_I64: { int64_t *d = (int64_t *)dst + at; for (size_t i = 0; i < n; ++i) d[i] = (int64_t)src[i] + add; return true; }
        case VS_I32: { int32_t *d = (int32_t *)dst + at; for (size_t i = 0; i < n; ++i) d[i] = (int32_t)((int64_t)src[i] + add); return true; }
        case VS_U32: { uint32_t *d = (uint32_t *)dst + at; for (size_t i = 0; i < n; ++i) d[i] = (uint32_t)((int64_t)src[i] + add); return true; }
        case VS_U16: { uint16_t *d = (uint16_t *)dst + at; for (size_t i = 0; i < n; ++i) d[i] = (uint16_t)((int64_t)src[i] + add); return true; }
        case VS_F32: { float *d = (float *)dst + at; for (size_t i = 0; i < n; ++i) d[i] = (float)src[i]; return true; }
        case VS_F16: { __half *d = (__half *)dst + at; for (size_t i = 0; i < n; ++i) d[i] = __float2half((float)src[i]); return true; }
        default: return false;
    }
}
bool convert_run_half(const __half *src, size_t n, void *dst, int dd, size_t at) {
    if (dd == VS_F16) { memcpy((__half *)dst + at, src, n * 2); return true; }
    if (dd == VS_F32) { float *d = (float *)dst + at; for (size_t i = 0; i < n; ++i) d[i] = __half2float(src[i]); return true; }
    return false;
}

}  // namespace

namespace {

int npz_open_impl(const char *path, vs_npz **out) {
    NPZ_REQUIRE(path && out, VS_ERR_INVALID, "vs_npz_open: NULL argument");
    File fh;
    fh.f = fopen(path, "rb");
    NPZ_REQUIRE(fh.f, VS_ERR_INVALID, "%s: cannot open", path);
    NPZ_REQUIRE(fseeko(fh.f, 0, SEEK_END) == 0, VS_ERR_INVALID, "%s: seek failed", path);
    const uint64_t fsize = (uint64_t)ftello(fh.f);
    const uint64_t tail = fsize < 65536 + 22 + 20 ? fsize : 65536 + 22 + 20;
    std::vector<uint8_t> buf((size_t)tail);
    NPZ_REQUIRE(fseeko(fh.f, (off_t)(fsize - tail), SEEK_SET) == 0 && fread(buf.data(), 1, (size_t)tail, fh.f) == tail, VS_ERR_INVALID,
                "%s: read failed", path);
    int64_t e = -1;
    for (int64_t i = (int64_t)tail - 22; i >= 0; --i)
        if (rd32(&buf[(size_t)i]) == 0x06054b50u) { e = i; break; }
    NPZ_REQUIRE(e >= 0, VS_ERR_INVALID, "%s: not a zip archive (no end-of-central-directory record)", path);
    uint64_t n_entries = rd16(&buf[(size_t)e + 10]), cd_size = rd32(&buf[(size_t)e + 12]), cd_off = rd32(&buf[(size_t)e + 16]);
    if (n_entries == 0xffff || cd_size == 0xffffffffu || cd_off == 0xffffffffu) {   // ZIP64
        NPZ_REQUIRE(e >= 20 && rd32(&buf[(size_t)e - 20]) == 0x07064b50u, VS_ERR_INVALID, "%s: ZIP64 locator missing", path);
        const uint64_t e64 = rd64(&buf[(size_t)e - 20 + 8]);
        uint8_t r[56];
        NPZ_REQUIRE(fseeko(fh.f, (off_t)e64, SEEK_SET) == 0 && fread(r, 1, 56, fh.f) == 56 && rd32(r) == 0x06064b50u, VS_ERR_INVALID,
                    "%s: bad ZIP64 end-of-central-directory record", path);
        n_entries = rd64(r + 32); cd_size = rd64(r + 40); cd_off = rd64(r + 48);
    }
    // sizes come from the file: bound them by the file before they size a buffer or an index
    NPZ_REQUIRE(cd_off <= fsize && cd_size <= fsize - cd_off, VS_ERR_INVALID, "%s: central directory lies outside the file", path);
    NPZ_REQUIRE(n_entries <= cd_size / 46, VS_ERR_INVALID, "%s: central directory too small for its entry count", path);
    std::vector<uint8_t> cd((size_t)cd_size);
    NPZ_REQUIRE(fseeko(fh.f, (off_t)cd_off, SEEK_SET) == 0 && fread(cd.data(), 1, (size_t)cd_size, fh.f) == cd_size, VS_ERR_INVALID,
                "%s: cannot read the central directory", path);
    std::unique_ptr<vs_npz> z(new vs_npz);
    z->path = path;
    size_t p = 0;
    for (uint64_t i = 0; i < n_entries; ++i) {
        NPZ_REQUIRE(p + 46 <= cd.size() && rd32(&cd[p]) == 0x02014b50u, VS_ERR_INVALID, "%s: corrupt central directory", path);
        // name, extra field and comment lengths are file data too: the whole entry must lie inside the directory
        NPZ_REQUIRE(p + 46 + (size_t)rd16(&cd[p + 28]) + rd16(&cd[p + 30]) + rd16(&cd[p + 32]) <= cd.size(), VS_ERR_INVALID,
                    "%s: central directory entry %llu runs past the directory", path, (unsigned long long)i);
        NpzMember m;
        m.method = rd16(&cd[p + 10]);
        m.crc = rd32(&cd[p + 16]);
        m.comp_size = rd32(&cd[p + 20]); m.raw_size = rd32(&cd[p + 24]);
        const uint16_t nl = rd16(&cd[p + 28]), xl = rd16(&cd[p + 30]), cl = rd16(&cd[p + 32]);
        m.local_off = rd32(&cd[p + 42]);
        std::string name((const char *)&cd[p + 46], nl);
        size_t x = p + 46 + nl;
        const size_t xend = x + xl;
        while (x + 4 <= xend) {   // ZIP64 extended information: only the fields that overflowed, in this order
            const uint16_t id = rd16(&cd[x]), sz = rd16(&cd[x + 2]);
            if (x + 4 + (size_t)sz > xend) break;   // a field that claims more than the extra block holds
            if (id == 0x0001) {
                size_t f = x + 4;
                if (m.raw_size == 0xffffffffu && f + 8 <= x + 4 + sz) { m.raw_size = rd64(&cd[f]); f += 8; }
                if (m.comp_size == 0xffffffffu && f + 8 <= x + 4 + sz) { m.comp_size = rd64(&cd[f]); f += 8; }
                if (m.local_off == 0xffffffffu && f + 8 <= x + 4 + sz) { m.local_off = rd64(&cd[f]); f += 8; }
            }
            x += 4 + (size_t)sz;
        }
        if (name.size() > 4 && name.compare(name.size() - 4, 4, ".npy") == 0) name.resize(name.size() - 4);
        m.name = name;
        z->members.push_back(m);
        p += 46 + (size_t)nl + xl + cl;
    }
    *out = z.release();
    return VS_OK;
}

int npz_member_info_impl(vs_npz *z, const char *name, int *dtype, int *ndim, int64_t *shape4, int64_t *n_elems) {
    NPZ_REQUIRE(z && name, VS_ERR_INVALID, "vs_npz_member_info: NULL argument");
    NpzMember *m = find_member(z, name);
    NPZ_REQUIRE(m, VS_ERR_INVALID, "%s: no member %s", z->path.c_str(), name);
    int rc = parse_npy_header(z, *m);
    if (rc) return rc;
    int64_t n = 1;
    for (int i = 0; i < m->ndim; ++i) n *= m->shape[i];
    if (dtype) *dtype = m->dtype;
    if (ndim) *ndim = m->ndim;
    if (shape4) for (int i = 0; i < 4; ++i) shape4[i] = i < m->ndim ? m->shape[i] : 0;
    if (n_elems) *n_elems = n;
    return VS_OK;
}

// `n` elements of source dtype `sd` at p (aligned) -> dst[at ...] in dtype `dd`
bool convert_any(int sd, const uint8_t *p, size_t n, void *dst, int dd, size_t at, int64_t add) {
    switch (sd) {
        case VS_I64: return convert_run((const int64_t *)p, n, dst, dd, at, add);
        case VS_I32: return convert_run((const int32_t *)p, n, dst, dd, at, add);
        case VS_U32: return convert_run((const uint32_t *)p, n, dst, dd, at, add);
        case VS_U16: return convert_run((const uint16_t *)p, n, dst, dd, at, add);
        case VS_F32: return convert_run((const float *)p, n, dst, dd, at, 0);
        case VS_F64: return convert_run((const double *)p, n, dst, dd, at, 0);
        default: return convert_run_half((const __half *)p, n, dst, dd, at);
    }
}

size_t dtype_bytes(int dt) {
    switch (dt) {
        case VS_I64: case VS_F64: return 8;
        case VS_I32: case VS_U32: case VS_F32: return 4;
        default: return 2;
    }
}

// converted elements go straight to their place in a host array
struct HostSink {
    void *dst; int dd; size_t written = 0;
    bool put(int sd, const uint8_t *p, size_t n, int64_t add) {
        const bool ok = convert_any(sd, p, n, dst, dd, written, add);
        written += n;
        return ok;
    }
    bool finish() { return true; }
};

// converted elements go through two pinned staging buffers to their place in a DEVICE array: while one buffer is on
// its way (cudaMemcpyAsync on the sink's own stream) the inflating thread fills the other
struct DeviceSink {
    uint8_t *d_dst; int dd; cudaStream_t st;
    uint8_t *buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    size_t cap = 0, fill = 0, d_off = 0, written = 0;
    int cur = 0;
    bool failed = false;
    bool init(size_t bytes) {
        cap = bytes;
        for (int i = 0; i < 2; ++i)
            if (cudaHostAlloc((void **)&buf[i], cap, cudaHostAllocDefault) != cudaSuccess ||
                cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) return false;
        return true;
    }
    ~DeviceSink() {
        for (int i = 0; i < 2; ++i) { if (buf[i]) cudaFreeHost(buf[i]); if (ev[i]) cudaEventDestroy(ev[i]); }
    }
    bool flush() {
        if (fill == 0) return true;
        if (cudaMemcpyAsync(d_dst + d_off, buf[cur], fill, cudaMemcpyHostToDevice, st) != cudaSuccess) return !(failed = true);
        if (cudaEventRecord(ev[cur], st) != cudaSuccess) return !(failed = true);
        d_off += fill;
        fill = 0;
        cur ^= 1;
        return cudaEventSynchronize(ev[cur]) == cudaSuccess || !(failed = true);   // the other buffer's last copy has left it
    }
    bool put(int sd, const uint8_t *p, size_t n, int64_t add) {
        const size_t di = dtype_bytes(dd), si = dtype_bytes(sd);
        while (n) {
            const size_t room = (cap - fill) / di, c = room < n ? room : n;
            if (!convert_any(sd, p, c, buf[cur], dd, fill / di, add)) return false;
            fill += c * di; written += c; p += c * si; n -= c;
            if (fill + di > cap && !flush()) return false;
        }
        return true;
    }
    bool finish() { return flush() && cudaStreamSynchronize(st) == cudaSuccess; }
};

template <typename Sink>
int npz_read_core(vs_npz *z, const char *name, Sink &out, int dst_dtype, int64_t skip_elems, int64_t n_elems, int64_t add_offset) {
    NpzMember *m = find_member(z, name);
    NPZ_REQUIRE(m, VS_ERR_INVALID, "%s: no member %s", z->path.c_str(), name);
    int rc = parse_npy_header(z, *m);
    if (rc) return rc;
    NPZ_REQUIRE(m->dtype != VS_NONE, VS_ERR_UNSUPPORTED, "%s: member %s has dtype %s (int32/int64/uint16/uint32/float16/float32/float64 only)",
                z->path.c_str(), name, m->descr);
    int64_t total = 1;
    for (int i = 0; i < m->ndim; ++i) total *= m->shape[i];
    NPZ_REQUIRE(skip_elems >= 0 && n_elems >= 0 && skip_elems + n_elems <= total, VS_ERR_INVALID,
                "%s: member %s holds %lld elements, asked for [%lld, %lld)", z->path.c_str(), name, (long long)total,
                (long long)skip_elems, (long long)(skip_elems + n_elems));
    const bool src_int = m->dtype == VS_I64 || m->dtype == VS_I32 || m->dtype == VS_U16 || m->dtype == VS_U32;
    const bool dst_int = dst_dtype == VS_I64 || dst_dtype == VS_I32 || dst_dtype == VS_U16 || dst_dtype == VS_U32;
    NPZ_REQUIRE(src_int == dst_int && (dst_int || dst_dtype == VS_F32 || dst_dtype == VS_F16), VS_ERR_INVALID,
                "%s: cannot convert member %s (%s) to dtype code %d", z->path.c_str(), name, m->descr, dst_dtype);
    const size_t item = (size_t)m->item;
    uint64_t pos = 0;                          // raw bytes consumed so far
    const uint64_t first = m->header_len + (uint64_t)skip_elems * item, last = first + (uint64_t)n_elems * item;
    uint8_t carry[8];
    size_t n_carry = 0;
    bool ok = true;
    auto emit = [&](const uint8_t *p, size_t n_el) { ok = ok && out.put(m->dtype, p, n_el, add_offset); };
    rc = stream_member(z, *m, [&](const uint8_t *p, size_t n) {
        uint64_t lo = pos, hi = pos + n;
        pos = hi;
        if (hi <= first) return true;
        if (lo < first) { p += first - lo; lo = first; }
        if (hi > last) hi = last;
        size_t len = (size_t)(hi - lo);
        if (n_carry) {   // finish the element split across two buffers (inflate output is not element-aligned)
            const size_t need = item - n_carry, take = need < len ? need : len;
            memcpy(carry + n_carry, p, take);
            n_carry += take; p += take; len -= take;
            if (n_carry == item) { alignas(8) uint8_t one[8]; memcpy(one, carry, item); emit(one, 1); n_carry = 0; }
        }
        const size_t whole = len / item;
        if (whole) {
            if (((uintptr_t)p & (item - 1)) == 0) emit(p, whole);
            else {   // unaligned tail of the buffer after a carry: go through an aligned bounce buffer
                alignas(8) uint8_t tmp[4096];
                size_t done = 0;
                while (done < whole) {
                    const size_t c = (whole - done) < 4096 / item ? (whole - done) : 4096 / item;
                    memcpy(tmp, p + done * item, c * item);
                    emit(tmp, c);
                    done += c;
                }
            }
        }
        const size_t rest = len - whole * item;
        if (rest) { memcpy(carry, p + whole * item, rest); n_carry = rest; }
        return ok && pos < last;
    });
    if (rc) return rc;
    NPZ_REQUIRE(ok && out.finish(), VS_ERR_INVALID, "%s: conversion / transfer of member %s failed", z->path.c_str(), name);
    NPZ_REQUIRE((int64_t)out.written == n_elems, VS_ERR_INVALID, "%s: member %s ended after %lld of %lld elements", z->path.c_str(), name,
                (long long)out.written, (long long)n_elems);
    return VS_OK;
}

int npz_read_impl(vs_npz *z, const char *name, void *dst, int dst_dtype, int64_t skip_elems, int64_t n_elems, int64_t add_offset) {
    NPZ_REQUIRE(z && name && (dst || n_elems == 0), VS_ERR_INVALID, "vs_npz_read: NULL argument");
    HostSink out{dst, dst_dtype};
    return npz_read_core(z, name, out, dst_dtype, skip_elems, n_elems, add_offset);
}

// raw bytes of a small member's array data (format / shape / _is_array), at most `cap`
int npz_read_small(vs_npz *z, const char *name, uint8_t *dst, size_t cap, size_t *got) {
    NpzMember *m = find_member(z, name);
    NPZ_REQUIRE(m, VS_ERR_INVALID, "%s: no member %s", z->path.c_str(), name);
    int rc = parse_npy_header(z, *m);
    if (rc) return rc;
    std::vector<uint8_t> all;
    rc = stream_member(z, *m, [&](const uint8_t *p, size_t n) { all.insert(all.end(), p, p + n); return all.size() < m->header_len + cap; });
    if (rc) return rc;
    const size_t n = all.size() > m->header_len ? all.size() - (size_t)m->header_len : 0;
    *got = n < cap ? n : cap;
    memcpy(dst, all.data() + m->header_len, *got);
    return VS_OK;
}

}  // namespace

// The ABI never throws: allocation failures inside the readers come back as VS_ERR_NOMEM.
#define NPZ_NOTHROW(expr)                                                    \
    try {                                                                    \
        return (expr);                                                       \
    } catch (const std::bad_alloc &) {                                       \
        vs::set_error("out of host memory in the .npz reader");              \
        return VS_ERR_NOMEM;                                                 \
    } catch (...) {                                                          \
        vs::set_error("unexpected C++ exception in the .npz reader");        \
        return VS_ERR_INVALID;                                               \
    }

extern "C" {

int vs_npz_open(const char *path, vs_npz **out) { NPZ_NOTHROW(npz_open_impl(path, out)) }

int vs_npz_close(vs_npz *z) {
    delete z;
    return VS_OK;
}

int vs_npz_member_info(vs_npz *z, const char *name, int *dtype, int *ndim, int64_t *shape4, int64_t *n_elems) {
    NPZ_NOTHROW(npz_member_info_impl(z, name, dtype, ndim, shape4, n_elems))
}

int vs_npz_read(vs_npz *z, const char *name, void *dst, int dst_dtype, int64_t skip_elems, int64_t n_elems, int64_t add_offset) {
    NPZ_NOTHROW(npz_read_impl(z, name, dst, dst_dtype, skip_elems, n_elems, add_offset))
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Loader: shard files -> a device index, without a host copy of the index.  Upstream
// (SparseIndex.init_index, index.py:163-179) inflates every shard on one Python thread into scipy matrices, slices,
// vstacks, converts and only then copies to the device: minutes and more than twice the index in host memory at 21M
// rows.  Here every (shard, member) pair is a job of a small thread pool: the member is inflated in 1 MB pieces,
// narrowed on the fly (int64 / int32 columns -> uint16 when the vocabulary allows, float64 / float32 values ->
// float16 when asked, row pointers + the shard's entry offset -> int64) into one of two pinned staging buffers and
// sent to its final place in the device CSR arrays with cudaMemcpyAsync while the thread inflates the next piece.
// The `[:, shift:]` column slice and the row concatenation cost nothing on the host: the shift is applied by the
// index build on the GPU (build_index.cu drops and renumbers while it lays the rows out), concatenation is where the
// pieces land.  Peak host memory: 2 x 4 MB pinned per thread.
#include <algorithm>
#include <atomic>
#include <mutex>
#include <thread>

#include "index.cuh"

namespace vs {
__global__ void all_ones_kernel(const void *v, int dtype, uint64_t n, int *not_ones) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (; i < n; i += stride) {
        const float x = dtype == VS_F16 ? __half2float(((const __half *)v)[i]) : ((const float *)v)[i];
        bad |= (x != 1.0f);
    }
    if (bad) atomicExch(not_ones, 1);
}
}  // namespace vs

namespace {

struct LoadJob {
    int shard;
    const char *member;
    uint8_t *d_dst;
    int dst_dtype;
    int64_t skip, n, add;
    uint64_t bytes;
};

int load_npz_impl(int device, const char *const *paths, int n_paths, int shift, int value_dtype, int binary_if_ones, int threads,
                  void *stream, vs_index **out, int64_t *shard_rows) {
    NPZ_REQUIRE(out != nullptr, VS_ERR_INVALID, "out is NULL");
    *out = nullptr;
    NPZ_REQUIRE(paths != nullptr && n_paths >= 1 && shift >= 0, VS_ERR_INVALID, "vs_index_load_npz: bad argument");
    NPZ_REQUIRE(value_dtype == VS_F32 || value_dtype == VS_F16 || value_dtype == VS_BF16, VS_ERR_INVALID, "value dtype must be f32 / f16 / bf16");
    struct Shard { std::unique_ptr<vs_npz> z; int64_t rows = 0, nnz = 0, row_off = 0, nnz_off = 0; };
    std::vector<Shard> sh((size_t)n_paths);
    int64_t n_rows = 0, nnz = 0, n_cols_file = -1;
    for (int i = 0; i < n_paths; ++i) {
        vs_npz *z = nullptr;
        int rc = npz_open_impl(paths[i], &z);
        if (rc) return rc;
        sh[i].z.reset(z);
        uint8_t fmt[8] = {0};
        size_t got = 0;
        rc = npz_read_small(z, "format", fmt, 3, &got);
        if (rc) return rc;
        NPZ_REQUIRE(got == 3 && memcmp(fmt, "csr", 3) == 0, VS_ERR_INVALID, "%s: expected a CSR .npz (format member is not b'csr')", paths[i]);
        int64_t shape[2] = {0, 0};
        rc = npz_read_impl(z, "shape", shape, VS_I64, 0, 2, 0);
        if (rc) return rc;
        int64_t n_ptr = 0, n_idx = 0, n_dat = 0;
        rc = npz_member_info_impl(z, "indptr", nullptr, nullptr, nullptr, &n_ptr);
        if (rc == VS_OK) rc = npz_member_info_impl(z, "indices", nullptr, nullptr, nullptr, &n_idx);
        if (rc == VS_OK) rc = npz_member_info_impl(z, "data", nullptr, nullptr, nullptr, &n_dat);
        if (rc) return rc;
        NPZ_REQUIRE(n_ptr == shape[0] + 1 && n_idx == n_dat && shape[0] >= 0 && shape[1] >= 1, VS_ERR_INVALID, "%s: inconsistent CSR members", paths[i]);
        if (n_cols_file < 0) n_cols_file = shape[1];
        NPZ_REQUIRE(shape[1] == n_cols_file, VS_ERR_INVALID, "%s: column count %lld differs from previous shards (%lld)", paths[i],
                    (long long)shape[1], (long long)n_cols_file);
        sh[i].rows = shape[0]; sh[i].nnz = n_idx; sh[i].row_off = n_rows; sh[i].nnz_off = nnz;
        if (shard_rows) shard_rows[i] = shape[0];
        n_rows += shape[0];
        nnz += n_idx;
    }
    NPZ_REQUIRE(n_cols_file - shift >= 1, VS_ERR_INVALID, "shift=%d leaves no columns of %lld", shift, (long long)n_cols_file);
    NPZ_REQUIRE(cudaSetDevice(device) == cudaSuccess, VS_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    cudaStream_t st = (cudaStream_t)stream;
    const int col_dtype = n_cols_file <= 65535 ? VS_U16 : VS_I32;
    const int stage_val = value_dtype == VS_F16 ? VS_F16 : VS_F32;   // upstream's fp16=True: astype(float16) while loading
    const size_t csz = dtype_bytes(col_dtype), vsz = dtype_bytes(stage_val);
    int64_t *d_crow = nullptr;
    uint8_t *d_col = nullptr, *d_val = nullptr;
    int *d_flag = nullptr;
    auto cleanup = [&]() { cudaFree(d_crow); cudaFree(d_col); cudaFree(d_val); cudaFree(d_flag); };
    if (cudaMalloc(&d_crow, (size_t)(n_rows + 1) * 8) != cudaSuccess || cudaMalloc(&d_col, nnz ? (size_t)nnz * csz : 16) != cudaSuccess ||
        cudaMalloc(&d_val, nnz ? (size_t)nnz * vsz : 16) != cudaSuccess || cudaMalloc(&d_flag, 4) != cudaSuccess) {
        cleanup();
        NPZ_REQUIRE(false, VS_ERR_NOMEM, "out of device memory for the CSR staging arrays (%lld rows, %lld entries)", (long long)n_rows, (long long)nnz);
    }
    cudaMemsetAsync(d_crow, 0, 8, st);
    cudaMemsetAsync(d_flag, 0, 4, st);
    // ---- jobs, largest first
    std::vector<LoadJob> jobs;
    for (int i = 0; i < n_paths; ++i) {
        jobs.push_back({i, "indices", d_col + (size_t)sh[i].nnz_off * csz, col_dtype, 0, sh[i].nnz, 0, (uint64_t)sh[i].nnz * 8});
        jobs.push_back({i, "data", d_val + (size_t)sh[i].nnz_off * vsz, stage_val, 0, sh[i].nnz, 0, (uint64_t)sh[i].nnz * 4});
        // drop each shard's leading 0, add the entries of the shards before it
        jobs.push_back({i, "indptr", (uint8_t *)(d_crow + sh[i].row_off + 1), VS_I64, 1, sh[i].rows, sh[i].nnz_off, (uint64_t)sh[i].rows * 8});
    }
    std::sort(jobs.begin(), jobs.end(), [](const LoadJob &a, const LoadJob &b) { return a.bytes > b.bytes; });
    int n_thr = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (n_thr < 1) n_thr = 1;
    if (n_thr > (int)jobs.size()) n_thr = (int)jobs.size();
    std::atomic<size_t> next{0};
    std::mutex err_mu;
    int first_rc = VS_OK;
    std::string first_err;
    auto worker = [&]() {
        auto fail = [&](int rc, const char *msg) {
            std::lock_guard<std::mutex> g(err_mu);
            if (first_rc == VS_OK) { first_rc = rc; first_err = msg; }
        };
        if (cudaSetDevice(device) != cudaSuccess) { fail(VS_ERR_CUDA, "cudaSetDevice failed in a loader thread"); return; }
        cudaStream_t cs = nullptr;
        if (cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess) { fail(VS_ERR_CUDA, "cudaStreamCreate failed in a loader thread"); return; }
        {
            DeviceSink sink{nullptr, VS_I64, cs};
            if (!sink.init(4u << 20)) fail(VS_ERR_NOMEM, "cannot allocate pinned staging buffers");
            else
                for (;;) {
                    const size_t j = next.fetch_add(1);
                    if (j >= jobs.size() || first_rc != VS_OK) break;
                    const LoadJob &job = jobs[j];
                    sink.d_dst = job.d_dst; sink.dd = job.dst_dtype; sink.fill = 0; sink.d_off = 0; sink.written = 0;
                    int rc = VS_OK;
                    try {
                        rc = npz_read_core(sh[job.shard].z.get(), job.member, sink, job.dst_dtype, job.skip, job.n, job.add);
                    } catch (...) {
                        rc = VS_ERR_NOMEM;
                        vs::set_error("out of host memory in a loader thread");
                    }
                    if (rc != VS_OK) fail(rc, vs_last_error());   // the message is thread-local: carry it over
                }
        }
        cudaStreamSynchronize(cs);
        cudaStreamDestroy(cs);
    };
    {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_thr; ++t) pool.emplace_back(worker);
        for (auto &t : pool) t.join();
    }
    if (first_rc != VS_OK) { cleanup(); NPZ_REQUIRE(false, first_rc, "%s", first_err.c_str()); }
    // ---- bag-of-token files store a 1 per entry: keep column ids only
    bool binary = false;
    if (binary_if_ones) {
        int h_flag = 0;
        if (nnz) vs::all_ones_kernel<<<1024, 256, 0, st>>>(d_val, stage_val, (uint64_t)nnz, d_flag);
        cudaError_t e = cudaMemcpyAsync(&h_flag, d_flag, 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { cleanup(); NPZ_REQUIRE(false, VS_ERR_CUDA, "all-ones check failed: %s", cudaGetErrorString(e)); }
        binary = h_flag == 0;
    }
    int rc = vs::create_csr_from_device(device, n_rows, n_cols_file - shift, nnz, d_crow, VS_I64, d_col, col_dtype, binary ? nullptr : d_val,
                                        binary ? VS_NONE : stage_val, binary ? VS_NONE : value_dtype, shift, st, out);
    cleanup();
    return rc;
}

}  // namespace

extern "C" int vs_index_load_npz(int device, const char *const *paths, int n_paths, int shift, int value_dtype, int binary_if_ones,
                                 int threads, void *stream, vs_index **out, int64_t *shard_rows) {
    NPZ_NOTHROW(load_npz_impl(device, paths, n_paths, shift, value_dtype, binary_if_ones, threads, stream, out, shard_rows))
}

// ---------------------------------------------------------------------------------------------------------------
// Writer: a zip of deflated `.npy` members (what scipy.sparse.save_npz / numpy.savez_compressed produce,
// reference index.py:195-197), compressed by a thread pool.  A member's bytes (the .npy header the caller built +
// the array data) are cut into 4 MB blocks; every block is deflated independently (raw deflate, Z_SYNC_FLUSH, the
// last one Z_FINISH) and the pieces are concatenated in order: sync-flushed deflate blocks end byte-aligned without
// the final bit, so the concatenation is ONE valid deflate stream that zipfile / scipy / upstream's loader inflate as
// usual (the pigz construction).  CRC-32 per block, combined in order.  ZIP64 records for members or offsets >= 4 GB.
#include <condition_variable>
#include <mutex>
#include <thread>

namespace {

constexpr size_t kWriteBlock = 4u << 20;

struct OutBlock {
    std::vector<uint8_t> comp;
    uint32_t crc = 0;
    size_t raw = 0;
    bool done = false;
};

void put16(std::vector<uint8_t> &v, uint16_t x) { v.push_back((uint8_t)x); v.push_back((uint8_t)(x >> 8)); }
void put32(std::vector<uint8_t> &v, uint32_t x) { for (int i = 0; i < 4; ++i) v.push_back((uint8_t)(x >> (8 * i))); }
void put64(std::vector<uint8_t> &v, uint64_t x) { for (int i = 0; i < 8; ++i) v.push_back((uint8_t)(x >> (8 * i))); }

// byte `at` of the member = header bytes followed by data bytes
void copy_span(const vs_npz_member_in &m, uint64_t at, size_t n, uint8_t *dst) {
    size_t done = 0;
    if (at < (uint64_t)m.header_bytes) {
        const size_t h = (size_t)std::min<uint64_t>(n, (uint64_t)m.header_bytes - at);
        memcpy(dst, (const uint8_t *)m.header + at, h);
        done = h;
    }
    if (done < n) memcpy(dst + done, (const uint8_t *)m.data + (at + done - (uint64_t)m.header_bytes), n - done);
}

bool deflate_block(const uint8_t *src, size_t n, bool last, int level, std::vector<uint8_t> &out) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, level, Z_DEFLATED, -MAX_WBITS, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    out.resize(deflateBound(&zs, (uLong)n) + 16);
    zs.next_in = const_cast<uint8_t *>(src);
    zs.avail_in = (uInt)n;
    zs.next_out = out.data();
    zs.avail_out = (uInt)out.size();
    const int rc = deflate(&zs, last ? Z_FINISH : Z_SYNC_FLUSH);
    const bool ok = last ? rc == Z_STREAM_END : (rc == Z_OK && zs.avail_in == 0);
    out.resize(out.size() - zs.avail_out);
    deflateEnd(&zs);
    return ok;
}

int npz_write_impl(const char *path, const vs_npz_member_in *members, int n_members, int level, int threads) {
    NPZ_REQUIRE(path && members && n_members > 0, VS_ERR_INVALID, "vs_npz_write: bad argument");
    if (threads < 1) threads = 1;
    if (level < 0 || level > 9) level = 6;
    // test hook: write ZIP64 records for every member (the path a 21M-row index takes: its members exceed 4 GB)
    const bool force_z64 = getenv("VSEARCH_B200_NPZ_FORCE_ZIP64") != nullptr;
    File fh;
    fh.f = fopen(path, "wb");
    NPZ_REQUIRE(fh.f, VS_ERR_INVALID, "%s: cannot create", path);
    struct Entry { std::string name; uint32_t crc; uint64_t comp, raw, off; bool z64; };
    std::vector<Entry> dir;
    for (int mi = 0; mi < n_members; ++mi) {
        const vs_npz_member_in &m = members[mi];
        NPZ_REQUIRE(m.name && m.header_bytes >= 0 && m.data_bytes >= 0 && (m.header || m.header_bytes == 0) && (m.data || m.data_bytes == 0),
                    VS_ERR_INVALID, "vs_npz_write: bad member %d", mi);
        Entry e;
        e.name = std::string(m.name) + ".npy";
        e.raw = (uint64_t)m.header_bytes + (uint64_t)m.data_bytes;
        e.off = (uint64_t)ftello(fh.f);
        e.z64 = force_z64 || e.raw >= 0x7fffffffull || e.off >= 0x7fffffffull;
        e.crc = 0; e.comp = 0;
        // local header (sizes patched after the data is written)
        std::vector<uint8_t> lh;
        put32(lh, 0x04034b50u); put16(lh, e.z64 ? 45 : 20); put16(lh, 0); put16(lh, 8); put16(lh, 0); put16(lh, 0x21);
        put32(lh, 0); put32(lh, 0); put32(lh, 0);
        put16(lh, (uint16_t)e.name.size()); put16(lh, e.z64 ? 20 : 0);
        lh.insert(lh.end(), e.name.begin(), e.name.end());
        if (e.z64) { put16(lh, 1); put16(lh, 16); put64(lh, 0); put64(lh, 0); }
        NPZ_REQUIRE(fwrite(lh.data(), 1, lh.size(), fh.f) == lh.size(), VS_ERR_INVALID, "%s: write failed", path);
        // parallel deflate, in-order write
        const uint64_t n_blocks = e.raw == 0 ? 1 : (e.raw + kWriteBlock - 1) / kWriteBlock;
        const uint64_t window = (uint64_t)threads * 2;
        std::vector<OutBlock> ring((size_t)window);
        std::mutex mu;
        std::condition_variable cv;
        uint64_t next_claim = 0, next_write = 0;
        bool failed = false;
        auto worker = [&]() {
            std::vector<uint8_t> raw(kWriteBlock);
            for (;;) {
                uint64_t b;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return failed || next_claim >= n_blocks || next_claim < next_write + window; });
                    if (failed || next_claim >= n_blocks) return;
                    b = next_claim++;
                }
                const uint64_t at = b * kWriteBlock;
                const size_t n = (size_t)std::min<uint64_t>(kWriteBlock, e.raw - at);
                copy_span(m, at, n, raw.data());
                OutBlock ob;
                ob.raw = n;
                ob.crc = (uint32_t)crc32(0L, raw.data(), (uInt)n);
                const bool ok = deflate_block(raw.data(), n, b + 1 == n_blocks, level, ob.comp);
                {
                    std::lock_guard<std::mutex> lk(mu);
                    if (!ok) failed = true;
                    ob.done = true;
                    ring[(size_t)(b % window)] = std::move(ob);
                }
                cv.notify_all();
            }
        };
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(worker);
        bool io_ok = true;
        while (next_write < n_blocks) {
            OutBlock ob;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return failed || ring[(size_t)(next_write % window)].done; });
                if (failed) break;
                ob = std::move(ring[(size_t)(next_write % window)]);
                ring[(size_t)(next_write % window)] = OutBlock();
            }
            if (fwrite(ob.comp.data(), 1, ob.comp.size(), fh.f) != ob.comp.size()) io_ok = false;
            e.crc = next_write == 0 ? ob.crc : (uint32_t)crc32_combine(e.crc, ob.crc, (z_off_t)ob.raw);
            e.comp += ob.comp.size();
            {
                std::lock_guard<std::mutex> lk(mu);
                ++next_write;
                if (!io_ok) failed = true;
            }
            cv.notify_all();
            if (!io_ok) break;
        }
        { std::lock_guard<std::mutex> lk(mu); if (next_write < n_blocks) failed = true; }
        cv.notify_all();
        for (auto &t : pool) t.join();
        NPZ_REQUIRE(!failed, VS_ERR_INVALID, "%s: compressing or writing member %s failed", path, m.name);
        // patch the local header
        const uint64_t end = (uint64_t)ftello(fh.f);
        std::vector<uint8_t> fix;
        put32(fix, e.crc);
        if (e.z64) { put32(fix, 0xffffffffu); put32(fix, 0xffffffffu); } else { put32(fix, (uint32_t)e.comp); put32(fix, (uint32_t)e.raw); }
        NPZ_REQUIRE(fseeko(fh.f, (off_t)(e.off + 14), SEEK_SET) == 0 && fwrite(fix.data(), 1, fix.size(), fh.f) == fix.size(), VS_ERR_INVALID,
                    "%s: write failed", path);
        if (e.z64) {
            std::vector<uint8_t> x;
            put64(x, e.raw); put64(x, e.comp);
            NPZ_REQUIRE(fseeko(fh.f, (off_t)(e.off + 30 + e.name.size() + 4), SEEK_SET) == 0 && fwrite(x.data(), 1, 16, fh.f) == 16, VS_ERR_INVALID,
                        "%s: write failed", path);
        }
        NPZ_REQUIRE(fseeko(fh.f, (off_t)end, SEEK_SET) == 0, VS_ERR_INVALID, "%s: seek failed", path);
        dir.push_back(e);
    }
    // central directory
    const uint64_t cd_off = (uint64_t)ftello(fh.f);
    std::vector<uint8_t> cd;
    for (const Entry &e : dir) {
        const bool z64 = e.z64 || e.comp >= 0xffffffffull;
        put32(cd, 0x02014b50u); put16(cd, 45); put16(cd, z64 ? 45 : 20); put16(cd, 0); put16(cd, 8); put16(cd, 0); put16(cd, 0x21);
        put32(cd, e.crc);
        put32(cd, z64 ? 0xffffffffu : (uint32_t)e.comp); put32(cd, z64 ? 0xffffffffu : (uint32_t)e.raw);
        put16(cd, (uint16_t)e.name.size()); put16(cd, z64 ? 28 : 0); put16(cd, 0); put16(cd, 0); put16(cd, 0);
        put32(cd, 0x01800000u);   // external attributes: regular file 0600, like numpy
        put32(cd, z64 ? 0xffffffffu : (uint32_t)e.off);
        cd.insert(cd.end(), e.name.begin(), e.name.end());
        if (z64) { put16(cd, 1); put16(cd, 24); put64(cd, e.raw); put64(cd, e.comp); put64(cd, e.off); }
    }
    const uint64_t cd_size = cd.size();
    const bool big = force_z64 || cd_off >= 0xffffffffull || dir.size() >= 0xffff;
    if (big) {
        put32(cd, 0x06064b50u); put64(cd, 44); put16(cd, 45); put16(cd, 45); put32(cd, 0); put32(cd, 0);
        put64(cd, dir.size()); put64(cd, dir.size()); put64(cd, cd_size); put64(cd, cd_off);
        put32(cd, 0x07064b50u); put32(cd, 0); put64(cd, cd_off + cd_size); put32(cd, 1);
    }
    put32(cd, 0x06054b50u); put16(cd, 0); put16(cd, 0);
    put16(cd, big ? 0xffff : (uint16_t)dir.size()); put16(cd, big ? 0xffff : (uint16_t)dir.size());
    put32(cd, big ? 0xffffffffu : (uint32_t)cd_size); put32(cd, big ? 0xffffffffu : (uint32_t)cd_off); put16(cd, 0);
    NPZ_REQUIRE(fwrite(cd.data(), 1, cd.size(), fh.f) == cd.size() && fflush(fh.f) == 0, VS_ERR_INVALID, "%s: write failed", path);
    return VS_OK;
}

}  // namespace

extern "C" int vs_npz_write(const char *path, const vs_npz_member_in *members, int n_members, int level, int threads) {
    NPZ_NOTHROW(npz_write_impl(path, members, n_members, level, threads))
}
