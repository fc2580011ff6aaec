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
template <typename Sink>
int stream_member(const vs_npz *z, const NpzMember &m, Sink &&sink) {
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
            if (!sink(in.data(), n)) return VS_OK;
        }
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
        if (produced && !sink(out.data(), produced)) stop = true;
    }
    inflateEnd(&zs);
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
    std::vector<uint8_t> cd((size_t)cd_size);
    NPZ_REQUIRE(fseeko(fh.f, (off_t)cd_off, SEEK_SET) == 0 && fread(cd.data(), 1, (size_t)cd_size, fh.f) == cd_size, VS_ERR_INVALID,
                "%s: cannot read the central directory", path);
    vs_npz *z = new vs_npz;
    z->path = path;
    size_t p = 0;
    for (uint64_t i = 0; i < n_entries; ++i) {
        if (p + 46 > cd.size() || rd32(&cd[p]) != 0x02014b50u) { delete z; NPZ_REQUIRE(false, VS_ERR_INVALID, "%s: corrupt central directory", path); }
        NpzMember m;
        m.method = rd16(&cd[p + 10]);
        m.comp_size = rd32(&cd[p + 20]); m.raw_size = rd32(&cd[p + 24]);
        const uint16_t nl = rd16(&cd[p + 28]), xl = rd16(&cd[p + 30]), cl = rd16(&cd[p + 32]);
        m.local_off = rd32(&cd[p + 42]);
        std::string name((const char *)&cd[p + 46], nl);
        size_t x = p + 46 + nl;
        const size_t xend = x + xl;
        while (x + 4 <= xend) {   // ZIP64 extended information: only the fields that overflowed, in this order
            const uint16_t id = rd16(&cd[x]), sz = rd16(&cd[x + 2]);
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
    *out = z;
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

int npz_read_impl(vs_npz *z, const char *name, void *dst, int dst_dtype, int64_t skip_elems, int64_t n_elems, int64_t add_offset) {
    NPZ_REQUIRE(z && name && (dst || n_elems == 0), VS_ERR_INVALID, "vs_npz_read: NULL argument");
    NpzMember *m = find_member(z, name);
    NPZ_REQUIRE(m, VS_ERR_INVALID, "%s: no member %s", z->path.c_str(), name);
    int rc = parse_npy_header(z, *m);
    if (rc) return rc;
    NPZ_REQUIRE(m->dtype != VS_NONE, VS_ERR_UNSUPPORTED, "%s: member %s has dtype %s (int32/int64/uint16/uint32/float32/float16 only)",
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
    size_t written = 0;
    bool ok = true;
    auto emit = [&](const uint8_t *p, size_t n_el) {
        switch (m->dtype) {
            case VS_I64: ok = convert_run((const int64_t *)p, n_el, dst, dst_dtype, written, add_offset); break;
            case VS_I32: ok = convert_run((const int32_t *)p, n_el, dst, dst_dtype, written, add_offset); break;
            case VS_U32: ok = convert_run((const uint32_t *)p, n_el, dst, dst_dtype, written, add_offset); break;
            case VS_U16: ok = convert_run((const uint16_t *)p, n_el, dst, dst_dtype, written, add_offset); break;
            case VS_F32: ok = convert_run((const float *)p, n_el, dst, dst_dtype, written, 0); break;
            default: ok = convert_run_half((const __half *)p, n_el, dst, dst_dtype, written); break;
        }
        written += n_el;
    };
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
    NPZ_REQUIRE(ok, VS_ERR_INVALID, "%s: conversion of member %s failed", z->path.c_str(), name);
    NPZ_REQUIRE((int64_t)written == n_elems, VS_ERR_INVALID, "%s: member %s ended after %lld of %lld elements", z->path.c_str(), name,
                (long long)written, (long long)n_elems);
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
