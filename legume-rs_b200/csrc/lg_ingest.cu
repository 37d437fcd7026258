// lg_ingest.cu — column-block ingest: the reference's zarr backend as a feed of the path.
//
// data-beans/src/sparse_backend/zarr.rs keeps a feature x cell matrix as 1-D arrays of a Zarr V3 hierarchy written by
// zarrs 0.23 (a FilesystemStore directory):
//
//   zarr.json                      group; attributes nrow / ncol / nnz                      (zarr.rs:515-523)
//   by_column/{indptr,indices}     uint64 arrays, by_column/data float32                     (zarr.rs:31-64)
//   <array>/zarr.json              shape [n], regular chunk grid [c] (c = max(1 MiB / elem, 8192) capped at n,
//                                  utilities/io_helpers.rs:105-115), codecs bytes(little) + zstd(level 5, no checksum)
//                                  (zarr.rs:41, 285-310)
//   <array>/c/<i>                  chunk i (default chunk-key encoding, "/" separator); every chunk holds c elements,
//                                  the last one padded with the fill value; a missing chunk reads as the fill value
//
// This file reads a column range the way SparseIo::read_columns_csc / csc_column_arrays do (zarr.rs:573-587, 982-994): the
// range's indptr slice, then the chunks of indices / data that hold its entries, inflated on the host cores (libzstd is
// bound at run time with dlopen, like NCCL: no link-time dependency) and handed to lg_csc_upload.  A `.zarr.zip` archive of
// such a directory (zarr_io.rs:30-85: zarrs' ZipStorageAdapter; written by common_io.rs:591-640 with STORED entries under the
// prefix `<stem>/`, formerly `<stem>.zarr/`) is read in place through its central directory (ZIP64 included; deflated entries
// through libz, also bound at run time).  The hdf5 twin (hdf5.rs, blosc) is not read.  No reference-written file exists in this environment, so the layout is
// held to the Zarr V3 specification and zarrs' documented defaults: parity unpinned for this row.
#include <dlfcn.h>
#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>  // the z_stream layout only: libz itself is bound with dlopen when a deflated entry is met

#include <atomic>
#include <cerrno>
#include <chrono>
#include <memory>
#include <cmath>
#include <fstream>
#include <mutex>
#include <sstream>
#include <thread>
#include <unordered_map>

#include "lg_common.cuh"

namespace {

// ---- libzstd, bound at run time ----------------------------------------------------------------
struct ZstdApi {
    void* handle = nullptr;
    size_t (*decompress)(void*, size_t, const void*, size_t) = nullptr;
    unsigned (*is_error)(size_t) = nullptr;
    const char* (*error_name)(size_t) = nullptr;
    std::string err;
};
ZstdApi* zstd_api() {
    static ZstdApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* n : {"libzstd.so.1", "libzstd.so"}) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            api.err = "libzstd not found (dlopen libzstd.so.1)";
            return;
        }
        api.decompress = reinterpret_cast<decltype(api.decompress)>(dlsym(api.handle, "ZSTD_decompress"));
        api.is_error = reinterpret_cast<decltype(api.is_error)>(dlsym(api.handle, "ZSTD_isError"));
        api.error_name = reinterpret_cast<decltype(api.error_name)>(dlsym(api.handle, "ZSTD_getErrorName"));
        if (!api.decompress || !api.is_error || !api.error_name) api.err = "libzstd: symbol missing";
    });
    return &api;
}

// ---- the little JSON the metadata needs -----------------------------------------------------------
struct JVal {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;
    const JVal* get(const char* key) const {
        if (kind != Obj) return nullptr;
        for (const auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
};
struct JParser {
    const std::string& s;
    size_t i = 0;
    bool ok = true;
    explicit JParser(const std::string& t) : s(t) {}
    void ws() {
        while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) ++i;
    }
    bool lit(const char* w) {
        const size_t n = strlen(w);
        if (s.compare(i, n, w) == 0) {
            i += n;
            return true;
        }
        return false;
    }
    std::string string_() {
        std::string out;
        ++i;  // opening quote
        while (i < s.size() && s[i] != '"') {
            if (s[i] == '\\' && i + 1 < s.size()) {
                const char c = s[++i];
                if (c == 'n') out += '\n';
                else if (c == 't') out += '\t';
                else if (c == 'u') {  // \uXXXX: the metadata is ASCII; keep a placeholder
                    out += '?';
                    i += 4;
                } else out += c;
                ++i;
            } else {
                out += s[i++];
            }
        }
        if (i >= s.size()) ok = false;
        ++i;
        return out;
    }
    JVal value(int depth = 0) {
        JVal v;
        ws();
        if (i >= s.size() || depth > 64) {
            ok = false;
            return v;
        }
        const char c = s[i];
        if (c == '{') {
            v.kind = JVal::Obj;
            ++i;
            ws();
            if (i < s.size() && s[i] == '}') {
                ++i;
                return v;
            }
            while (ok) {
                ws();
                if (i >= s.size() || s[i] != '"') {
                    ok = false;
                    break;
                }
                std::string k = string_();
                ws();
                if (i >= s.size() || s[i] != ':') {
                    ok = false;
                    break;
                }
                ++i;
                v.obj.emplace_back(std::move(k), value(depth + 1));
                ws();
                if (i < s.size() && s[i] == ',') {
                    ++i;
                    continue;
                }
                if (i < s.size() && s[i] == '}') {
                    ++i;
                    break;
                }
                ok = false;
            }
        } else if (c == '[') {
            v.kind = JVal::Arr;
            ++i;
            ws();
            if (i < s.size() && s[i] == ']') {
                ++i;
                return v;
            }
            while (ok) {
                v.arr.push_back(value(depth + 1));
                ws();
                if (i < s.size() && s[i] == ',') {
                    ++i;
                    continue;
                }
                if (i < s.size() && s[i] == ']') {
                    ++i;
                    break;
                }
                ok = false;
            }
        } else if (c == '"') {
            v.kind = JVal::Str;
            v.str = string_();
        } else if (lit("true")) {
            v.kind = JVal::Bool;
            v.b = true;
        } else if (lit("false")) {
            v.kind = JVal::Bool;
        } else if (lit("null")) {
            v.kind = JVal::Null;
        } else {
            char* end = nullptr;
            errno = 0;
            v.num = strtod(s.c_str() + i, &end);
            if (end == s.c_str() + i) ok = false;
            else i = (size_t)(end - s.c_str());
            v.kind = JVal::Num;
        }
        return v;
    }
};

bool read_file(const std::string& path, std::string* out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    *out = ss.str();
    return true;
}

// ---- libz, bound at run time (deflated zip entries only) ----------------------------------------
struct ZlibApi {
    void* handle = nullptr;
    int (*inflate_init2)(z_streamp, int, const char*, int) = nullptr;
    int (*inflate_fn)(z_streamp, int) = nullptr;
    int (*inflate_end)(z_streamp) = nullptr;
    std::string err;
};
ZlibApi* zlib_api() {
    static ZlibApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* n : {"libz.so.1", "libz.so"}) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            api.err = "libz not found (dlopen libz.so.1): a deflated zip entry cannot be read";
            return;
        }
        api.inflate_init2 = reinterpret_cast<decltype(api.inflate_init2)>(dlsym(api.handle, "inflateInit2_"));
        api.inflate_fn = reinterpret_cast<decltype(api.inflate_fn)>(dlsym(api.handle, "inflate"));
        api.inflate_end = reinterpret_cast<decltype(api.inflate_end)>(dlsym(api.handle, "inflateEnd"));
        if (!api.inflate_init2 || !api.inflate_fn || !api.inflate_end) api.err = "libz: symbol missing";
    });
    return &api;
}

// ---- a store: a directory, or a zip archive of one ------------------------------------------------
struct ZipEntry {
    uint64_t off = 0, csize = 0, usize = 0;  // local header offset, stored and inflated sizes
    uint16_t method = 0;                     // 0 stored, 8 deflated
};
struct ZStore {
    std::string root;  // the directory, or the archive
    bool zip = false;
    int fd = -1;
    std::string prefix;  // the archive's entries sit under `<stem>/` (or `<stem>.zarr/`, or nothing): zarr_io.rs:30-50
    std::unordered_map<std::string, ZipEntry> entries;
    ~ZStore() {
        if (fd >= 0) close(fd);
    }
};
inline uint16_t le16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t le32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint64_t le64(const unsigned char* p) { return (uint64_t)le32(p) | ((uint64_t)le32(p + 4) << 32); }
bool pread_all(int fd, void* dst, size_t n, uint64_t off) {
    char* d = static_cast<char*>(dst);
    while (n) {
        const ssize_t g = pread(fd, d, n, (off_t)off);
        if (g <= 0) return false;
        d += g;
        off += (uint64_t)g;
        n -= (size_t)g;
    }
    return true;
}
// the central directory of the archive -> entries (names as stored)
bool zip_open(const std::string& path, ZStore* st, std::string* err) {
    st->fd = open(path.c_str(), O_RDONLY | O_CLOEXEC);
    if (st->fd < 0) {
        *err = "cannot open " + path;
        return false;
    }
    const off_t fsize = lseek(st->fd, 0, SEEK_END);
    if (fsize < 22) {
        *err = path + ": not a zip archive";
        return false;
    }
    const size_t tail = (size_t)std::min<off_t>(fsize, 65536 + 22 + 20);
    std::vector<unsigned char> buf(tail);
    if (!pread_all(st->fd, buf.data(), tail, (uint64_t)(fsize - (off_t)tail))) {
        *err = path + ": read error";
        return false;
    }
    ssize_t e = -1;
    for (ssize_t i = (ssize_t)tail - 22; i >= 0; --i)
        if (le32(&buf[i]) == 0x06054b50u) {
            e = i;
            break;
        }
    if (e < 0) {
        *err = path + ": no end-of-central-directory record (not a zip archive)";
        return false;
    }
    uint64_t nent = le16(&buf[e + 10]), cd_size = le32(&buf[e + 12]), cd_off = le32(&buf[e + 16]);
    if (nent == 0xffffu || cd_size == 0xffffffffu || cd_off == 0xffffffffu) {  // ZIP64: locator just before the record
        if (e < 20 || le32(&buf[e - 20]) != 0x07064b50u) {
            *err = path + ": ZIP64 locator missing";
            return false;
        }
        const uint64_t r64 = le64(&buf[e - 20 + 8]);
        unsigned char rec[56];
        if (!pread_all(st->fd, rec, sizeof rec, r64) || le32(rec) != 0x06064b50u) {
            *err = path + ": ZIP64 end-of-central-directory record unreadable";
            return false;
        }
        nent = le64(rec + 32);
        cd_size = le64(rec + 40);
        cd_off = le64(rec + 48);
    }
    if (cd_off + cd_size > (uint64_t)fsize) {
        *err = path + ": central directory outside the file";
        return false;
    }
    std::vector<unsigned char> cd((size_t)cd_size);
    if (cd_size && !pread_all(st->fd, cd.data(), (size_t)cd_size, cd_off)) {
        *err = path + ": central directory unreadable";
        return false;
    }
    size_t p = 0;
    for (uint64_t i = 0; i < nent; ++i) {
        if (p + 46 > cd.size() || le32(&cd[p]) != 0x02014b50u) {
            *err = path + ": central directory entry " + std::to_string(i) + " is malformed";
            return false;
        }
        ZipEntry ze;
        ze.method = le16(&cd[p + 10]);
        ze.csize = le32(&cd[p + 20]);
        ze.usize = le32(&cd[p + 24]);
        const size_t nlen = le16(&cd[p + 28]), xlen = le16(&cd[p + 30]), clen = le16(&cd[p + 32]);
        ze.off = le32(&cd[p + 42]);
        if (p + 46 + nlen + xlen + clen > cd.size()) {
            *err = path + ": central directory entry overruns the directory";
            return false;
        }
        const std::string name(reinterpret_cast<const char*>(&cd[p + 46]), nlen);
        // ZIP64 extended information: the 64-bit values of the fields that read 0xffffffff, in this order
        size_t x = p + 46 + nlen;
        const size_t xend = x + xlen;
        while (x + 4 <= xend) {
            const uint16_t id = le16(&cd[x]), sz = le16(&cd[x + 2]);
            if (id == 0x0001) {
                size_t q = x + 4;
                if (ze.usize == 0xffffffffu && q + 8 <= xend) {
                    ze.usize = le64(&cd[q]);
                    q += 8;
                }
                if (ze.csize == 0xffffffffu && q + 8 <= xend) {
                    ze.csize = le64(&cd[q]);
                    q += 8;
                }
                if (ze.off == 0xffffffffu && q + 8 <= xend) ze.off = le64(&cd[q]);
            }
            x += 4 + sz;
        }
        if (!name.empty() && name.back() != '/') st->entries.emplace(name, ze);
        p += 46 + nlen + xlen + clen;
    }
    // where the hierarchy starts inside the archive (detect_zip_zarr_prefix, zarr_io.rs:30-50)
    std::string file = path.substr(path.find_last_of('/') == std::string::npos ? 0 : path.find_last_of('/') + 1);
    auto strip = [](std::string v, const std::string& suf) {
        if (v.size() >= suf.size() && v.compare(v.size() - suf.size(), suf.size(), suf) == 0) v.resize(v.size() - suf.size());
        return v;
    };
    const std::string stem = strip(file, ".zip");
    const std::string new_prefix = strip(stem, ".zarr") + "/", legacy_prefix = stem + "/";
    auto any_under = [&](const std::string& pre) {
        for (const auto& kv : st->entries)
            if (kv.first.compare(0, pre.size(), pre) == 0) return true;
        return false;
    };
    st->prefix = any_under(new_prefix) ? new_prefix : (new_prefix != legacy_prefix && any_under(legacy_prefix) ? legacy_prefix : "");
    st->zip = true;
    return true;
}
// the bytes under `key` (a path relative to the root of the hierarchy): 1 read, 0 absent, -1 error (*err set)
int store_read(const ZStore& st, const std::string& key, std::string* out, std::string* err) {
    if (!st.zip) return read_file(st.root + "/" + key, out) ? 1 : 0;
    const auto it = st.entries.find(st.prefix + key);
    if (it == st.entries.end()) return 0;
    const ZipEntry& ze = it->second;
    unsigned char lh[30];
    if (!pread_all(st.fd, lh, sizeof lh, ze.off) || le32(lh) != 0x04034b50u) {
        *err = st.root + ": local header of " + key + " unreadable";
        return -1;
    }
    const uint64_t data_off = ze.off + 30 + le16(lh + 26) + le16(lh + 28);
    std::string comp((size_t)ze.csize, '\0');
    if (ze.csize && !pread_all(st.fd, &comp[0], (size_t)ze.csize, data_off)) {
        *err = st.root + ": " + key + " is truncated";
        return -1;
    }
    if (ze.method == 0) {
        out->swap(comp);
        return 1;
    }
    if (ze.method != 8) {
        *err = st.root + ": " + key + " uses zip compression method " + std::to_string(ze.method) + " (stored and deflated entries are read)";
        return -1;
    }
    ZlibApi* z = zlib_api();
    if (!z->err.empty()) {
        *err = z->err;
        return -1;
    }
    out->assign((size_t)ze.usize, '\0');
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (z->inflate_init2(&zs, -15, ZLIB_VERSION, (int)sizeof(z_stream)) != Z_OK) {  // raw deflate stream
        *err = "libz: inflateInit2 failed";
        return -1;
    }
    zs.next_in = reinterpret_cast<Bytef*>(&comp[0]);
    zs.avail_in = (uInt)comp.size();
    zs.next_out = reinterpret_cast<Bytef*>(&(*out)[0]);
    zs.avail_out = (uInt)out->size();
    const int rc = z->inflate_fn(&zs, Z_FINISH);
    const bool ok = rc == Z_STREAM_END && zs.total_out == ze.usize;
    z->inflate_end(&zs);
    if (!ok) {
        *err = st.root + ": " + key + " does not inflate to its recorded size";
        return -1;
    }
    return 1;
}

struct ZArray {
    const ZStore* store = nullptr;
    std::string dir;       // key of the array inside the store: by_column/<name>
    uint64_t n = 0, chunk = 0;
    int elem = 0;          // bytes per element
    bool is_float = false;
    bool zstd = false;
    std::string sep = "/";
    double fill = 0.0;
};

}  // namespace

struct lg_zarr {
    ZStore store;
    std::string root, err;
    uint64_t nrows = 0, ncols = 0, nnz = 0;
    ZArray indptr, indices, data;
    // host arrays of lg_zarr_read_columns, kept between calls: a visitor reads block after block of the same size, and
    // mapping + first-touching + unmapping 12 bytes per non-zero every time cost a fifth of the call
    std::unique_ptr<uint64_t[]> buf_ip, buf_ix;
    std::unique_ptr<float[]> buf_v;
    size_t cap_ip = 0, cap_e = 0;
};

namespace {

int zfail(lg_zarr* z, int code, const std::string& msg) {
    if (z) z->err = msg;
    return code;
}

// zarr.json of one array -> ZArray; `want_float` / `want_bytes`: what the reference writes there
bool open_array(const ZStore& st, const std::string& dir, bool want_float, int want_bytes, ZArray* a, std::string* err) {
    std::string txt;
    if (store_read(st, dir + "/zarr.json", &txt, err) != 1) {
        if (err->empty()) *err = "cannot read " + dir + "/zarr.json in " + st.root;
        return false;
    }
    JParser p(txt);
    const JVal m = p.value();
    if (!p.ok || m.kind != JVal::Obj) {
        *err = dir + "/zarr.json: not JSON";
        return false;
    }
    const JVal* fmt = m.get("zarr_format");
    const JVal* node = m.get("node_type");
    if (!fmt || fmt->kind != JVal::Num || (int)fmt->num != 3 || !node || node->str != "array") {
        *err = dir + ": not a Zarr V3 array (zarr_format 3, node_type array)";
        return false;
    }
    const JVal* shape = m.get("shape");
    if (!shape || shape->kind != JVal::Arr || shape->arr.size() != 1 || shape->arr[0].kind != JVal::Num) {
        *err = dir + ": the backend's arrays are 1-D";
        return false;
    }
    a->n = (uint64_t)shape->arr[0].num;
    const JVal* dt = m.get("data_type");
    const std::string dts = dt && dt->kind == JVal::Str ? dt->str : "";
    if (want_float ? dts != "float32" : dts != "uint64") {
        *err = dir + ": data_type " + dts + ", expected " + (want_float ? "float32" : "uint64");
        return false;
    }
    a->elem = want_bytes;
    a->is_float = want_float;
    const JVal* grid = m.get("chunk_grid");
    const JVal* gconf = grid ? grid->get("configuration") : nullptr;
    const JVal* cs = gconf ? gconf->get("chunk_shape") : nullptr;
    const JVal* gname = grid ? grid->get("name") : nullptr;
    if (!gname || gname->str != "regular" || !cs || cs->kind != JVal::Arr || cs->arr.size() != 1 || cs->arr[0].num < 1.0) {
        *err = dir + ": chunk_grid must be regular and 1-D";
        return false;
    }
    a->chunk = (uint64_t)cs->arr[0].num;
    if (const JVal* ke = m.get("chunk_key_encoding")) {
        const JVal* kn = ke->get("name");
        if (kn && kn->str != "default") {
            *err = dir + ": chunk_key_encoding " + kn->str + " (only `default` is read)";
            return false;
        }
        const JVal* kc = ke->get("configuration");
        const JVal* sp = kc ? kc->get("separator") : nullptr;
        if (sp && sp->kind == JVal::Str) a->sep = sp->str;
    }
    if (const JVal* fv = m.get("fill_value")) {
        if (fv->kind == JVal::Num) a->fill = fv->num;
        else if (fv->kind == JVal::Str) a->fill = (fv->str == "NaN") ? NAN : (fv->str == "Infinity" ? INFINITY : (fv->str == "-Infinity" ? -INFINITY : 0.0));
    }
    const JVal* codecs = m.get("codecs");
    if (!codecs || codecs->kind != JVal::Arr) {
        *err = dir + ": no codecs";
        return false;
    }
    bool have_bytes = false;
    for (const JVal& c : codecs->arr) {
        const JVal* nm = c.get("name");
        const std::string name = nm ? nm->str : "";
        if (name == "bytes") {
            have_bytes = true;
            const JVal* cc = c.get("configuration");
            const JVal* en = cc ? cc->get("endian") : nullptr;
            if (en && en->str != "little") {
                *err = dir + ": big-endian bytes codec";
                return false;
            }
        } else if (name == "zstd") {
            a->zstd = true;
        } else {
            *err = dir + ": codec `" + name + "` is not read (the backend writes bytes + zstd)";
            return false;
        }
    }
    if (!have_bytes) {
        *err = dir + ": the array-to-bytes codec must be `bytes`";
        return false;
    }
    a->dir = dir;
    a->store = &st;
    return true;
}

// elements [e0, e1) of a 1-D array into dst (a->elem bytes each): chunk by chunk, inflated on `nthreads` threads
bool read_range(const ZArray& a, uint64_t e0, uint64_t e1, void* dst, int nthreads, std::string* err) {
    if (e1 > a.n || e0 > e1) {
        *err = a.dir + ": element range outside the array";
        return false;
    }
    if (e0 == e1) return true;
    const uint64_t c0 = e0 / a.chunk, c1 = (e1 - 1) / a.chunk + 1;
    ZstdApi* z = a.zstd ? zstd_api() : nullptr;
    if (z && !z->err.empty()) {
        *err = z->err;
        return false;
    }
    std::atomic<uint64_t> next{c0};
    std::mutex mu;
    std::string first_err;
    auto work = [&]() {
        std::vector<char> raw((size_t)a.chunk * a.elem);
        std::string comp;
        for (;;) {
            const uint64_t c = next.fetch_add(1);
            if (c >= c1) break;
            const std::string path = a.dir + "/c" + a.sep + std::to_string(c);
            const uint64_t lo = std::max<uint64_t>(e0, c * a.chunk), hi = std::min<uint64_t>(e1, (c + 1) * a.chunk);
            char* out = static_cast<char*>(dst) + (size_t)(lo - e0) * a.elem;
            std::string rerr;
            const int got_chunk = store_read(*a.store, path, &comp, &rerr);
            if (got_chunk < 0) {
                std::lock_guard<std::mutex> g(mu);
                if (first_err.empty()) first_err = rerr;
                continue;
            }
            if (got_chunk == 0) {  // an absent chunk is the fill value (Zarr V3)
                if (a.is_float) {
                    const float f = (float)a.fill;
                    for (uint64_t i = 0; i < hi - lo; ++i) memcpy(out + i * 4, &f, 4);
                } else {
                    const uint64_t u = (uint64_t)a.fill;
                    for (uint64_t i = 0; i < hi - lo; ++i) memcpy(out + i * 8, &u, 8);
                }
                continue;
            }
            const char* src = comp.data();
            if (a.zstd) {
                const size_t got = z->decompress(raw.data(), raw.size(), comp.data(), comp.size());
                if (z->is_error(got) || got != raw.size()) {
                    std::lock_guard<std::mutex> g(mu);
                    if (first_err.empty())
                        first_err = path + ": " + (z->is_error(got) ? std::string(z->error_name(got)) : "chunk inflates to " + std::to_string(got) + " bytes, expected " + std::to_string(raw.size()));
                    continue;
                }
                src = raw.data();
            } else if (comp.size() != raw.size()) {
                std::lock_guard<std::mutex> g(mu);
                if (first_err.empty()) first_err = path + ": chunk of " + std::to_string(comp.size()) + " bytes, expected " + std::to_string(raw.size());
                continue;
            }
            memcpy(out, src + (size_t)(lo - c * a.chunk) * a.elem, (size_t)(hi - lo) * a.elem);
        }
    };
    const uint64_t nch = c1 - c0;
    int nt = nthreads < 1 ? 1 : nthreads;
    if ((uint64_t)nt > nch) nt = (int)nch;
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    if (!first_err.empty()) {
        *err = first_err;
        return false;
    }
    return true;
}

int ingest_threads() {
    if (const char* e = getenv("LG_INGEST_THREADS")) return std::max(1, atoi(e));
    const unsigned hc = std::thread::hardware_concurrency();
    return (int)(hc ? (hc > 32 ? 32 : hc) : 4);
}

}  // namespace

extern "C" int lg_zarr_open(const char* path, lg_zarr** out, char* err, size_t err_len) {
    auto say = [&](const std::string& m) {
        if (err && err_len) snprintf(err, err_len, "%s", m.c_str());
        return LG_ERR_INVALID;
    };
    if (!path || !out) return say("lg_zarr_open: null argument");
    *out = nullptr;
    std::string root(path);
    while (root.size() > 1 && root.back() == '/') root.pop_back();
    std::unique_ptr<lg_zarr> zz(new lg_zarr());
    zz->store.root = root;
    const bool is_zip = root.size() > 4 && root.compare(root.size() - 4, 4, ".zip") == 0;  // zarr_io.rs:59: by extension
    std::string serr;
    if (is_zip && !zip_open(root, &zz->store, &serr)) return say("lg_zarr_open: " + serr);
    std::string txt;
    if (store_read(zz->store, "zarr.json", &txt, &serr) != 1)
        return say(serr.empty() ? "lg_zarr_open: cannot read " + root + "/zarr.json (a Zarr V3 directory store or a .zarr.zip of one is expected)"
                                : "lg_zarr_open: " + serr);
    JParser p(txt);
    const JVal g = p.value();
    const JVal* node = g.get("node_type");
    if (!p.ok || !node || node->str != "group") return say("lg_zarr_open: " + root + "/zarr.json is not a Zarr V3 group");
    lg_zarr* z = zz.release();
    z->root = root;
    const JVal* at = g.get("attributes");
    auto attr = [&](const char* k, uint64_t* v) {
        const JVal* a = at ? at->get(k) : nullptr;
        if (!a || a->kind != JVal::Num || a->num < 0) return false;
        *v = (uint64_t)a->num;
        return true;
    };
    std::string e;
    const bool have_shape = attr("nrow", &z->nrows) && attr("ncol", &z->ncols);
    if (!open_array(z->store, "by_column/indptr", false, 8, &z->indptr, &e) || !open_array(z->store, "by_column/indices", false, 8, &z->indices, &e) ||
        !open_array(z->store, "by_column/data", true, 4, &z->data, &e)) {
        delete z;
        return say("lg_zarr_open: " + e);
    }
    if (!have_shape) {
        delete z;
        return say("lg_zarr_open: the root group carries no nrow / ncol attributes");
    }
    if (!attr("nnz", &z->nnz)) z->nnz = z->data.n;
    if (z->indptr.n != z->ncols + 1 || z->indices.n != z->data.n || z->nnz != z->data.n) {
        delete z;
        return say("lg_zarr_open: array lengths do not fit the attributes (indptr " + std::to_string(z->indptr.n) + ", indices " +
                   std::to_string(z->indices.n) + ", data " + std::to_string(z->data.n) + "; ncol " + std::to_string(z->ncols) + ", nnz " +
                   std::to_string(z->nnz) + ")");
    }
    *out = z;
    return LG_OK;
}

extern "C" void lg_zarr_close(lg_zarr* z) { delete z; }
extern "C" const char* lg_zarr_last_error(const lg_zarr* z) { return z ? z->err.c_str() : "null handle"; }
extern "C" int lg_zarr_shape(const lg_zarr* z, uint64_t* nrows, uint64_t* ncols, uint64_t* nnz) {
    if (!z) return LG_ERR_INVALID;
    if (nrows) *nrows = z->nrows;
    if (ncols) *ncols = z->ncols;
    if (nnz) *nnz = z->nnz;
    return LG_OK;
}

// entries of columns [col_lo, col_hi): [*first, *last)
extern "C" int lg_zarr_column_extent(lg_zarr* z, uint64_t col_lo, uint64_t col_hi, uint64_t* first, uint64_t* last) {
    if (!z || !first || !last) return LG_ERR_INVALID;
    if (col_lo > col_hi || col_hi > z->ncols) return zfail(z, LG_ERR_INVALID, "lg_zarr_column_extent: column range outside the matrix");
    uint64_t a = 0, b = 0;
    std::string e;
    if (!read_range(z->indptr, col_lo, col_lo + 1, &a, 1, &e) || !read_range(z->indptr, col_hi, col_hi + 1, &b, 1, &e))
        return zfail(z, LG_ERR_INVALID, e);
    if (b < a || b > z->nnz) return zfail(z, LG_ERR_INVALID, "lg_zarr_column_extent: indptr not monotone");
    *first = a;
    *last = b;
    return LG_OK;
}

// csc_column_arrays of the range: indptr (col_hi - col_lo + 1 entries, rebased to 0), indices and data (*last - *first
// entries each, sized by the caller from lg_zarr_column_extent)
extern "C" int lg_zarr_read_columns_host(lg_zarr* z, uint64_t col_lo, uint64_t col_hi, uint64_t* indptr, uint64_t* indices, float* data) {
    if (!z || !indptr) return LG_ERR_INVALID;
    if (col_lo > col_hi || col_hi > z->ncols) return zfail(z, LG_ERR_INVALID, "lg_zarr_read_columns: column range outside the matrix");
    const int nt = ingest_threads();
    std::string e;
    if (!read_range(z->indptr, col_lo, col_hi + 1, indptr, nt, &e)) return zfail(z, LG_ERR_INVALID, e);
    const uint64_t base = indptr[0], end = indptr[col_hi - col_lo];
    if (end < base || end > z->nnz) return zfail(z, LG_ERR_INVALID, "lg_zarr_read_columns: indptr not monotone");
    for (uint64_t j = 0; j <= col_hi - col_lo; ++j) {
        if (indptr[j] < base || indptr[j] > end || (j && indptr[j] < indptr[j - 1])) return zfail(z, LG_ERR_INVALID, "lg_zarr_read_columns: indptr not monotone");
        indptr[j] -= base;
    }
    if (end > base) {
        if (!indices || !data) return zfail(z, LG_ERR_INVALID, "lg_zarr_read_columns: null indices / data");
        if (!read_range(z->indices, base, end, indices, nt, &e) || !read_range(z->data, base, end, data, nt, &e)) return zfail(z, LG_ERR_INVALID, e);
    }
    return LG_OK;
}

// the range as a device-resident block: read_columns_csc of the backend followed by the feed of the path (lg_csc_upload:
// one-byte row gaps, byte counts, pinned ring)
extern "C" int lg_zarr_read_columns(lg_ctx* ctx, lg_zarr* z, uint64_t col_lo, uint64_t col_hi, lg_csc** out) {
    if (!ctx || !z || !out) return LG_ERR_INVALID;
    *out = nullptr;
    uint64_t first = 0, last = 0;
    int rc = lg_zarr_column_extent(z, col_lo, col_hi, &first, &last);
    if (rc != LG_OK) return lg_fail(ctx, rc, z->err);
    const auto t0 = std::chrono::steady_clock::now();
    // uninitialised host arrays (a std::vector would first write 12 bytes per entry of zeros), grown when a block needs more
    const size_t ne = (size_t)(last - first), nc = (size_t)(col_hi - col_lo + 1);
    if (nc > z->cap_ip) {
        z->buf_ip.reset(new uint64_t[nc]);
        z->cap_ip = nc;
    }
    if (ne > z->cap_e || !z->buf_ix) {
        z->buf_ix.reset();
        z->buf_v.reset();
        z->buf_ix.reset(new uint64_t[ne ? ne : 1]);
        z->buf_v.reset(new float[ne ? ne : 1]);
        z->cap_e = ne ? ne : 1;
    }
    rc = lg_zarr_read_columns_host(z, col_lo, col_hi, z->buf_ip.get(), z->buf_ix.get(), z->buf_v.get());
    if (rc != LG_OK) return lg_fail(ctx, rc, z->err);
    const auto t1 = std::chrono::steady_clock::now();
    rc = lg_csc_upload(ctx, z->buf_ip.get(), z->buf_ix.get(), z->buf_v.get(), z->nrows, 0, col_hi - col_lo, nullptr, out);
    if (getenv("LG_INGEST_TRACE")) {
        const auto t2 = std::chrono::steady_clock::now();
        const double dec = std::chrono::duration<double, std::milli>(t1 - t0).count(), up = std::chrono::duration<double, std::milli>(t2 - t1).count();
        fprintf(stderr, "[lg_zarr_read_columns] %zu columns, %zu entries: inflate %.1f ms (%.2f GB/s of arrays, %d threads), upload %.1f ms\n", nc - 1, ne,
                dec, (12.0 * ne + 8.0 * nc) / dec * 1e-6, ingest_threads(), up);
    }
    return rc;
}
