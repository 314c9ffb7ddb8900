// .rts scene ingest: replaces getnum / read / gettexnum / getppmnum / getppmpaths of the
// reference (raygpu/kernel.cu:1113-1169, 1186-1530, 1172-1183, 1979-2018).
//
// Format: text, one record per line, comma separated.  First character
// '/' = comment, '*' = settings line, anything else = object line of up to 38 columns.
// The reference parses every number with std::stof / std::stoi (prefix parse, so
// "45.000000" -> 45 for the integer fields); this loader produces bit-identical floats
// (std::from_chars is correctly rounded, like strtof) but parses the file in one pass over
// a memory map, in parallel over line ranges, instead of two stringstream passes.
//
// Deliberate differences from the reference, all on inputs it has undefined behaviour for:
//   * blank lines are skipped (the reference throws from stof("")),
//   * the token "r" ("random number", kernel.cu:1098-1102, 1308) becomes a deterministic hash
//     of (line, column) in [0,1) instead of rand() seeded from the tick count,
//   * fields a short line does not name are zero / the struct defaults (kernel.cu:48-74)
//     instead of indeterminate heap contents,
//   * there is no phantom object at index `lines` (kernel.cu:1158, 1518).
#include "drb_internal.h"

#include <algorithm>
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <mutex>
#include <functional>
#include <thread>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

thread_local std::string g_last_error;

struct Token { const char* p; size_t n; };

// std::stof semantics: skip leading white space, parse the longest valid prefix, fail if none.
bool parse_float(Token t, float& out)
{
    const char* p = t.p; const char* e = t.p + t.n;
    while (p < e && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\v' || *p == '\f')) ++p;
    if (p < e && *p != '+') {
        auto r = std::from_chars(p, e, out);
        if (r.ec == std::errc()) return true;
        if (r.ec == std::errc::result_out_of_range) {       // stof throws; saturate like strtof
            out = (*p == '-') ? -HUGE_VALF : HUGE_VALF;
            return true;
        }
    }
    // rare spellings from_chars does not take ("+1", hex floats): strtof on a terminated copy
    char buf[128];
    size_t n = std::min((size_t)(e - p), sizeof buf - 1);
    memcpy(buf, p, n); buf[n] = 0;
    char* endp = nullptr;
    float v = strtof(buf, &endp);
    if (endp == buf) return false;
    out = v;
    return true;
}

// std::stoi semantics: base 10 prefix
bool parse_int(Token t, int32_t& out)
{
    const char* p = t.p; const char* e = t.p + t.n;
    while (p < e && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\v' || *p == '\f')) ++p;
    bool neg = false;
    if (p < e && (*p == '+' || *p == '-')) { neg = *p == '-'; ++p; }
    if (p >= e || *p < '0' || *p > '9') return false;
    long long v = 0;
    while (p < e && *p >= '0' && *p <= '9') { v = v * 10 + (*p - '0'); if (v > 4294967296LL) v = 4294967296LL; ++p; }
    if (neg) v = -v;
    if (v > INT32_MAX) v = INT32_MAX;
    if (v < INT32_MIN) v = INT32_MIN;
    out = (int32_t)v;
    return true;
}

inline bool token_is(Token t, const char* s) { size_t n = strlen(s); return t.n == n && memcmp(t.p, s, n) == 0; }

float hash_unit(uint64_t line, uint32_t col)
{
    uint64_t z = line * 0x9E3779B97F4A7C15ull + col * 0xBF58476D1CE4E5B9ull + 0x94D049BB133111EBull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

void object_defaults(drb_object& o)
{
    memset(&o, 0, sizeof o);
    // kernel.cu:55-64, 70-71
    o.norm[0] = -2; o.norm[1] = -3; o.norm[2] = -20;
    for (float* n : { o.n1, o.n2, o.n3 }) { n[0] = -2; n[1] = -3; n[2] = -20; }
    o.t1[0] = 0; o.t1[1] = 1; o.t2[0] = 0; o.t2[1] = 0; o.t3[0] = 1; o.t3[1] = 0;
    o.texnum = -1; o.rtexnum = -1;
}

struct ParseError { std::string msg; };

// one object line -> drb_object; column map kernel.cu:1316-1503
void parse_object_line(const char* p, const char* e, uint64_t lineno, const std::vector<std::string>& tex, drb_object& o)
{
    object_defaults(o);
    int col = 0;
    const char* q = p;
    for (;;) {
        const char* c = q;                              // tokens are ~9 characters: a byte loop beats a memchr call
        while (c < e && *c != ',') ++c;
        if (c == e) c = nullptr;
        Token t{ q, (size_t)((c ? c : e) - q) };
        float f = 0; int32_t iv = 0;
        const bool is_r = t.n == 1 && *t.p == 'r';
        auto F = [&]() -> float {
            if (is_r) return hash_unit(lineno, (uint32_t)col);
            if (!parse_float(t, f)) throw ParseError{ "line " + std::to_string(lineno) + " column " + std::to_string(col) + ": not a number: '" + std::string(t.p, t.n) + "'" };
            return f;
        };
        auto I = [&]() -> int32_t {
            if (is_r) return 0;                         // stoi("0.xxxxx") of the reference's random string
            if (!parse_int(t, iv)) throw ParseError{ "line " + std::to_string(lineno) + " column " + std::to_string(col) + ": not an integer: '" + std::string(t.p, t.n) + "'" };
            return iv;
        };
        switch (col) {
        case 0: o.pos[0] = F(); break;   case 1: o.pos[1] = F(); break;   case 2: o.pos[2] = F(); break;
        case 3: o.type = I(); break;
        case 4: o.col[0] = F(); break;   case 5: o.col[1] = F(); break;   case 6: o.col[2] = F(); break;
        case 7: o.add_y = F(); break;    case 8: o.add_x = F(); break;
        case 9: o.dim[0] = F(); break;   case 10: o.dim[1] = F(); break;  case 11: o.dim[2] = F(); break;
        case 12: o.mat = I(); break;
        case 13: o.rot[0] = F(); break;  case 14: o.rot[1] = F(); break;  case 15: o.rot[2] = F(); break;
        case 16: o.norm[0] = F(); break; case 17: o.norm[1] = F(); break; case 18: o.norm[2] = F(); break;
        case 19: o.n1[0] = F(); break;   case 20: o.n1[1] = F(); break;   case 21: o.n1[2] = F(); break;
        case 22: o.n2[0] = F(); break;   case 23: o.n2[1] = F(); break;   case 24: o.n2[2] = F(); break;
        case 25: o.n3[0] = F(); break;   case 26: o.n3[1] = F(); break;   case 27: o.n3[2] = F(); break;
        case 28: o.t1[0] = F(); break;   case 29: o.t1[1] = F(); break;
        case 30: o.t2[0] = F(); break;   case 31: o.t2[1] = F(); break;
        case 32: o.t3[0] = F(); break;   case 33: o.t3[1] = F(); break;
        case 34: o.smooth = (I() == 1) ? 1 : 0; break;
        case 35: o.checker = (I() == 1) ? 1 : 0; break;
        case 36: if (!token_is(t, "no")) o.texnum = drb_find_texture(tex, std::string(t.p, t.n)); break;
        case 37: if (!token_is(t, "no")) o.rtexnum = drb_find_texture(tex, std::string(t.p, t.n)); break;
        default: break;                                 // the reference ignores further columns
        }
        ++col;
        if (!c) break;
        q = c + 1;
    }
    o.ncols = col;
}

// settings line, column map kernel.cu:1230-1293
void parse_settings_line(const char* p, const char* e, uint64_t lineno, const std::vector<std::string>& tex, drb_settings& s)
{
    int col = 0;
    const char* q = p;
    for (;;) {
        const char* c = (const char*)memchr(q, ',', (size_t)(e - q));
        Token t{ q, (size_t)((c ? c : e) - q) };
        float f = 0; int32_t iv = 0;
        auto F = [&]() -> float {
            if (!parse_float(t, f)) throw ParseError{ "settings line " + std::to_string(lineno) + " column " + std::to_string(col) + ": not a number: '" + std::string(t.p, t.n) + "'" };
            return f;
        };
        auto I = [&]() -> int32_t {
            if (!parse_int(t, iv)) throw ParseError{ "settings line " + std::to_string(lineno) + " column " + std::to_string(col) + ": not an integer: '" + std::string(t.p, t.n) + "'" };
            return iv;
        };
        switch (col) {
        case 1: s.cam[0] = F(); break;  case 2: s.cam[1] = F(); break;  case 3: s.cam[2] = F(); break;
        case 4: s.aperture = F(); break;
        case 5: s.look[0] = F(); break; case 6: s.look[1] = F(); break; case 7: s.look[2] = F(); break;
        case 8: s.focus = F(); break;
        case 9: s.fov = I(); break;
        case 10: s.max_depth = I(); break;
        case 11: s.spp = I(); break;
        case 12: s.bg_intensity = F(); break;
        case 13: if (!token_is(t, "no")) s.backtex = drb_find_texture(tex, std::string(t.p, t.n)); break;
        case 14: s.width = I(); break;
        case 15: s.height = I(); break;
        default: break;
        }
        ++col;
        if (!c) break;
        q = c + 1;
    }
}

// Whole-file parse, every phase parallel over byte ranges that end on line boundaries:
//   A  each thread splits its range into lines (std::getline semantics: a final line without '\n' counts, an empty
//      tail does not) and classifies them: blank (skipped), '/' comment, '*' settings, object
//   B  (serial, a few microseconds) prefix sums give every range its first line number and first object index;
//      settings lines are applied in file order
//   C  each thread parses the object lines of its own range straight into the (uninitialised) object array and
//      checks that the tracer can represent them
// Results do not depend on the thread count: line numbers, object indices, the first warning and the first error
// are the ones a sequential pass would report.
int parse_buffer(const char* text, size_t len, const char* tex_dir, drb_host_scene** out)
{
    auto hs = new drb_host_scene();
    drb_settings_default(&hs->settings);
    hs->tex_paths = drb_scan_textures(tex_dir);

    struct Line { const char* b; const char* e; };
    struct Range {
        const char* b; const char* e;
        std::vector<Line> obj;                // object lines
        std::vector<int64_t> obj_local;       // their line index inside the range
        std::vector<std::pair<int64_t, Line>> settings;
        int64_t nlines = 0, blank = 0, first_blank = -1;
        int64_t line_base = 0, obj_base = 0;
        // phase C
        int64_t bad = 0, first_bad_line = -1; int first_bad_type = 0, first_bad_cols = 0;
        int64_t err_line = -1; std::string err;
    };
    unsigned hw = std::thread::hardware_concurrency();
    const int nthreads = (int)std::min<size_t>(hw ? hw : 4, std::max<size_t>(1, len >> 20));
    std::vector<Range> rg((size_t)nthreads);
    const char* end = text + len;
    {
        const char* b = text;
        for (int t = 0; t < nthreads; ++t) {
            const char* e = end;
            if (t + 1 < nthreads) {
                const char* guess = text + len * (size_t)(t + 1) / (size_t)nthreads;
                if (guess < b) guess = b;
                const char* nl = (const char*)memchr(guess, '\n', (size_t)(end - guess));
                e = nl ? nl + 1 : end;
            }
            rg[(size_t)t].b = b; rg[(size_t)t].e = e;
            b = e;
        }
    }
    auto run = [&](auto&& fn) {
        if (nthreads <= 1) { fn(0); return; }
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t) pool.emplace_back(fn, t);
        for (auto& th : pool) th.join();
    };

    // ---- A: lines ----
    run([&](int t) {
        Range& r = rg[(size_t)t];
        r.obj.reserve((size_t)(r.e - r.b) / 200 + 16);
        r.obj_local.reserve((size_t)(r.e - r.b) / 200 + 16);
        const char* p = r.b;
        while (p < r.e) {
            const char* nl = (const char*)memchr(p, '\n', (size_t)(r.e - p));
            const char* b = p; const char* e = nl ? nl : r.e;
            p = nl ? nl + 1 : r.e;
            const int64_t idx = r.nlines++;
            if (e > b && e[-1] == '\r') --e;                           // text-mode read on the reference's platform
            if (b == e) { r.blank++; if (r.first_blank < 0) r.first_blank = idx; continue; }
            if (*b == '/') continue;
            if (*b == '*') { r.settings.push_back({ idx, Line{ b, e } }); continue; }
            r.obj.push_back(Line{ b, e });
            r.obj_local.push_back(idx);
        }
    });

    // ---- B: numbering, settings ----
    int64_t nlines = 0, n = 0;
    for (Range& r : rg) { r.line_base = nlines; r.obj_base = n; nlines += r.nlines; n += (int64_t)r.obj.size(); }
    for (Range& r : rg) {
        if (r.blank) {
            hs->skipped += r.blank;
            if (hs->first_warning.empty()) hs->first_warning = "line " + std::to_string(r.line_base + r.first_blank + 1) + ": blank line skipped";
        }
    }
    for (Range& r : rg)
        for (auto& sl : r.settings) {
            try { parse_settings_line(sl.second.b, sl.second.e, (uint64_t)(r.line_base + sl.first + 1), hs->tex_paths, hs->settings); }
            catch (ParseError& pe) { drb_set_error("%s", pe.msg.c_str()); delete hs; return DRB_ERR_PARSE; }
        }

    // ---- C: objects ----
    hs->objects.resize((size_t)n);                                     // default-initialised: first touched by its parser
    run([&](int t) {
        Range& r = rg[(size_t)t];
        for (size_t k = 0; k < r.obj.size(); ++k) {
            const int64_t line = r.line_base + r.obj_local[k];         // zero-based
            drb_object& o = hs->objects[(size_t)r.obj_base + k];
            try { parse_object_line(r.obj[k].b, r.obj[k].e, (uint64_t)line + 1, hs->tex_paths, o); }
            catch (ParseError& pe) { r.err_line = line; r.err = pe.msg; return; }
            // objects the tracer cannot represent are kept (ids are line indices) but reported
            const bool ok = (o.type == 2 && o.ncols >= 16) || (o.type == 0 && o.ncols >= 10);
            if (!ok) {
                if (!r.bad++) { r.first_bad_line = line; r.first_bad_type = o.type; r.first_bad_cols = o.ncols; }
            }
        }
    });
    for (Range& r : rg)                                                // ranges are in file order: the first error wins
        if (r.err_line >= 0) { drb_set_error("%s", r.err.c_str()); delete hs; return DRB_ERR_PARSE; }
    for (Range& r : rg) {
        if (!r.bad) continue;
        hs->skipped += r.bad;
        if (hs->first_warning.empty())
            hs->first_warning = "line " + std::to_string(r.first_bad_line + 1) + ": object with type " + std::to_string(r.first_bad_type) + " and " +
                                std::to_string(r.first_bad_cols) + " columns is not renderable (types: 0 sphere, 2 triangle) and is left out of the tree";
    }
    if (!hs->first_warning.empty()) drb_set_error("%s", hs->first_warning.c_str());
    *out = hs;
    return DRB_OK;
}

// ---- binary scene cache ("optional binary cache keyed by file hash") -------------------
// 64-bit content hash: 1 MiB chunks hashed independently (multiply-fold over 8-byte words, threads over
// chunks), chunk hashes folded in order.  Not cryptographic; it only has to notice an edited scene file.
inline uint64_t fold64(uint64_t a, uint64_t b)
{
    unsigned __int128 r = (unsigned __int128)a * b;
    return (uint64_t)r ^ (uint64_t)(r >> 64);
}
constexpr uint64_t kHashK0 = 0x9E3779B97F4A7C15ull, kHashK1 = 0xD6E8FEB86659FD93ull;
constexpr size_t kHashChunk = size_t(1) << 20;

uint64_t hash_chunk(const unsigned char* p, size_t n, uint64_t index)
{
    uint64_t h = fold64(index + kHashK0, n ^ kHashK1);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) { uint64_t w; memcpy(&w, p + i, 8); h = fold64(h ^ w, kHashK0) + kHashK1; }
    if (i < n) { uint64_t w = 0; memcpy(&w, p + i, n - i); h = fold64(h ^ w, kHashK1) + kHashK0; }
    return h;
}

uint64_t hash_bytes(const void* data, size_t len)
{
    const unsigned char* p = (const unsigned char*)data;
    const size_t nchunks = (len + kHashChunk - 1) / kHashChunk;
    std::vector<uint64_t> hc(nchunks);
    auto work = [&](size_t lo, size_t hi) {
        for (size_t c = lo; c < hi; ++c) hc[c] = hash_chunk(p + c * kHashChunk, std::min(kHashChunk, len - c * kHashChunk), c);
    };
    unsigned hw = std::thread::hardware_concurrency();
    size_t nthreads = std::min<size_t>(hw ? hw : 4, std::max<size_t>(1, nchunks / 8));
    if (nthreads <= 1) work(0, nchunks);
    else {
        std::vector<std::thread> pool;
        for (size_t t = 0; t < nthreads; ++t) pool.emplace_back(work, nchunks * t / nthreads, nchunks * (t + 1) / nthreads);
        for (auto& th : pool) th.join();
    }
    uint64_t h = fold64(len ^ kHashK1, kHashK0);
    for (uint64_t c : hc) h = fold64(h ^ c, kHashK0) + kHashK1;
    return h;
}

// texture ids inside objects and settings.backtex are indices into the scanned texture list, so the list is
// part of the key
uint64_t hash_tex_list(const std::vector<std::string>& tex)
{
    std::string all;
    for (const auto& t : tex) { all += t; all.push_back('\0'); }
    return hash_bytes(all.data(), all.size());
}

struct CacheHeader {
    char magic[8];                 // "DRBSCN01"
    uint32_t abi_version, sizeof_object, sizeof_settings, warning_len;
    uint64_t source_hash, source_len, tex_hash, nobjects;
    int64_t skipped;
    drb_settings settings;
    uint32_t pad;
};
static_assert(sizeof(CacheHeader) % 8 == 0, "objects follow the header and the warning text 8-byte aligned");
const char kCacheMagic[8] = { 'D', 'R', 'B', 'S', 'C', 'N', '0', '1' };

// returns a scene on a hit, nullptr on any mismatch or damage (the caller then parses the text)
drb_host_scene* cache_read(const std::string& path, uint64_t source_hash, uint64_t source_len, const std::vector<std::string>& tex)
{
    int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return nullptr;
    struct stat st;
    drb_host_scene* hs = nullptr;
    if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && (size_t)st.st_size >= sizeof(CacheHeader)) {
        size_t len = (size_t)st.st_size;
        void* m = mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m != MAP_FAILED) {
            CacheHeader h;
            memcpy(&h, m, sizeof h);
            size_t wl = ((size_t)h.warning_len + 7) & ~size_t(7);
            bool ok = memcmp(h.magic, kCacheMagic, 8) == 0 && h.abi_version == (uint32_t)DRB_ABI_VERSION &&
                      h.sizeof_object == sizeof(drb_object) && h.sizeof_settings == sizeof(drb_settings) &&
                      h.source_hash == source_hash && h.source_len == source_len && h.tex_hash == hash_tex_list(tex) &&
                      h.warning_len < (1u << 20) && h.nobjects <= (len - sizeof h) / sizeof(drb_object) &&
                      len == sizeof h + wl + (size_t)h.nobjects * sizeof(drb_object);
            if (ok) {
                hs = new drb_host_scene();
                hs->settings = h.settings;
                hs->tex_paths = tex;
                hs->skipped = h.skipped;
                const char* q = (const char*)m + sizeof h;
                hs->first_warning.assign(q, h.warning_len);
                const drb_object* o = (const drb_object*)(q + wl);
                // copy in parallel: page faults on both sides are most of the cost of a 100+ MB copy
                const size_t nobj = (size_t)h.nobjects;
                hs->objects.resize(nobj);                              // default-initialised, see drb_default_init_alloc
                unsigned hw = std::thread::hardware_concurrency();
                const size_t nt = std::min<size_t>(hw ? hw : 4, std::max<size_t>(1, nobj / 65536));
                auto copy = [&](size_t lo, size_t hi) { if (hi > lo) memcpy(hs->objects.data() + lo, o + lo, (hi - lo) * sizeof(drb_object)); };
                if (nt <= 1) copy(0, nobj);
                else {
                    std::vector<std::thread> pool;
                    for (size_t t = 0; t < nt; ++t) pool.emplace_back(copy, nobj * t / nt, nobj * (t + 1) / nt);
                    for (auto& th : pool) th.join();
                }
            }
            munmap(m, len);
        }
    }
    close(fd);
    return hs;
}

// best effort: a cache that cannot be written is not an error of the load
bool cache_write(const std::string& path, const drb_host_scene& hs, uint64_t source_hash, uint64_t source_len)
{
    CacheHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, kCacheMagic, 8);
    h.abi_version = (uint32_t)DRB_ABI_VERSION;
    h.sizeof_object = (uint32_t)sizeof(drb_object);
    h.sizeof_settings = (uint32_t)sizeof(drb_settings);
    h.warning_len = (uint32_t)std::min<size_t>(hs.first_warning.size(), (1u << 20) - 1);
    h.source_hash = source_hash; h.source_len = source_len;
    h.tex_hash = hash_tex_list(hs.tex_paths);
    h.nobjects = hs.objects.size();
    h.skipped = hs.skipped;
    h.settings = hs.settings;
    std::string tmp = path + ".tmp." + std::to_string((long)getpid());
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return false;
    static const char zeros[8] = { 0 };
    size_t wl = ((size_t)h.warning_len + 7) & ~size_t(7);
    bool ok = fwrite(&h, sizeof h, 1, f) == 1;
    if (ok && h.warning_len) ok = fwrite(hs.first_warning.data(), 1, h.warning_len, f) == h.warning_len;
    if (ok && wl > h.warning_len) ok = fwrite(zeros, 1, wl - h.warning_len, f) == wl - h.warning_len;
    if (ok && h.nobjects) ok = fwrite(hs.objects.data(), sizeof(drb_object), (size_t)h.nobjects, f) == (size_t)h.nobjects;
    ok = (fclose(f) == 0) && ok;
    if (ok) ok = rename(tmp.c_str(), path.c_str()) == 0;      // readers never see a half-written file
    if (!ok) unlink(tmp.c_str());
    return ok;
}

} // namespace

void drb_set_error(const char* fmt, ...)
{
    char buf[1024];
    va_list ap; va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
}
void drb_clear_error() { g_last_error.clear(); }

int drb_find_texture(const std::vector<std::string>& tex_paths, const std::string& query)
{
    for (size_t i = 0; i < tex_paths.size(); ++i) {
        std::string low = tex_paths[i];
        std::transform(low.begin(), low.end(), low.begin(), [](unsigned char c) { return (char)::tolower(c); });
        if (low.find(query) != std::string::npos) return (int)i;
    }
    return -1;
}

std::vector<std::string> drb_scan_textures(const char* tex_dir)
{
    std::vector<std::string> found;
    std::error_code ec;
    std::filesystem::path dir = (tex_dir && tex_dir[0]) ? std::filesystem::path(tex_dir) : std::filesystem::current_path(ec);
    if (ec) return found;
    for (const auto& e : std::filesystem::directory_iterator(dir, ec)) {
        std::string name = e.path().filename().string();
        if (name.find("ppm") != std::string::npos || name.find("PPM") != std::string::npos) found.push_back(e.path().string());
    }
    std::sort(found.begin(), found.end());
    return found;
}

void drb_host_scene_summarise(const drb_host_scene* hs)
{
    if (!hs || hs->renderable >= 0) return;
    const size_t n = hs->objects.size(), ntex = hs->tex_paths.size();
    const unsigned hw = std::thread::hardware_concurrency();
    const int nt = (int)std::max<size_t>(1, std::min<size_t>(hw ? hw : 1, n / 65536 + 1));
    std::vector<int64_t> part((size_t)nt, 0);
    std::vector<std::vector<char>> used((size_t)nt, std::vector<char>(ntex, 0));
    auto count = [&](int t) {
        const size_t lo = n * (size_t)t / (size_t)nt, hi = n * ((size_t)t + 1) / (size_t)nt;
        int64_t c = 0;
        std::vector<char>& u = used[(size_t)t];
        for (size_t i = lo; i < hi; ++i) {
            const drb_object& o = hs->objects[i];
            c += drb_object_renderable(o) ? 1 : 0;
            if (o.texnum >= 0 && (size_t)o.texnum < ntex) u[(size_t)o.texnum] = 1;
            if (o.rtexnum >= 0 && (size_t)o.rtexnum < ntex) u[(size_t)o.rtexnum] = 1;
        }
        part[(size_t)t] = c;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(count, t);
    count(0);
    for (auto& th : pool) th.join();
    int64_t total = 0;
    hs->tex_used.assign(ntex, 0);
    for (int t = 0; t < nt; ++t) {
        total += part[(size_t)t];
        for (size_t k = 0; k < ntex; ++k) hs->tex_used[k] |= used[(size_t)t][k];
    }
    if (hs->settings.backtex >= 0 && (size_t)hs->settings.backtex < ntex) hs->tex_used[(size_t)hs->settings.backtex] = 1;
    hs->renderable = total;
}

extern "C" {

const char* drb_last_error(void) { return g_last_error.c_str(); }
int drb_abi_version(void) { return DRB_ABI_VERSION; }

void drb_settings_default(drb_settings* s)
{
    if (!s) return;
    memset(s, 0, sizeof *s);
    s->cam[0] = 0; s->cam[1] = 0; s->cam[2] = 2;        // kernel.cu:125
    s->aperture = 0.01f;                                // :127
    s->focus = 3;                                       // :128
    s->fov = 45; s->max_depth = 50; s->spp = 1;         // :130-132
    s->bg_intensity = 1;                                // :108-109
    s->backtex = -1;                                    // :123
    s->width = 1280; s->height = 720;                   // :29-30
}

int drb_host_scene_parse(const char* text, size_t len, const char* tex_dir, drb_host_scene** out)
{
    if (!out || (!text && len)) { drb_set_error("drb_host_scene_parse: null argument"); return DRB_ERR_ARG; }
    drb_clear_error();
    *out = nullptr;
    return parse_buffer(text ? text : "", len, tex_dir, out);
}

int drb_host_scene_load(const char* rts_path, const char* tex_dir, drb_host_scene** out)
{
    if (!rts_path || !out) { drb_set_error("drb_host_scene_load: null argument"); return DRB_ERR_ARG; }
    drb_clear_error();
    *out = nullptr;
    int fd = open(rts_path, O_RDONLY);
    if (fd < 0) { drb_set_error("cannot open scene file '%s': %s", rts_path, strerror(errno)); return DRB_ERR_IO; }
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { close(fd); drb_set_error("'%s' is not a regular file", rts_path); return DRB_ERR_IO; }
    size_t len = (size_t)st.st_size;
    int rc;
    if (len == 0) rc = parse_buffer("", 0, tex_dir, out);
    else {
        void* m = mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { close(fd); drb_set_error("mmap of '%s' failed: %s", rts_path, strerror(errno)); return DRB_ERR_IO; }
        madvise(m, len, MADV_SEQUENTIAL);
        rc = parse_buffer((const char*)m, len, tex_dir, out);
        munmap(m, len);
    }
    close(fd);
    return rc;
}

uint64_t drb_hash_bytes(const void* data, size_t len) { return (data || !len) ? hash_bytes(data, len) : 0; }

int drb_host_scene_load_cached(const char* rts_path, const char* tex_dir, const char* cache_path, drb_host_scene** out, int* cache_hit)
{
    if (!rts_path || !out) { drb_set_error("drb_host_scene_load_cached: null argument"); return DRB_ERR_ARG; }
    drb_clear_error();
    *out = nullptr;
    if (cache_hit) *cache_hit = 0;
    std::string cpath = (cache_path && cache_path[0]) ? std::string(cache_path) : std::string(rts_path) + ".drbcache";
    int fd = open(rts_path, O_RDONLY);
    if (fd < 0) { drb_set_error("cannot open scene file '%s': %s", rts_path, strerror(errno)); return DRB_ERR_IO; }
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { close(fd); drb_set_error("'%s' is not a regular file", rts_path); return DRB_ERR_IO; }
    size_t len = (size_t)st.st_size;
    void* m = nullptr;
    if (len) {
        m = mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { close(fd); drb_set_error("mmap of '%s' failed: %s", rts_path, strerror(errno)); return DRB_ERR_IO; }
    }
    const uint64_t h = hash_bytes(m, len);
    int rc = DRB_OK;
    drb_host_scene* hs = cache_read(cpath, h, len, drb_scan_textures(tex_dir));
    if (hs) {
        if (cache_hit) *cache_hit = 1;
        if (!hs->first_warning.empty()) drb_set_error("%s", hs->first_warning.c_str());
        *out = hs;
    } else {
        rc = parse_buffer(len ? (const char*)m : "", len, tex_dir, out);
        if (rc == DRB_OK) cache_write(cpath, **out, h, len);
    }
    if (m) munmap(m, len);
    close(fd);
    return rc;
}

int drb_host_scene_create(const drb_settings* settings, const drb_object* objects, int64_t nobjects,
                          const char* const* tex_paths, int ntex, drb_host_scene** out)
{
    if (!out || nobjects < 0 || (nobjects && !objects) || ntex < 0 || (ntex && !tex_paths)) { drb_set_error("drb_host_scene_create: bad argument"); return DRB_ERR_ARG; }
    drb_clear_error();
    auto hs = new drb_host_scene();
    if (settings) hs->settings = *settings; else drb_settings_default(&hs->settings);
    hs->objects.assign(objects, objects + nobjects);
    for (int i = 0; i < ntex; ++i) hs->tex_paths.emplace_back(tex_paths[i] ? tex_paths[i] : "");
    *out = hs;
    return DRB_OK;
}

int64_t drb_host_scene_num_renderable(const drb_host_scene* hs)
{
    if (!hs) return 0;
    drb_host_scene_summarise(hs);
    return hs->renderable;
}

void drb_host_scene_free(drb_host_scene* hs) { if (hs) { drb_host_scene_unpin(hs); delete hs; } }
int64_t drb_host_scene_num_objects(const drb_host_scene* hs) { return hs ? (int64_t)hs->objects.size() : 0; }
const drb_object* drb_host_scene_objects(const drb_host_scene* hs) { return hs && !hs->objects.empty() ? hs->objects.data() : nullptr; }
int drb_host_scene_settings(const drb_host_scene* hs, drb_settings* out)
{
    if (!hs || !out) { drb_set_error("drb_host_scene_settings: null argument"); return DRB_ERR_ARG; }
    *out = hs->settings;
    return DRB_OK;
}
int drb_host_scene_num_textures(const drb_host_scene* hs) { return hs ? (int)hs->tex_paths.size() : 0; }
const char* drb_host_scene_texture_path(const drb_host_scene* hs, int i)
{
    return (hs && i >= 0 && i < (int)hs->tex_paths.size()) ? hs->tex_paths[(size_t)i].c_str() : "";
}
int64_t drb_host_scene_num_skipped(const drb_host_scene* hs) { return hs ? hs->skipped : 0; }

// writer in the exporter's format: plugin/rtsexport.py:205-207 (header + settings), :312-314 (objects)
int drb_rts_write(const char* path, const drb_settings* s, const drb_object* objs, int64_t n,
                  const char* const* tex_names, int ntex, const char* backtex_name)
{
    if (!path || !s || n < 0 || (n && !objs)) { drb_set_error("drb_rts_write: bad argument"); return DRB_ERR_ARG; }
    FILE* f = fopen(path, "w");
    if (!f) { drb_set_error("cannot create '%s': %s", path, strerror(errno)); return DRB_ERR_IO; }
    std::vector<char> iobuf(1 << 22);
    setvbuf(f, iobuf.data(), _IOFBF, iobuf.size());
    auto texname = [&](int k) -> const char* { return (k >= 0 && k < ntex && tex_names && tex_names[k]) ? tex_names[k] : "no"; };
    fprintf(f, "/exported from dogeray_b200\n");
    fprintf(f, "*,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%s,%i,%i\n", s->cam[0], s->cam[1], s->cam[2], s->aperture,
            s->look[0], s->look[1], s->look[2], s->focus, (double)s->fov, (double)s->max_depth, (double)s->spp, s->bg_intensity,
            backtex_name ? backtex_name : texname(s->backtex), s->width, s->height);
    // object lines are formatted in parallel, one contiguous range per thread, then written in order
    auto format_range = [&](int64_t lo, int64_t hi, std::string& out) {
        out.reserve((size_t)(hi - lo) * 400);
        char line[1024];
        for (int64_t i = lo; i < hi; ++i) {
            const drb_object& o = objs[i];
            const int k = snprintf(line, sizeof line,
                "%f,%f,%f,%d,%f,%f,%f,%f,%g,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%f,%s,%s\n",
                o.pos[0], o.pos[1], o.pos[2], o.type, o.col[0], o.col[1], o.col[2], o.add_y, o.add_x,
                o.dim[0], o.dim[1], o.dim[2], (double)o.mat, o.rot[0], o.rot[1], o.rot[2],
                o.norm[0], o.norm[1], o.norm[2], o.n1[0], o.n1[1], o.n1[2], o.n2[0], o.n2[1], o.n2[2], o.n3[0], o.n3[1], o.n3[2],
                o.t1[0], o.t1[1], o.t2[0], o.t2[1], o.t3[0], o.t3[1], (double)o.smooth, (double)o.checker,
                texname(o.texnum), texname(o.rtexnum));
            if (k > 0) out.append(line, (size_t)std::min<int>(k, (int)sizeof line - 1));
        }
    };
    unsigned hw = std::thread::hardware_concurrency();
    const int nthreads = (int)std::min<int64_t>(hw ? hw : 4, std::max<int64_t>(1, n / 20000));
    const int64_t chunk = 65536;                                  // bounded memory: format nthreads chunks at a time
    for (int64_t base = 0; base < n; base += chunk * nthreads) {
        std::vector<std::string> parts((size_t)nthreads);
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t) {
            const int64_t lo = std::min(n, base + chunk * t), hi = std::min(n, lo + chunk);
            if (lo < hi) pool.emplace_back(format_range, lo, hi, std::ref(parts[(size_t)t]));
        }
        for (auto& th : pool) th.join();
        for (auto& part : parts)
            if (!part.empty() && fwrite(part.data(), 1, part.size(), f) != part.size()) break;
    }
    bool ok = !ferror(f);
    ok = (fclose(f) == 0) && ok;
    if (!ok) { drb_set_error("short write to '%s'", path); return DRB_ERR_IO; }
    return DRB_OK;
}

} // extern "C"
