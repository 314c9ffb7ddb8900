// Image files of the hot path: .ppm textures in, .bmp / .ppm frames out, and the host display
// transform.  Replaces sdkLoadPPM4 (called at raygpu/kernel.cu:1926), SDL_SaveBMP (kernel.cu:2513)
// and the per-pixel clamp/divide of the draw loop (kernel.cu:2287).
#include "drb_internal.h"

#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

// P6 (RGB) or P5 (grey) with maxval <= 255; '#' comments allowed in the header.
// Output RGBA8, alpha 0, rows top-down: what readtextures copies into its uchar4 array.
int drb_load_ppm(const std::string& path, drb_image& out)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { drb_set_error("cannot open texture '%s': %s", path.c_str(), strerror(errno)); return DRB_ERR_IO; }
    auto fail = [&](const char* why) { fclose(f); drb_set_error("texture '%s': %s", path.c_str(), why); return DRB_ERR_PARSE; };
    int c0 = fgetc(f), c1 = fgetc(f);
    if (c0 != 'P' || (c1 != '6' && c1 != '5')) return fail("not a P6/P5 file");
    const int ch = (c1 == '6') ? 3 : 1;
    long vals[3] = {0, 0, 0};
    for (int got = 0; got < 3;) {
        int c = fgetc(f);
        if (c == EOF) return fail("truncated header");
        if (c == '#') { while (c != '\n' && c != EOF) c = fgetc(f); continue; }
        if (c == ' ' || c == '\t' || c == '\n' || c == '\r') continue;
        if (c < '0' || c > '9') return fail("bad header");
        long v = 0;
        while (c >= '0' && c <= '9') { v = v * 10 + (c - '0'); if (v > (1L << 30)) return fail("header value too large"); c = fgetc(f); }
        // c is now the single separator after the number (for maxval: the byte before the raster)
        vals[got++] = v;
        if (got == 3 && !(c == ' ' || c == '\t' || c == '\n' || c == '\r')) return fail("bad header");
    }
    if (vals[0] <= 0 || vals[1] <= 0 || vals[0] > 65536 || vals[1] > 65536) return fail("bad dimensions");
    if (vals[2] <= 0 || vals[2] > 255) return fail("only 8-bit maxval is supported");
    const size_t n = (size_t)vals[0] * (size_t)vals[1];
    std::vector<uint8_t> raw(n * ch);
    if (fread(raw.data(), 1, raw.size(), f) != raw.size()) return fail("truncated raster");
    fclose(f);
    out.w = (int)vals[0]; out.h = (int)vals[1];
    out.rgba.assign(n * 4, 0);
    for (size_t i = 0; i < n; ++i) {
        if (ch == 3) { out.rgba[4*i] = raw[3*i]; out.rgba[4*i+1] = raw[3*i+1]; out.rgba[4*i+2] = raw[3*i+2]; }
        else { out.rgba[4*i] = out.rgba[4*i+1] = out.rgba[4*i+2] = raw[i]; }
    }
    return DRB_OK;
}

namespace {
void put16(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
void put32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
}

extern "C" {

int drb_read_ppm(const char* path, uint8_t** rgba, int* width, int* height)
{
    if (!path || !rgba || !width || !height) { drb_set_error("drb_read_ppm: null argument"); return DRB_ERR_ARG; }
    drb_image img;
    int rc = drb_load_ppm(path, img);
    if (rc != DRB_OK) return rc;
    *rgba = (uint8_t*)malloc(img.rgba.size());
    if (!*rgba) { drb_set_error("out of memory"); return DRB_ERR_NOMEM; }
    memcpy(*rgba, img.rgba.data(), img.rgba.size());
    *width = img.w; *height = img.h;
    return DRB_OK;
}

void drb_free(void* p) { free(p); }

// 14-byte file header + 108-byte BITMAPV4HEADER, 32 bpp, BI_BITFIELDS,
// masks R 00FF0000 G 0000FF00 B 000000FF A FF000000, CSType "Win ", bottom-up rows, pixels B,G,R,A=255.
int drb_write_bmp(const char* path, const uint8_t* rgb8, int w, int h)
{
    if (!path || !rgb8 || w <= 0 || h <= 0) { drb_set_error("drb_write_bmp: bad argument"); return DRB_ERR_ARG; }
    const uint32_t off = 14 + 108, img = (uint32_t)w * (uint32_t)h * 4u;
    std::vector<uint8_t> buf((size_t)off + img, 0);
    uint8_t* p = buf.data();
    p[0] = 'B'; p[1] = 'M'; put32(p + 2, off + img); put32(p + 10, off);
    uint8_t* d = p + 14;
    put32(d + 0, 108); put32(d + 4, (uint32_t)w); put32(d + 8, (uint32_t)h); put16(d + 12, 1); put16(d + 14, 32);
    put32(d + 16, 3); put32(d + 20, img);
    put32(d + 40, 0x00FF0000u); put32(d + 44, 0x0000FF00u); put32(d + 48, 0x000000FFu); put32(d + 52, 0xFF000000u);
    put32(d + 56, 0x57696E20u);
    for (int y = 0; y < h; ++y) {
        const uint8_t* src = rgb8 + (size_t)y * w * 3;
        uint8_t* dst = p + off + (size_t)(h - 1 - y) * w * 4;
        for (int x = 0; x < w; ++x) { dst[4*x] = src[3*x+2]; dst[4*x+1] = src[3*x+1]; dst[4*x+2] = src[3*x]; dst[4*x+3] = 255; }
    }
    FILE* f = fopen(path, "wb");
    if (!f) { drb_set_error("cannot create '%s': %s", path, strerror(errno)); return DRB_ERR_IO; }
    bool ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
    ok = (fclose(f) == 0) && ok;
    if (!ok) { drb_set_error("short write to '%s'", path); return DRB_ERR_IO; }
    return DRB_OK;
}

int drb_write_ppm(const char* path, const uint8_t* rgb8, int w, int h)
{
    if (!path || !rgb8 || w <= 0 || h <= 0) { drb_set_error("drb_write_ppm: bad argument"); return DRB_ERR_ARG; }
    FILE* f = fopen(path, "wb");
    if (!f) { drb_set_error("cannot create '%s': %s", path, strerror(errno)); return DRB_ERR_IO; }
    fprintf(f, "P6\n%d %d\n255\n", w, h);
    size_t n = (size_t)w * h * 3;
    bool ok = fwrite(rgb8, 1, n, f) == n;
    ok = (fclose(f) == 0) && ok;
    if (!ok) { drb_set_error("short write to '%s'", path); return DRB_ERR_IO; }
    return DRB_OK;
}

// kernel.cu:1083-1085 then :2287: the kernel stores trunc(255 * mean) per frame, the draw loop clamps
// to [0,255].  With a float accumulator the two steps collapse into one.
int drb_tonemap(const float* accum, int w, int h, double nsamples, uint8_t* rgb8)
{
    if (!accum || !rgb8 || w <= 0 || h <= 0 || !(nsamples > 0)) { drb_set_error("drb_tonemap: bad argument"); return DRB_ERR_ARG; }
    const float scale = (float)(1.0 / nsamples);
    const size_t n = (size_t)w * h * 3;
    for (size_t i = 0; i < n; ++i) {
        float v = accum[i] * 255.0f * scale;
        int q = (v != v) ? 0 : (v >= 255.0f ? 255 : (v <= 0.0f ? 0 : (int)v));
        rgb8[i] = (uint8_t)q;
    }
    return DRB_OK;
}

} // extern "C"
