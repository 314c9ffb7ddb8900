/* TEST INFRASTRUCTURE -- stand-in for the CUDA Samples' <helper_functions.h>
 * (v11.2 common/inc, not vendored by the reference).  kernel.cu uses exactly one
 * symbol from it, sdkLoadPPM4 (kernel.cu:1926): load a P6/P5 file as RGBA8 with
 * alpha 0, rows top-down, allocating *data when it is NULL. */
#ifndef DOGERAY_ORACLE_STUB_HELPER_FUNCTIONS_H
#define DOGERAY_ORACLE_STUB_HELPER_FUNCTIONS_H
#include <cstdio>
#include <cstdlib>
#include <cstring>
static inline bool sdkLoadPPM4(const char* file, unsigned char** data, unsigned int* w, unsigned int* h)
{
    *w = 1; *h = 1;
    FILE* f = fopen(file, "rb");
    char magic[3] = {0, 0, 0};
    int vals[3] = {1, 1, 255}, got = 0, ch = 3;
    bool ok = f && fscanf(f, "%2s", magic) == 1 && (!strcmp(magic, "P6") || !strcmp(magic, "P5"));
    if (ok) ch = magic[1] == '6' ? 3 : 1;
    while (ok && got < 3) {
        int c = fgetc(f);
        if (c == EOF) { ok = false; break; }
        if (c == '#') { while (c != '\n' && c != EOF) c = fgetc(f); continue; }
        if (c == ' ' || c == '\t' || c == '\n' || c == '\r') continue;
        ungetc(c, f);
        if (fscanf(f, "%d", &vals[got]) != 1) ok = false;
        ++got;
    }
    if (ok) fgetc(f);
    size_t n = ok ? (size_t)vals[0] * vals[1] : 1;
    if (!*data) *data = (unsigned char*)calloc(n * 4, 1);
    if (ok) {
        *w = (unsigned)vals[0]; *h = (unsigned)vals[1];
        unsigned char* raw = (unsigned char*)malloc(n * ch);
        if (fread(raw, 1, n * ch, f) != n * ch) ok = false;
        for (size_t i = 0; ok && i < n; ++i)
            for (int k = 0; k < 3; ++k) (*data)[4 * i + k] = raw[ch == 3 ? 3 * i + k : i];
        free(raw);
    }
    if (f) fclose(f);
    return ok;
}
#endif
