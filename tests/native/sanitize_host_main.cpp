// Test harness: the host-side translation units (rts_loader.cpp, image_io.cpp) built with
// -fsanitize=address,undefined and driven over scene files given on the command line plus a list of malformed
// inputs (tests/test_sanitizers.py).  Loads each file plainly, through a cold and a warm scene cache, compares the
// three and writes the scene back out.  Exit status = number of mismatches; sanitizer reports abort the process.
#include "dogeray_b200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <filesystem>
// stand-in for the symbol scene.cu provides
struct drb_host_scene;
void drb_host_scene_unpin(drb_host_scene*) {}
int main(int argc, char** argv)
{
    int bad = 0;
    for (int i = 1; i < argc; ++i) {
        std::string dir = std::filesystem::path(argv[i]).parent_path().string();
        drb_host_scene* a = nullptr; drb_host_scene* b = nullptr; drb_host_scene* c = nullptr;
        int hit = 0;
        std::string cache = std::string(getenv("DRB_SAN_TMP") ? getenv("DRB_SAN_TMP") : "/tmp") + "/san_cache.bin";
        remove(cache.c_str());
        int r0 = drb_host_scene_load(argv[i], dir.c_str(), &a);
        int r1 = drb_host_scene_load_cached(argv[i], dir.c_str(), cache.c_str(), &b, &hit);
        int r2 = drb_host_scene_load_cached(argv[i], dir.c_str(), cache.c_str(), &c, &hit);
        if (r0 != r1 || r1 != r2) { printf("%s: rc %d %d %d\n", argv[i], r0, r1, r2); bad++; }
        if (r0 == 0) {
            int64_t n = drb_host_scene_num_objects(a);
            if (!hit || n != drb_host_scene_num_objects(c) || (n && memcmp(drb_host_scene_objects(a), drb_host_scene_objects(c), n * sizeof(drb_object)))) { printf("%s: mismatch\n", argv[i]); bad++; }
            drb_settings st; drb_host_scene_settings(a, &st);
            std::string out = std::string(getenv("DRB_SAN_TMP") ? getenv("DRB_SAN_TMP") : "/tmp") + "/san_out.rts";
            if (drb_rts_write(out.c_str(), &st, drb_host_scene_objects(a), n, nullptr, 0, nullptr) != 0) { printf("%s: write failed\n", argv[i]); bad++; }
        }
        drb_host_scene_free(a); drb_host_scene_free(b); drb_host_scene_free(c);
    }
    // truncated / garbage text
    const char* junk[] = { "", "\n\n", "*", "*,1", "2", "2,", ",,,,", "1,2,3,2,1,1,1,0,0,1,2,3,0,1,2,3", "r,r,r,0,r,r,r,0,0,r", "*,1,2,3,4,5,6,7,8,9,10,11,12,x.ppm,10,10\n/", "1e99999,2,3,2" };
    for (const char* j : junk) { drb_host_scene* h = nullptr; drb_host_scene_parse(j, strlen(j), "/tmp", &h); drb_host_scene_free(h); }
    // image writers and the .ppm reader, including headers that lie about the size
    {
        const std::string tmp = getenv("DRB_SAN_TMP") ? getenv("DRB_SAN_TMP") : "/tmp";
        const int W = 13, H = 7;
        std::vector<uint8_t> rgb((size_t)W * H * 3);
        for (size_t i = 0; i < rgb.size(); ++i) rgb[i] = (uint8_t)(i * 37);
        if (drb_write_bmp((tmp + "/san.bmp").c_str(), rgb.data(), W, H) != 0 || drb_write_ppm((tmp + "/san.ppm").c_str(), rgb.data(), W, H) != 0) bad++;
        uint8_t* px = nullptr; int w = 0, h = 0;
        if (drb_read_ppm((tmp + "/san.ppm").c_str(), &px, &w, &h) != 0 || w != W || h != H) bad++;
        else for (int i = 0; i < W * H; ++i) if (memcmp(px + 4 * i, rgb.data() + 3 * i, 3) != 0 || px[4 * i + 3] != 0) { bad++; break; }
        drb_free(px);
        const char* lies[] = { "", "P6", "P6\n", "P6\n4 4\n255\n", "P6\n99999 99999\n255\nabc", "P6\n-3 4\n255\n", "P3\n1 1\n255\n1 2 3", "P6\n# c\n2 1\n255\nabcdef",
                               "P6\n2 2\n65535\nabcdefgh" };
        for (const char* l : lies) {
            FILE* f = fopen((tmp + "/lie.ppm").c_str(), "wb"); fwrite(l, 1, strlen(l), f); fclose(f);
            px = nullptr;
            drb_read_ppm((tmp + "/lie.ppm").c_str(), &px, &w, &h);
            drb_free(px);
        }
    }
    printf("files %d bad %d\n", argc - 1, bad);
    return bad;
}
