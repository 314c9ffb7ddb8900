// dogeray-b200: headless drop-in for `raygpu.exe [scene.rts]` (raygpu/kernel.cu:2021-2051, 2486-2516).
//
//   dogeray-b200 [scene.rts] [--spp N] [--depth D] [--res WxH] [--seed S] [--device K | --gpus N | --devices a,b,..]
//                [--dynamic] [--shard tiles|samples] [--out file.bmp|file.ppm]
//                [--snapshot-every N] [--save-acc file.acc] [--resume file.acc] [--cache]
//
// --snapshot-every N renders N samples per pixel at a time and rewrites the image after every chunk (the headless
// stand-in for the window's progressive accumulate loop, kernel.cu:2154-2224); --save-acc / --resume checkpoint the
// float accumulator and its sample count, so a render can be continued later with the next sample indices.
//
// Like the reference it opens `scene.rts` when no path is given, looks for textures among the *.ppm files of the
// current working directory, and writes `<scene path>.bmp` (the file SPACE exports, 32-bpp V4 header).  Unlike the
// reference there is no window: it renders the settings line's samples per pixel once and exits.
#include "dogeray_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static int fail(const char* what)
{
    fprintf(stderr, "dogeray-b200: %s: %s\n", what, drb_last_error());
    return 1;
}

int main(int argc, char** argv)
{
    std::string scene_path = "scene.rts", out_path;
    std::string save_acc, resume_acc;
    int spp = -1, depth = -1, w = -1, h = -1, device = 0, snapshot_every = 0;
    bool use_cache = false, dynamic = false, by_samples = false;
    std::vector<int> devices;                        // --gpus N = 0..N-1, --devices a,b,c = exactly those (repeats allowed)
    unsigned long long seed = 0;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&](const char* name) -> const char* {
            if (i + 1 >= argc) { fprintf(stderr, "dogeray-b200: %s needs a value\n", name); exit(2); }
            return argv[++i];
        };
        if (a == "--spp") spp = atoi(next("--spp"));
        else if (a == "--depth") depth = atoi(next("--depth"));
        else if (a == "--seed") seed = strtoull(next("--seed"), nullptr, 10);
        else if (a == "--device") device = atoi(next("--device"));
        else if (a == "--gpus") {
            int n = atoi(next("--gpus"));
            if (n < 1) { fprintf(stderr, "dogeray-b200: --gpus wants a positive count\n"); return 2; }
            devices.clear();
            for (int k = 0; k < n; ++k) devices.push_back(k);
        } else if (a == "--devices") {
            devices.clear();
            const char* p = next("--devices");
            while (*p) {
                char* e = nullptr;
                long v = strtol(p, &e, 10);
                if (e == p || v < 0) { fprintf(stderr, "dogeray-b200: --devices wants a comma-separated list of device numbers\n"); return 2; }
                devices.push_back((int)v);
                p = (*e == ',') ? e + 1 : e;
                if (*e && *e != ',') { fprintf(stderr, "dogeray-b200: --devices wants a comma-separated list of device numbers\n"); return 2; }
            }
            if (devices.empty()) { fprintf(stderr, "dogeray-b200: --devices wants at least one device\n"); return 2; }
        }
        else if (a == "--out") out_path = next("--out");
        else if (a == "--snapshot-every") snapshot_every = atoi(next("--snapshot-every"));
        else if (a == "--save-acc") save_acc = next("--save-acc");
        else if (a == "--resume") resume_acc = next("--resume");
        else if (a == "--cache") use_cache = true;
        else if (a == "--dynamic") dynamic = true;
        else if (a == "--shard") {
            const std::string v = next("--shard");
            if (v == "samples") by_samples = true;
            else if (v != "tiles") { fprintf(stderr, "dogeray-b200: --shard wants tiles or samples\n"); return 2; }
        }
        else if (a == "--res") { if (sscanf(next("--res"), "%dx%d", &w, &h) != 2) { fprintf(stderr, "dogeray-b200: --res wants WxH\n"); return 2; } }
        else if (a == "-h" || a == "--help") {
            printf("usage: dogeray-b200 [scene.rts] [--spp N] [--depth D] [--res WxH] [--seed S] [--device K | --gpus N | --devices a,b,..]\n"
                   "                    [--out file.bmp|file.ppm]\n"
                   "                    [--snapshot-every N] [--save-acc file.acc] [--resume file.acc] [--cache]\n"
                   "  --gpus / --devices  one frame over several GPUs from this process (interleaved tiles, same image as one GPU;\n"
                   "                      the scene crosses PCIe once, the GPUs exchange it and the image over NVLink)\n"
                   "  --dynamic           with several GPUs: tile shards are claimed from a shared queue (load balancing)\n"
                   "  --shard samples     with several GPUs: split the samples instead of the tiles (sum re-associated)\n"
                   "  --cache  keep the parsed scene in <scene>.drbcache (keyed by a hash of the text) and reuse it\n");
            return 0;
        } else if (!a.empty() && a[0] == '-') { fprintf(stderr, "dogeray-b200: unknown option %s\n", a.c_str()); return 2; }
        else scene_path = a;
    }
    printf("Opening:%s\n", scene_path.c_str());
    drb_host_scene* hs = nullptr;
    int cache_hit = 0;
    if ((use_cache ? drb_host_scene_load_cached(scene_path.c_str(), nullptr, nullptr, &hs, &cache_hit)
                   : drb_host_scene_load(scene_path.c_str(), nullptr, &hs)) != DRB_OK) return fail("cannot load scene");
    if (use_cache) printf("scene cache: %s\n", cache_hit ? "hit" : "miss (written)");
    printf("%lld tris\n%d textures total\n", (long long)drb_host_scene_num_objects(hs), drb_host_scene_num_textures(hs));
    if (drb_host_scene_num_skipped(hs)) fprintf(stderr, "dogeray-b200: warning: %s\n", drb_last_error());
    if (devices.empty()) devices.push_back(device);
    std::vector<drb_scene*> scenes(devices.size(), nullptr);
    printf("Building BVH..\n");
    if (drb_scene_create_multi(hs, devices.data(), (int)devices.size(), 0u, scenes.data()) != DRB_OK) return fail("cannot create device scene");
    drb_scene* scene = scenes[0];
    if (scenes.size() > 1) printf("%zu device scenes (%s)\n", scenes.size(), by_samples ? "sample sharding" : dynamic ? "tile shards from a shared queue" : "interleaved tile sharding");
    drb_build_info bi;
    drb_scene_build_info(scene, &bi);
    printf("Done! %lld nodes total (upload %.2f ms, build %.2f ms)\n", (long long)bi.nnodes, bi.upload_ms, bi.build_ms);
    drb_settings st;
    drb_scene_settings(scene, &st);
    if (spp > 0) st.spp = spp;
    if (depth >= 0) st.max_depth = depth;
    if (w > 0 && h > 0) { st.width = w; st.height = h; }
    drb_opts opts;
    drb_opts_default(&opts);
    opts.seed = seed;
    std::vector<float> accum((size_t)st.width * st.height * 3, 0.0f);
    // checkpoint header: magic, width, height, samples accumulated so far, seed
    struct AccHeader { char magic[8]; int32_t w, h; uint64_t samples, seed; };
    uint64_t have = 0;
    if (!resume_acc.empty()) {
        FILE* f = fopen(resume_acc.c_str(), "rb");
        AccHeader hd;
        if (!f || fread(&hd, sizeof hd, 1, f) != 1 || memcmp(hd.magic, "DRBACC1", 8) != 0 || hd.w != st.width || hd.h != st.height ||
            fread(accum.data(), sizeof(float), accum.size(), f) != accum.size()) {
            fprintf(stderr, "dogeray-b200: cannot resume from %s (missing, truncated, or another image size)\n", resume_acc.c_str());
            return 1;
        }
        fclose(f);
        have = hd.samples; seed = hd.seed;
        printf("resumed %llu samples from %s\n", (unsigned long long)have, resume_acc.c_str());
    }
    if (out_path.empty()) out_path = scene_path + ".bmp";
    const bool ppm = out_path.size() > 4 && out_path.substr(out_path.size() - 4) == ".ppm";
    std::vector<uint8_t> rgb(accum.size());
    auto write_image = [&](uint64_t nsamples) -> int {
        if (drb_tonemap(accum.data(), st.width, st.height, (double)(nsamples ? nsamples : 1), rgb.data()) != DRB_OK) return DRB_ERR_ARG;
        return ppm ? drb_write_ppm(out_path.c_str(), rgb.data(), st.width, st.height) : drb_write_bmp(out_path.c_str(), rgb.data(), st.width, st.height);
    };
    const uint32_t total = (uint32_t)(st.spp > 0 ? st.spp : 0);
    const uint32_t chunk = snapshot_every > 0 ? (uint32_t)snapshot_every : (total ? total : 1);
    double ms_total = 0; uint64_t rays_total = 0;
    for (uint32_t done = 0; done < total; done += chunk) {
        drb_opts opts;
        drb_opts_default(&opts);
        opts.seed = seed;
        opts.sample_base = (uint32_t)have;
        opts.sample_count = total - done < chunk ? total - done : chunk;
        opts.flags = DRB_FLAG_ACCUMULATE | (dynamic ? DRB_FLAG_DYNAMIC_TILES : 0u) | (by_samples ? DRB_FLAG_SHARD_SAMPLES : 0u);
        drb_stats stats;
        if (drb_render_multi(scenes.data(), (int)scenes.size(), &st, &opts, accum.data(), &stats) != DRB_OK) return fail("render failed");
        have += opts.sample_count; ms_total += stats.total_ms; rays_total += stats.rays;
        if (write_image(have) != DRB_OK) return fail("cannot write image");
        if (snapshot_every > 0) printf("%llu samples -> %s\n", (unsigned long long)have, out_path.c_str());
    }
    if (total == 0 && write_image(have) != DRB_OK) return fail("cannot write image");
    printf("Time = %.3f ms  %llu samples  %.1f Mrays/s\n", ms_total, (unsigned long long)have, ms_total > 0 ? rays_total / (ms_total * 1e-3) / 1e6 : 0.0);
    if (!save_acc.empty()) {
        AccHeader hd;
        memset(&hd, 0, sizeof hd);
        memcpy(hd.magic, "DRBACC1", 8); hd.w = st.width; hd.h = st.height; hd.samples = have; hd.seed = seed;
        FILE* f = fopen(save_acc.c_str(), "wb");
        if (!f || fwrite(&hd, sizeof hd, 1, f) != 1 || fwrite(accum.data(), sizeof(float), accum.size(), f) != accum.size() || fclose(f) != 0) {
            fprintf(stderr, "dogeray-b200: cannot write %s\n", save_acc.c_str());
            return 1;
        }
    }
    printf("exported image:%s\n", out_path.c_str());
    for (drb_scene* sc : scenes) drb_scene_free(sc);
    drb_host_scene_free(hs);
    return 0;
}
