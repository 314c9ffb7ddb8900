/* A plain C99 consumer of include/dogeray_b200.h, the way a maintainer of the reference would link the library
 * (INTEGRATION.md 1): compiled with -std=c99 -pedantic -Wall -Werror by tests/test_abi.py.  Prints the struct sizes
 * the C compiler sees, parses a two-triangle scene, and -- when a CUDA device is present -- renders a 16x8 frame
 * through drb_render and drb_frame_i3.  Exit status 0 = everything it could check held. */
#include "dogeray_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const char* SCENE =
    "*,0,0,4,0.0,0,0,0,4,45,3,2,1,no,16,8\n"
    "-1,-1,0,2,0.8,0.2,0.2,0.0,0,1,-1,0,1,0,1,0\n"
    "-1,-1,0,2,0.2,0.8,0.2,0.0,0,0,1,0,1,-1,1,0\n";

int main(void)
{
    drb_host_scene* hs = NULL;
    drb_scene* scene = NULL;
    drb_settings st;
    drb_opts opts;
    drb_stats stats;
    int rc, i, lit = 0;
    printf("sizes settings=%u object=%u opts=%u stats=%u build_info=%u abi=%d\n", (unsigned)sizeof(drb_settings), (unsigned)sizeof(drb_object),
           (unsigned)sizeof(drb_opts), (unsigned)sizeof(drb_stats), (unsigned)sizeof(drb_build_info), drb_abi_version());
    if (drb_host_scene_parse(SCENE, strlen(SCENE), "/nonexistent-texture-dir", &hs) != DRB_OK) { printf("parse: %s\n", drb_last_error()); return 1; }
    if (drb_host_scene_num_objects(hs) != 2) { printf("expected 2 objects\n"); return 1; }
    if (drb_host_scene_settings(hs, &st) != DRB_OK || st.width != 16 || st.height != 8 || st.spp != 2 || st.max_depth != 3) { printf("settings\n"); return 1; }
    printf("parsed objects=%ld width=%d height=%d\n", (long)drb_host_scene_num_objects(hs), st.width, st.height);
    rc = drb_scene_create(hs, 0, &scene);
    if (drb_device_count() <= 0) {
        printf("no device: drb_scene_create -> %d (%s)\n", rc, drb_last_error());
        drb_host_scene_free(hs);
        return rc == DRB_ERR_CUDA && scene == NULL ? 0 : 1;
    }
    if (rc != DRB_OK) { printf("create: %s\n", drb_last_error()); return 1; }
    {
        float accum[16 * 8 * 3];
        int32_t frame[16 * 8 * 3];
        drb_opts_default(&opts);
        opts.seed = 3;
        if (drb_render(scene, &st, &opts, accum, &stats) != DRB_OK) { printf("render: %s\n", drb_last_error()); return 1; }
        for (i = 0; i < 16 * 8 * 3; ++i) lit += accum[i] > 0.0f;
        memset(frame, 0, sizeof frame);
        if (drb_frame_i3(scene, &st, &opts, 1, frame) != DRB_OK) { printf("frame: %s\n", drb_last_error()); return 1; }
        /* out[x*H + y] = trunc(255 * mean): the same numbers as the float sums, in CudaStarter's layout */
        for (i = 0; i < 16 * 8; ++i) {
            const int x = i / 8, y = i % 8, c = 1;
            const float mean = accum[(y * 16 + x) * 3 + c] * (1.0f / 2.0f);
            if (frame[i * 3 + c] != (int32_t)(mean * 255.0f)) { printf("frame_i3 mismatch at x=%d y=%d: %d vs %f\n", x, y, (int)frame[i * 3 + c], mean * 255.0f); return 1; }
        }
        printf("rendered paths=%lu rays=%lu lit=%d\n", (unsigned long)stats.paths, (unsigned long)stats.rays, lit);
        if (stats.paths != 16u * 8u * 2u || stats.rays < stats.paths || lit == 0) return 1;
        /* the device list of one process: two handles on device 0 stand in for two GPUs */
        {
            int devices[2] = { 0, 0 };
            drb_scene* pair[2] = { NULL, NULL };
            float multi[16 * 8 * 3];
            drb_stats ms;
            if (drb_scene_create_multi(hs, devices, 2, 0u, pair) != DRB_OK) { printf("create_multi: %s\n", drb_last_error()); return 1; }
            opts.flags = DRB_FLAG_DYNAMIC_TILES;
            if (drb_render_multi(pair, 2, &st, &opts, multi, &ms) != DRB_OK) { printf("render_multi: %s\n", drb_last_error()); return 1; }
            if (memcmp(multi, accum, sizeof multi) != 0 || ms.rays != stats.rays) { printf("render_multi differs from drb_render\n"); return 1; }
            printf("multi paths=%lu rays=%lu identical\n", (unsigned long)ms.paths, (unsigned long)ms.rays);
            drb_scene_free(pair[0]); drb_scene_free(pair[1]);
        }
    }
    drb_scene_free(scene);
    drb_host_scene_free(hs);
    return 0;
}
