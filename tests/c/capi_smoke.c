/* capi_smoke.c — the C ABI used from C: compiled by tests/test_capi_c.py with
 *   gcc -std=c99 -Wall -Wextra -pedantic -Werror
 * against include/depthhead_cuda.h (the header must be valid C, not only C++), linked to
 * libdepthhead_cuda.so.
 *
 *   capi_smoke <model.json>                      host-only entry points (no GPU needed)
 *   capi_smoke <model.json> <frame.u16> <w> <h>  + one dh_predict and a 3-frame dh_predict_batch on
 *                                                device 0; prints "pose x y z r0 r1 r2" per result
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "depthhead_cuda.h"

static char* slurp(const char* path, size_t* len) {
    FILE* f = fopen(path, "rb");
    char* buf;
    long n;
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf = (char*)malloc((size_t)n + 1);
    if (!buf || fread(buf, 1, (size_t)n, f) != (size_t)n) {
        fclose(f);
        free(buf);
        return NULL;
    }
    fclose(f);
    buf[n] = 0;
    *len = (size_t)n;
    return buf;
}

#define CHECK(call)                                                              \
    do {                                                                         \
        int rc_ = (call);                                                        \
        if (rc_ != DH_OK) {                                                      \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, dh_last_error());      \
            return 2;                                                            \
        }                                                                        \
    } while (0)

int main(int argc, char** argv) {
    size_t len = 0, needed = 0;
    char* json;
    dh_forest* forest = NULL;
    dh_ctx* ctx = NULL;
    const float K[9] = {560.0f, 0.0f, 320.0f, 0.0f, 560.0f, 240.0f, 0.0f, 0.0f, 1.0f};
    if (argc < 2) return 64;
    json = slurp(argv[1], &len);
    if (!json) return 65;
    printf("abi %d build %s\n", dh_abi_version(), dh_build_id());
    if (sizeof(dh_result) != 56) return 66;

    /* a broken document is an error code and a message, never a crash */
    if (dh_forest_from_json("{\"stepwidth\": 1", 15, &forest) != DH_E_JSON || forest != NULL) return 67;
    if (strlen(dh_last_error()) == 0) return 68;

    CHECK(dh_forest_from_json(json, len, &forest));
    printf("trees %d nodes %ld leaves %ld votes %ld stepwidth %u iterations %u sigma %g\n", (int)dh_forest_n_trees(forest),
           (long)dh_forest_n_nodes(forest), (long)dh_forest_n_leaves(forest), (long)dh_forest_n_votes(forest),
           dh_forest_get_stepwidth(forest), dh_forest_get_meanshift_iterations(forest), (double)dh_forest_get_sigma(forest));
    CHECK(dh_forest_set_stepwidth(forest, 7u));
    CHECK(dh_forest_set_meanshift_iterations(forest, 11u));
    CHECK(dh_forest_set_sigma(forest, 4.0f));
    CHECK(dh_forest_set_sigma(forest, -1.0f)); /* ignored like update_sigma(val <= 0) */
    if (dh_forest_get_stepwidth(forest) != 7u || dh_forest_get_meanshift_iterations(forest) != 11u || dh_forest_get_sigma(forest) != 4.0f)
        return 69;
    if (dh_forest_set_stepwidth(forest, 0u) != DH_E_SHAPE) return 70;
    CHECK(dh_forest_to_json(forest, NULL, 0, &needed));
    if (needed < 100) return 71;

    if (argc >= 5) {
        const uint32_t w = (uint32_t)atoi(argv[3]), h = (uint32_t)atoi(argv[4]);
        size_t flen = 0;
        uint16_t* frame = (uint16_t*)slurp(argv[2], &flen);
        uint16_t* three;
        dh_result one, batch[3];
        const float guess[3] = {10.0f, -20.0f, 900.0f};
        int i;
        if (!frame || flen != (size_t)w * h * 2) return 72;
        CHECK(dh_forest_set_stepwidth(forest, 10u));
        CHECK(dh_forest_set_meanshift_iterations(forest, 20u));
        CHECK(dh_forest_set_sigma(forest, 8.0f));
        CHECK(dh_ctx_create(0, &ctx));
        CHECK(dh_predict(ctx, forest, frame, w, h, K, NULL, NULL, &one));
        printf("pose %.1f %.1f %.1f %.17g %.17g %.17g\n", (double)one.mid_point[0], (double)one.mid_point[1], (double)one.mid_point[2],
               one.rotation[0], one.rotation[1], one.rotation[2]);
        CHECK(dh_predict(ctx, forest, frame, w, h, K, guess, NULL, &one));
        printf("pose %.1f %.1f %.1f %.17g %.17g %.17g\n", (double)one.mid_point[0], (double)one.mid_point[1], (double)one.mid_point[2],
               one.rotation[0], one.rotation[1], one.rotation[2]);
        three = (uint16_t*)malloc(flen * 3);
        if (!three) return 73;
        for (i = 0; i < 3; ++i) memcpy((char*)three + (size_t)i * flen, frame, flen);
        CHECK(dh_predict_batch(ctx, forest, three, 3u, w, h, K, DH_DEPTH_HOST, batch));
        for (i = 0; i < 3; ++i)
            printf("pose %.1f %.1f %.1f %.17g %.17g %.17g\n", (double)batch[i].mid_point[0], (double)batch[i].mid_point[1],
                   (double)batch[i].mid_point[2], batch[i].rotation[0], batch[i].rotation[1], batch[i].rotation[2]);
        /* degenerate shapes are error codes, not undefined behaviour */
        if (dh_predict(ctx, forest, frame, 40u, 40u, K, NULL, NULL, &one) != DH_E_SHAPE) return 74;
        free(three);
        free(frame);
        dh_ctx_free(ctx);
    }
    dh_forest_free(forest);
    free(json);
    printf("ok\n");
    return 0;
}
