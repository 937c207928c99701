/* Headless use of the C ABI from plain C (what a cgo / JNI / Rust FFI binding would call):
 *   compile the rule set, create a grid, paint a circle of sand, step, read back the census.
 *
 *   gcc -std=c99 -Wall -Wextra -pedantic -I include examples/headless_step.c \
 *       -L sandengine_b200 -lsandengine_b200 -Wl,-rpath,$PWD/sandengine_b200 -o headless_step
 *   ./headless_step data/materials.yaml [width height steps]
 *
 * Exit codes: 0 ok, 2 usage / IO, 3 the library reported an error (e.g. no CUDA device: there is no CPU fallback).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sandengine_b200.h"

static int die(const char* what, int code) {
    fprintf(stderr, "%s failed (%d): %s\n", what, code, se_last_error());
    return 3;
}

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s materials.yaml [width height steps]\n", argv[0]);
        return 2;
    }
    const uint32_t width = argc > 2 ? (uint32_t)atoi(argv[2]) : 512u;
    const uint32_t height = argc > 3 ? (uint32_t)atoi(argv[3]) : 512u;
    const uint32_t steps = argc > 4 ? (uint32_t)atoi(argv[4]) : 100u;

    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    char* yaml = (char*)malloc((size_t)n + 1);
    if (!yaml || fread(yaml, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(yaml); return 2; }
    fclose(f);

    printf("%s\n", se_version());
    se_rules* rules = NULL;
    int rc = se_rules_compile_yaml(yaml, (size_t)n, &rules);      /* parse + CUDA codegen + NVRTC (no device needed) */
    free(yaml);
    if (rc != SE_OK) return die("se_rules_compile_yaml", rc);

    int32_t n_rules = 0, n_types = 0, n_materials = 0, sand = -1;
    se_rules_counts(rules, &n_rules, &n_types, &n_materials);
    printf("rule set: %d rules, %d types, %d materials\n", (int)n_rules, (int)n_types, (int)n_materials);
    if ((rc = se_rules_material_id(rules, "sand", &sand)) != SE_OK) { se_rules_destroy(rules); return die("se_rules_material_id", rc); }

    se_create_params prm;
    memset(&prm, 0, sizeof prm);
    prm.width = width;
    prm.height = height;
    se_sim* sim = NULL;
    if ((rc = se_sim_create(rules, &prm, &sim)) != SE_OK) { se_rules_destroy(rules); return die("se_sim_create", rc); }

    /* Simulation::run() x 2: frame 1 clears the grid (falling_sand.glsl:743-746), frame 2 takes the brush stamp */
    if ((rc = se_sim_step(sim, 1)) != SE_OK) goto fail;
    {
        se_modification m;
        memset(&m, 0, sizeof m);
        m.position[0] = (int32_t)(width / 2);
        m.position[1] = (int32_t)(height / 4);
        m.mod_shape = SE_MODSHAPE_CIRCLE;
        m.mod_size = 20;
        m.mod_matID = sand;
        if ((rc = se_sim_push_modifications(sim, &m, 1)) != SE_OK) goto fail;
    }
    if ((rc = se_sim_step(sim, 1)) != SE_OK) goto fail;
    if ((rc = se_sim_step(sim, steps)) != SE_OK) goto fail;       /* a run of plain steps: fused on the device */
    {
        uint64_t census[256];
        int32_t frame = 0;
        if ((rc = se_sim_census(sim, census)) != SE_OK) goto fail;
        se_sim_get_frame(sim, &frame);
        printf("frame %d: %llu sand cells, %llu empty cells\n", (int)frame, (unsigned long long)census[sand], (unsigned long long)census[0]);
    }
    se_sim_destroy(sim);
    se_rules_destroy(rules);
    return 0;
fail:
    die("simulation", rc);
    se_sim_destroy(sim);
    se_rules_destroy(rules);
    return 3;
}
