/* standalone_host.c — libmpt.so driven from plain C: no Python, no merian, no PyTorch.
 *
 * What a host that is not merian does to swap the path tracer in: loader output (float vertex attributes, scene-graph nodes) ->
 * mptt_build (include/mpt_tables.h: merian::Scene's grouping, GeometryData order and vertex path) -> mpt_scene_upload (BVH8 build on the
 * host cores) -> mptt_materials_build (the reference's material buffer: MaterialHeader {u16 type id, u16 alpha texture} + payload,
 * material_system.cpp:140-195) -> mpt_materials_upload -> mpt_env_set -> one frame into host memory (mpt_render_host) -> PFM (mpt_pfm_save, the Image Write node's format).
 *
 * Scene: a floor, a back wall and a box under a uniform sky; 160x120, 64 spp, max path length 5.
 *   gcc -std=c99 -Iinclude examples/standalone_host.c -Lmerian_b200/csrc -lmpt -Wl,-rpath,$PWD/merian_b200/csrc -lm -o standalone_host
 *   ./standalone_host out.pfm     (exit code 2 + the library's message when there is no sm_100 device: there is no CPU fallback)
 * tests/test_standalone_host.py builds and runs it, and compares the image with the same scene rendered through the Python mirror. */
#include <math.h> /* cosf, sinf for the node transform */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mpt.h"
#include "mpt_tables.h"

#define IMG_W 160
#define IMG_H 120

static int fail(const char* what, const char* msg) {
    fprintf(stderr, "%s: %s\n", what, msg ? msg : "?");
    return 1;
}

/* a quad (two triangles, indexed) with a constant normal */
static void quad(float* pos, float* nrm, const float a[3], const float b[3], const float c[3], const float d[3], const float n[3]) {
    const float* v[4] = {a, b, c, d};
    for (int i = 0; i < 4; i++)
        for (int k = 0; k < 3; k++) { pos[i * 3 + k] = v[i][k]; nrm[i * 3 + k] = n[k]; }
}

int main(int argc, char** argv) {
    const char* out_path = argc > 1 ? argv[1] : "standalone_host.pfm";
    mpt_context* ctx = NULL;
    if (mpt_create(0, &ctx) != MPT_OK) {
        fprintf(stderr, "mpt_create: %s\n", mpt_last_error(NULL));
        return 2;
    }

    /* ---- geometry: three meshes, every vertex attribute a plain float array */
    static float pos[3][24 * 3], nrm[3][24 * 3];
    static const uint16_t quad_idx[6] = {0, 1, 2, 0, 2, 3};
    static uint16_t box_idx[36];
    {
        const float f0[3] = {-2, 0, -2}, f1[3] = {2, 0, -2}, f2[3] = {2, 0, 2}, f3[3] = {-2, 0, 2}, up[3] = {0, 1, 0};
        quad(pos[0], nrm[0], f0, f3, f2, f1, up); /* floor, facing +y */
        const float w0[3] = {-2, 0, -2}, w1[3] = {2, 0, -2}, w2[3] = {2, 2.5f, -2}, w3[3] = {-2, 2.5f, -2}, front[3] = {0, 0, 1};
        quad(pos[1], nrm[1], w0, w1, w2, w3, front); /* back wall, facing +z */
        /* unit box centred on the origin (6 quads, outward normals); the node transform places it */
        const float h = 0.5f;
        const float c[8][3] = {{-h, -h, -h}, {h, -h, -h}, {h, h, -h}, {-h, h, -h}, {-h, -h, h}, {h, -h, h}, {h, h, h}, {-h, h, h}};
        const int face[6][4] = {{4, 5, 6, 7}, {1, 0, 3, 2}, {5, 1, 2, 6}, {0, 4, 7, 3}, {7, 6, 2, 3}, {0, 1, 5, 4}};
        const float fn[6][3] = {{0, 0, 1}, {0, 0, -1}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}};
        for (int f = 0; f < 6; f++) {
            quad(pos[2] + f * 12, nrm[2] + f * 12, c[face[f][0]], c[face[f][1]], c[face[f][2]], c[face[f][3]], fn[f]);
            for (int k = 0; k < 6; k++) box_idx[f * 6 + k] = (uint16_t)(f * 4 + quad_idx[k]);
        }
    }
    mpt_src_mesh meshes[3];
    memset(meshes, 0, sizeof(meshes));
    for (int m = 0; m < 3; m++) {
        meshes[m].positions = pos[m];
        meshes[m].normals = nrm[m];
        meshes[m].vertex_count = m == 2 ? 24 : 4;
        meshes[m].indices = m == 2 ? box_idx : quad_idx;
        meshes[m].primitive_count = m == 2 ? 12 : 2;
        meshes[m].index_size = 2;
        meshes[m].material_id = (uint32_t)m;
        meshes[m].flags = MPT_MESH_IS_OPAQUE;
        meshes[m].instance_mask = 0xFF;
    }
    const uint32_t n0[2] = {0, 1}, n1[1] = {2};
    mpt_src_node nodes[2];
    memset(nodes, 0, sizeof(nodes));
    const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    const float cs = cosf(0.5f), sn = sinf(0.5f);
    const float place[16] = {cs, 0, sn, 0.3f, 0, 1, 0, 0.5f, -sn, 0, cs, 0.2f, 0, 0, 0, 1}; /* rotate about y, lift onto the floor */
    memcpy(nodes[0].transform, ident, sizeof(ident));
    nodes[0].meshes = n0; nodes[0].mesh_count = 2;
    memcpy(nodes[1].transform, place, sizeof(place));
    nodes[1].meshes = n1; nodes[1].mesh_count = 1;

    mpt_tables* tables = NULL;
    if (mptt_build(meshes, 3, nodes, 2, &tables) != MPT_OK) return fail("mptt_build", mpt_last_error(NULL));
    if (mpt_scene_upload(ctx, mptt_scene_desc(tables)) != MPT_OK) return fail("mpt_scene_upload", mpt_last_error(ctx));

    /* ---- materials: the reference's entry layout (MaterialHeader + payload, material_system.cpp:140-195) from typed descriptions */
    mpt_src_material mats[3];
    memset(mats, 0, sizeof(mats));
    const float colours[3][4] = {{0.7f, 0.7f, 0.7f, 1.f}, {0.2f, 0.4f, 0.7f, 1.f}, {0.8f, 0.3f, 0.2f, 1.f}};
    for (int m = 0; m < 3; m++) {
        mats[m].model = MPT_MATERIAL_DIFFUSE;
        memcpy(mats[m].base_color_factor, colours[m], sizeof(colours[m]));
        mats[m].base_color_texture = MPT_TEXTURE_NONE;
    }
    mpt_material_table* mtab = NULL;
    mpt_material_table_view mv;
    if (mptt_materials_build(mats, 3, &mtab) != MPT_OK || mptt_materials_get(mtab, &mv) != MPT_OK) return fail("mptt_materials_build", mpt_last_error(NULL));
    if (mpt_materials_upload(ctx, mv.entries, mv.entry_size, mv.material_count, mv.types, mv.type_count) != MPT_OK)
        return fail("mpt_materials_upload", mpt_last_error(ctx));

    /* ---- a uniform sky: 8x4 lat-long texels */
    static float sky[8 * 4 * 4];
    for (int i = 0; i < 8 * 4; i++) { sky[i * 4] = 0.9f; sky[i * 4 + 1] = 1.0f; sky[i * 4 + 2] = 1.2f; sky[i * 4 + 3] = 1.f; }
    if (mpt_env_set(ctx, MPT_ENV_LATLONG, ident, 1.0f, 8, 4, sky) != MPT_OK) return fail("mpt_env_set", mpt_last_error(ctx));

    /* ---- camera: Camera::write_to (camera.cpp:349-368) */
    mpt_camera cam;
    const float eye[3] = {0.0f, 1.4f, 4.0f}, tgt[3] = {0.0f, 0.6f, 0.0f}, upv[3] = {0, 1, 0};
    if (mptt_camera_look_at(eye, tgt, upv, 45.0f, (float)IMG_W / (float)IMG_H, 0.01f, 1e4f, 0.0f, NULL, &cam) != MPT_OK)
        return fail("mptt_camera_look_at", mpt_last_error(NULL));

    /* ---- one frame into host memory (mpt_render_host: H2D of nothing here, render, D2H of the irradiance image). A host that keeps device
     * memory drives the node object instead (include/mpt_node.h: mptn_load_config with the reference's property keys, mptn_process). */
    mpt_render_params P;
    memset(&P, 0, sizeof(P));
    P.width = IMG_W; P.height = IMG_H; P.frame = 0; P.spp = 64; P.max_path_length = 5; P.instance_mask = 0xFF;
    P.flags = MPT_RENDER_EMISSION_ON_PRIMARY; P.sample_offset = 0; P.sample_stride = 1;
    float* image = (float*)malloc((size_t)IMG_W * IMG_H * 4 * sizeof(float));
    if (!image) return fail("malloc", "out of memory");
    if (mpt_render_host(ctx, &cam, &P, NULL, NULL, image, NULL) != MPT_OK) return fail("mpt_render_host", mpt_last_error(ctx));

    double lum = 0.0;
    for (int i = 0; i < IMG_W * IMG_H; i++) lum += 0.299 * image[i * 4] + 0.587 * image[i * 4 + 1] + 0.114 * image[i * 4 + 2];
    printf("mean luminance %.6f\n", lum / (IMG_W * IMG_H));
    if (mpt_pfm_save(out_path, image, IMG_W, IMG_H, 4) != MPT_OK) return fail("mpt_pfm_save", mpt_last_error(NULL));
    printf("wrote %s\n", out_path);
    free(image);
    mptt_materials_free(mtab);
    mptt_free(tables);
    mpt_destroy(ctx);
    return 0;
}
