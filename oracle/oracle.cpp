// oracle.cpp — TEST INFRASTRUCTURE ONLY. Never linked into, imported by, or executed from the product
// (merian_b200/, libmpt.so). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load liboracle.so.
//
// CPU restatement (C++17 + OpenMP) of the path-tracing hot path behind merian-graph's `pt` node:
//   render_pt.slang (all), scene.slang, hit.slang, camera.slang, environment-map.slang,
//   material-system.slang, gltf-material.slang, diffuse-material.slang, gbuffer.slang (hit packing).
// The reference's device code is Slang compiled to SPIR-V at run time and its ray traversal lives in
// the Vulkan driver (acceleration-structure.slang:59-72); neither can be compiled or run in this
// container (no Vulkan, no slang, no meson — SURVEY.md §0.8), so oracle/_ref does not exist.
//
// PARITY UNPINNED for visibility and radiance: the reference's tests hold no golden vectors for this
// path; traversal/intersection follow the Vulkan specification (closest hit,
// back-face culling in object space, 8-bit instance masks, opaque / non-opaque candidates), implemented
// here with a median-split BVH2 and the watertight test of Woop, Benthin, Wald (JCGT 2013).
// The BSDF layer IS pinned (tests/test_bsdf.cpp restated in tests/test_bsdf_oracle.py).
#include "../include/mpt.h"
#include "oracle_bsdf.hpp"

#include <array>
#include <atomic>
#include <cstdio>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

// ------------------------------------------------------------------------------------------------
// BVH2 over generic boxes (oracle-only acceleration; not the product's BVH8)
struct AABB {
    float3 lo{1e30f, 1e30f, 1e30f}, hi{-1e30f, -1e30f, -1e30f};
    void grow(float3 p) {
        lo = float3(std::min(lo.x, p.x), std::min(lo.y, p.y), std::min(lo.z, p.z));
        hi = float3(std::max(hi.x, p.x), std::max(hi.y, p.y), std::max(hi.z, p.z));
    }
    void grow(const AABB& b) { grow(b.lo); grow(b.hi); }
};

struct Bvh2 {
    struct Node {
        AABB box;
        uint32_t left;  // internal: index of left child (right = left + 1); leaf: first prim slot
        uint32_t count; // 0 = internal, else leaf prim count
    };
    std::vector<Node> nodes;
    std::vector<uint32_t> prim_index;

    void build(const std::vector<AABB>& boxes) {
        const uint32_t n = uint32_t(boxes.size());
        prim_index.resize(n);
        std::iota(prim_index.begin(), prim_index.end(), 0u);
        nodes.clear();
        nodes.reserve(size_t(n) * 2 + 1);
        if (n == 0) return;
        std::vector<float3> cent(n);
        for (uint32_t i = 0; i < n; i++) cent[i] = (boxes[i].lo + boxes[i].hi) * 0.5f;
        nodes.push_back({});
        struct Job { uint32_t node, first, count; };
        std::vector<Job> stack{{0, 0, n}};
        while (!stack.empty()) {
            const Job j = stack.back();
            stack.pop_back();
            AABB box, cbox;
            for (uint32_t i = j.first; i < j.first + j.count; i++) {
                box.grow(boxes[prim_index[i]]);
                cbox.grow(cent[prim_index[i]]);
            }
            nodes[j.node].box = box;
            const float3 ext = cbox.hi - cbox.lo;
            int axis = ext.x > ext.y ? (ext.x > ext.z ? 0 : 2) : (ext.y > ext.z ? 1 : 2);
            if (j.count <= 4 || ext[axis] <= 0.f) {
                nodes[j.node].left = j.first;
                nodes[j.node].count = j.count;
                continue;
            }
            const uint32_t mid = j.count / 2;
            std::nth_element(prim_index.begin() + j.first, prim_index.begin() + j.first + mid,
                             prim_index.begin() + j.first + j.count,
                             [&](uint32_t a, uint32_t b) { return cent[a][axis] < cent[b][axis]; });
            const uint32_t l = uint32_t(nodes.size());
            nodes.push_back({});
            nodes.push_back({});
            nodes[j.node].left = l;
            nodes[j.node].count = 0;
            stack.push_back({l, j.first, mid});
            stack.push_back({l + 1, j.first + mid, j.count - mid});
        }
    }
};

// conservative slab test against [tmin, tmax]
inline bool ray_box(const AABB& b, float3 o, float3 inv_d, float tmin, float tmax) {
    float t0 = tmin, t1 = tmax;
    for (int a = 0; a < 3; a++) {
        float tn = (b.lo[a] - o[a]) * inv_d[a];
        float tf = (b.hi[a] - o[a]) * inv_d[a];
        if (tn > tf) std::swap(tn, tf);
        tf *= 1.0000005f; // 1 + 4 ulp: keeps the traversal conservative w.r.t. the triangle test
        tn *= 0.9999995f;
        if (std::isnan(tn) || std::isnan(tf)) continue; // 0 * inf: ray lies in the slab plane
        t0 = std::max(t0, std::min(tn, tf));
        t1 = std::min(t1, std::max(tn, tf));
    }
    return t0 <= t1;
}

// ------------------------------------------------------------------------------------------------
// Watertight ray/triangle intersection (Woop, Benthin, Wald 2013), fp32 with double fallback on zero
// edge functions. Compiled with -ffp-contract=off: every multiply/add below is a separately rounded
// IEEE operation, which merian_b200/csrc/intersect.cuh reproduces with __fmul_rn/__fsub_rn/__fadd_rn.
struct RayShear {
    int kx, ky, kz;
    float Sx, Sy, Sz;
    explicit RayShear(float3 d) {
        const float ax = std::fabs(d.x), ay = std::fabs(d.y), az = std::fabs(d.z);
        kz = (ax > ay) ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
        kx = kz + 1; if (kx == 3) kx = 0;
        ky = kx + 1; if (ky == 3) ky = 0;
        if (d[kz] < 0.0f) std::swap(kx, ky);
        Sx = d[kx] / d[kz];
        Sy = d[ky] / d[kz];
        Sz = 1.0f / d[kz];
    }
};

enum CullMode { CULL_BACK = 0, CULL_FRONT = 1, CULL_NONE = 2 }; // CULL_FRONT = back-face culling with flipped facing

// Returns true and (t, u, v) if the triangle is hit with tmin < t < tmax. u, v = weights of v1, v2
// (RayQuery::CommittedRayBarycentrics convention, hit.slang:58-60).
inline bool intersect_triangle(const RayShear& rs, float3 org, float3 v0, float3 v1, float3 v2, int cull,
                               float tmin, float tmax, float& t_out, float& u_out, float& v_out) {
    const float3 A = v0 - org, B = v1 - org, C = v2 - org;
    const float Ax = A[rs.kx] - rs.Sx * A[rs.kz], Ay = A[rs.ky] - rs.Sy * A[rs.kz];
    const float Bx = B[rs.kx] - rs.Sx * B[rs.kz], By = B[rs.ky] - rs.Sy * B[rs.kz];
    const float Cx = C[rs.kx] - rs.Sx * C[rs.kz], Cy = C[rs.ky] - rs.Sy * C[rs.kz];
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        U = float(double(Cx) * double(By) - double(Cy) * double(Bx));
        V = float(double(Ax) * double(Cy) - double(Ay) * double(Cx));
        W = float(double(Bx) * double(Ay) - double(By) * double(Ax));
    }
    // front-facing (counter-clockwise seen from the ray origin, normal = cross(e01, e02) towards the
    // origin) <=> U + V + W > 0; see DESIGN.md "facing".
    const bool any_neg = (U < 0.0f) || (V < 0.0f) || (W < 0.0f);
    const bool any_pos = (U > 0.0f) || (V > 0.0f) || (W > 0.0f);
    if (cull == CULL_BACK) { if (any_neg) return false; }
    else if (cull == CULL_FRONT) { if (any_pos) return false; }
    else { if (any_neg && any_pos) return false; }
    const float det = U + V + W;
    if (det == 0.0f) return false;
    const float Az = rs.Sz * A[rs.kz], Bz = rs.Sz * B[rs.kz], Cz = rs.Sz * C[rs.kz];
    const float T = U * Az + V * Bz + W * Cz;
    const float rcp = 1.0f / det;
    const float t = T * rcp;
    if (!(t > tmin && t < tmax)) return false;
    t_out = t;
    u_out = V * rcp;
    v_out = W * rcp;
    return true;
}

// ------------------------------------------------------------------------------------------------
// Scene tables
struct Texture {
    uint32_t w = 0, h = 0, format = 0, wrap_u = 0, wrap_v = 0, filter = 1;
    uint32_t mipmap_mode = MPT_MIPMAP_NONE; // VkSamplerMipmapMode of a texture that has a chain
    std::vector<uint8_t> rgba8;
    std::vector<float> rgba32f;
    std::vector<Texture> mips; // levels 1 .. (level 0 is this texture)
    bool valid() const { return w > 0; }
    const Texture& level(uint32_t l) const { return l == 0 ? *this : mips[l - 1]; }
    uint32_t levels() const { return uint32_t(mips.size()) + 1; }
};

inline float srgb_to_linear(float c) { // Vulkan / Khronos data format spec sRGB EOTF
    return c <= 0.04045f ? c / 12.92f : std::pow((c + 0.055f) / 1.055f, 2.4f);
}

inline float linear_to_srgb(float l) { return l <= 0.0031308f ? 12.92f * l : 1.055f * std::pow(l, 1.0f / 2.4f) - 0.055f; }

inline int wrap_coord(int i, int n, uint32_t mode) {
    if (mode == MPT_WRAP_CLAMP_TO_EDGE) return std::min(std::max(i, 0), n - 1);
    if (mode == MPT_WRAP_MIRRORED_REPEAT) {
        int m = i % (2 * n);
        if (m < 0) m += 2 * n;
        return m < n ? m : 2 * n - 1 - m;
    }
    int m = i % n;
    return m < 0 ? m + n : m;
}

inline float4 texel(const Texture& t, int x, int y) {
    const size_t o = (size_t(y) * t.w + x) * 4;
    if (t.format == MPT_TEX_RGBA32F) return float4(t.rgba32f[o], t.rgba32f[o + 1], t.rgba32f[o + 2], t.rgba32f[o + 3]);
    float4 c(t.rgba8[o] / 255.0f, t.rgba8[o + 1] / 255.0f, t.rgba8[o + 2] / 255.0f, t.rgba8[o + 3] / 255.0f);
    if (t.format == MPT_TEX_RGBA8_SRGB) { c.x = srgb_to_linear(c.x); c.y = srgb_to_linear(c.y); c.z = srgb_to_linear(c.z); }
    return c;
}

// LOD0Sampler::sample (textures.slang:338-366) = SampleLevel(uv, 0): Vulkan spec texel filtering.
inline float4 sample_texture(const Texture& t, float2 uv) {
    const float u = uv.x * float(t.w), v = uv.y * float(t.h);
    if (t.filter == MPT_FILTER_NEAREST) {
        return texel(t, wrap_coord(int(std::floor(u)), int(t.w), t.wrap_u), wrap_coord(int(std::floor(v)), int(t.h), t.wrap_v));
    }
    const float fu = u - 0.5f, fv = v - 0.5f;
    const int i0 = int(std::floor(fu)), j0 = int(std::floor(fv));
    const float a = fu - std::floor(fu), b = fv - std::floor(fv);
    const int x0 = wrap_coord(i0, int(t.w), t.wrap_u), x1 = wrap_coord(i0 + 1, int(t.w), t.wrap_u);
    const int y0 = wrap_coord(j0, int(t.h), t.wrap_v), y1 = wrap_coord(j0 + 1, int(t.h), t.wrap_v);
    const float4 t00 = texel(t, x0, y0), t10 = texel(t, x1, y0), t01 = texel(t, x0, y1), t11 = texel(t, x1, y1);
    return ((1 - a) * (1 - b)) * t00 + (a * (1 - b)) * t10 + ((1 - a) * b) * t01 + (a * b) * t11;
}

// Mip chain of create_texture_from_rgba8 / _rgba32f(generate_mipmaps) (src/merian/vk/memory/resource_allocator.cpp:298-358):
// floor(log2(max(w, h))) + 1 levels, level i = vkCmdBlitImage(level i - 1 -> max(1, dim >> i), VK_FILTER_LINEAR)
// (include/merian/vk/utils/blits.hpp:101-168). Blits convert through the format: sRGB texels are filtered in linear space and re-encoded.
inline void generate_mipmaps(Texture& t) {
    t.mips.clear();
    uint32_t n = 1;
    for (uint32_t m = std::max(t.w, t.h); m > 1; m >>= 1) n++;
    const Texture* src = &t;
    t.mips.reserve(n);
    for (uint32_t i = 1; i < n; i++) {
        Texture d;
        d.w = std::max(1u, t.w >> i); d.h = std::max(1u, t.h >> i);
        d.format = t.format; d.wrap_u = t.wrap_u; d.wrap_v = t.wrap_v; d.filter = t.filter;
        const bool f32 = t.format == MPT_TEX_RGBA32F;
        if (f32) d.rgba32f.resize(size_t(d.w) * d.h * 4); else d.rgba8.resize(size_t(d.w) * d.h * 4);
        const float sx = float(src->w) / float(d.w), sy = float(src->h) / float(d.h);
        for (uint32_t y = 0; y < d.h; y++)
            for (uint32_t x = 0; x < d.w; x++) {
                const float u = (float(x) + 0.5f) * sx - 0.5f, v = (float(y) + 0.5f) * sy - 0.5f;
                const float fu = std::floor(u), fv = std::floor(v), a = u - fu, b = v - fv;
                auto at = [&](int xx, int yy) { // clamp to edge; texel() decodes sRGB
                    return texel(*src, std::min(std::max(xx, 0), int(src->w) - 1), std::min(std::max(yy, 0), int(src->h) - 1));
                };
                const float4 r = ((1 - a) * (1 - b)) * at(int(fu), int(fv)) + (a * (1 - b)) * at(int(fu) + 1, int(fv)) +
                                 ((1 - a) * b) * at(int(fu), int(fv) + 1) + (a * b) * at(int(fu) + 1, int(fv) + 1);
                const size_t o = (size_t(y) * d.w + x) * 4;
                if (f32) {
                    d.rgba32f[o] = r.x; d.rgba32f[o + 1] = r.y; d.rgba32f[o + 2] = r.z; d.rgba32f[o + 3] = r.w;
                } else {
                    const bool srgb = t.format == MPT_TEX_RGBA8_SRGB;
                    const float e[4] = {srgb ? linear_to_srgb(r.x) : r.x, srgb ? linear_to_srgb(r.y) : r.y, srgb ? linear_to_srgb(r.z) : r.z, r.w};
                    for (int c = 0; c < 4; c++) d.rgba8[o + size_t(c)] = uint8_t(std::min(std::max(e[c] * 255.0f + 0.5f, 0.0f), 255.0f));
                }
            }
        t.mips.push_back(std::move(d));
        src = &t.mips.back();
    }
}

// GradLODSampler::sample (textures.slang:398-421) = SampleGrad(uv, ddx, ddy): LOD from the gradients as the Vulkan specification defines
// it for a sampler without anisotropy ("Scale Factor Operation", "LOD Operation", "Image Level(s) Selection"): rho = the longer of the two
// gradient vectors in texel units, lambda = log2(rho); lambda <= 0 magnifies (level 0); otherwise the level nearest to lambda
// (VK_SAMPLER_MIPMAP_MODE_NEAREST) or a blend of the two levels around it (LINEAR). The hardware may approximate rho and quantises
// lambda to 8 fractional bits; tests compare with a tolerance.
inline float4 sample_texture_grad(const Texture& t, float2 uv, float2 ddx, float2 ddy) {
    if (t.mips.empty()) return sample_texture(t, uv);
    const float rx = std::sqrt(ddx.x * float(t.w) * ddx.x * float(t.w) + ddx.y * float(t.h) * ddx.y * float(t.h));
    const float ry = std::sqrt(ddy.x * float(t.w) * ddy.x * float(t.w) + ddy.y * float(t.h) * ddy.y * float(t.h));
    const float rho = std::max(rx, ry);
    const float q = float(t.levels() - 1);
    if (!(rho > 1.0f)) return sample_texture(t, uv); // lambda <= 0 (or NaN gradients): magnification
    const float lambda = std::min(std::log2(rho), q);
    if (t.mipmap_mode == MPT_MIPMAP_NEAREST) {
        const float d = std::min(std::max(std::ceil(lambda + 0.5f) - 1.0f, 0.0f), q);
        return sample_texture(t.level(uint32_t(d)), uv);
    }
    const float dh = std::floor(lambda), delta = lambda - dh;
    const uint32_t hi = uint32_t(dh), lo = std::min(hi + 1u, t.levels() - 1u);
    return (1.0f - delta) * sample_texture(t.level(hi), uv) + delta * sample_texture(t.level(lo), uv);
}

struct Mesh {
    std::vector<mpt_packed_vertex> vertices;
    std::vector<uint32_t> indices; // widened to u32; empty for IndexType::None
    uint32_t primitive_count = 0;
    uint32_t index_flag = 0;
    bool opaque = true;
    void tri(uint32_t prim, uint32_t& a, uint32_t& b, uint32_t& c) const { // scene.slang:158-173
        if (index_flag == 0) { a = prim * 3; b = a + 1; c = a + 2; }
        else { a = indices[size_t(prim) * 3]; b = indices[size_t(prim) * 3 + 1]; c = indices[size_t(prim) * 3 + 2]; }
    }
};

struct Blas {
    std::vector<uint32_t> meshes;
    struct Prim { uint32_t geometry_index, primitive_id; };
    std::vector<Prim> prims;
    Bvh2 bvh;
    AABB bounds;
};

struct Hit { // hit.slang:25-66
    uint32_t instance_id = ~0u, instance_index = ~0u, primitive_id = ~0u, geometry_index = ~0u;
    float2 bary{0, 0};
    bool valid() const { return instance_id != ~0u; }
    float3 bary3() const { return float3(1.0f - bary.x - bary.y, bary.x, bary.y); }
};

struct Ray { float3 o; float tmin; float3 d; float tmax; };

struct Medium { // HomogeneousVolume: Vacuum (homogeneous-volume.slang:29-50) or IsotropicVolume (:81-112)
    bool vacuum = true;
    float3 mu_t{0, 0, 0};
    float eta = 1.0f;
    float3 transmittance(float t) const { return vacuum ? float3(1.0f) : exp3(-t * mu_t); }
};

// MaterialSample (material.slang:50-66) for both supported models
struct MaterialSample {
    bool is_gltf = false;
    std::optional<LambertDiffuseBRDF> diffuse;
    std::optional<GltfBSDF> gltf;
    float eta = 1.0f;
    float3 absorption{0, 0, 0};
    OptSample sample_eval(float3 wi, Rng& rng) const { return is_gltf ? gltf->sample_eval(wi, rng) : diffuse->sample_eval(wi, rng); }
    float3 eval(float3 wi, float3 wo, Rng& rng) const { return is_gltf ? gltf->eval(wi, wo, rng) : diffuse->eval(wi, wo, rng); }
    float pdf(float3 wi, float3 wo) const { return is_gltf ? gltf->pdf(wi, wo) : diffuse->pdf(wi, wo); }
    float3 get_albedo(float3 wi) const { return is_gltf ? gltf->get_albedo(wi) : diffuse->get_albedo(wi); }
    Medium get_interior_volume() const { // material.slang:62-65 / gltf-material.slang:322-324
        Medium m;
        m.vacuum = false;
        m.mu_t = is_gltf ? absorption : float3(0.f);
        m.eta = is_gltf ? eta : 1.0f;
        return m;
    }
};

struct ShadingData { // shading-data.slang:27-58
    float3 wi, pos;
    float2 uv;
    Frame frame;
    float3 face_normal;
    uint16_t bounce;
    float ior;
    bool front_facing;
    // the LODSampler2D handed to get_material_sample: LOD0Sampler (pt, textures.slang:338-366) or GradLODSampler(ddx, ddy) (GBuffer, :398-421)
    bool use_grad = false;
    float2 ddx{0, 0}, ddy{0, 0};
};

struct ShadingPoint { // hit.slang:105-147
    float3 pos;
    uint32_t enc_face_normal = 0;
    float3 emission{0, 0, 0};
    Frame shading_frame;
    bool has_surface_material = false;
    MaterialSample surface_material;
    float3 get_face_normal() const { return decode_normal(enc_face_normal); }
};

// hit.slang:14-23 (Waechter & Binder)
inline float3 offset_ray_origin(float3 pos, float3 n) {
    const float origin = 1.f / 16.f, float_scale = 1.f / 65536.f, int_scale = 256.f;
    const int ox = int(n.x * int_scale), oy = int(n.y * int_scale), oz = int(n.z * int_scale);
    const float ix = as_float_i(as_int(pos.x) + (pos.x < 0.f ? -ox : ox));
    const float iy = as_float_i(as_int(pos.y) + (pos.y < 0.f ? -oy : oy));
    const float iz = as_float_i(as_int(pos.z) + (pos.z < 0.f ? -oz : oz));
    return float3(std::fabs(pos.x) < origin ? pos.x + n.x * float_scale : ix,
                  std::fabs(pos.y) < origin ? pos.y + n.y * float_scale : iy,
                  std::fabs(pos.z) < origin ? pos.z + n.z * float_scale : iz);
}

constexpr float SCENE_RAY_TMIN = 0.f, SCENE_RAY_TMAX = 1e7f; // hit.slang:7-8
constexpr float ENV_MAP_DISTANCE = 1e7f;                      // environment-map.slang:11
constexpr float SCENE_THROUGHPUT_THRESHOLD = 1e-5f;           // scene.slang:24
constexpr float DEFAULT_MIN_ROUGHNESS = 0.0316f;              // material.slang:23
constexpr float DEFAULT_ALPHA_TEST_THRESHOLD = 0.5f;          // material-system.slang:7

struct Scene {
    std::vector<Mesh> meshes;
    std::vector<Blas> blas;
    std::vector<mpt_geometry_desc> geometries;
    std::vector<mpt_instance_desc> instances;
    std::vector<float> instance_transforms, inv_t_instance_transforms;
    std::vector<float> prev_instance_transforms; // Scene::prev_instance_transforms (scene.slang:38); empty = static scene
    std::vector<AABB> instance_bounds;
    std::vector<std::array<float, 12>> instance_inverse; // world -> object
    Bvh2 tlas;

    std::vector<uint8_t> material_entries;
    uint32_t material_entry_size = 0, material_count = 0;
    std::vector<mpt_material_type> material_types;
    std::vector<Texture> textures;
    // MaterialSystem hints (material_system.hpp:83-97): compile-time constants of the reference's composition
    float MIN_ROUGHNESS = DEFAULT_MIN_ROUGHNESS, ALPHA_TEST_THRESHOLD = DEFAULT_ALPHA_TEST_THRESHOLD;
    bool clamp_normals = true;

    uint32_t env_type = MPT_ENV_EMPTY;
    float env_to_local[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    float env_intensity = 1.f;
    Texture env_tex;

    std::atomic<uint64_t> rays_traced{0};

    // ---------------------------------------------------------------- materials
    const uint8_t* entry(uint32_t material_id) const { return material_entries.data() + size_t(material_id) * material_entry_size; }
    static uint16_t rd16(const uint8_t* p) { uint16_t v; std::memcpy(&v, p, 2); return v; }
    static float rdf(const uint8_t* p) { float v; std::memcpy(&v, p, 4); return v; }
    static float3 rdf3(const uint8_t* p) { return float3(rdf(p), rdf(p + 4), rdf(p + 8)); }
    float4 tex(uint16_t id, float2 uv) const { return sample_texture(textures[id], uv); }
    float4 tex(uint16_t id, const ShadingData& sd) const {
        return sd.use_grad ? sample_texture_grad(textures[id], sd.uv, sd.ddx, sd.ddy) : sample_texture(textures[id], sd.uv);
    }

    // material-system.slang:28-38
    bool alpha_test(uint32_t material_id, float2 uv) const {
        const uint16_t alpha_tex = rd16(entry(material_id) + 2);
        if (alpha_tex != MPT_TEXTURE_NONE) return tex(alpha_tex, uv).w < ALPHA_TEST_THRESHOLD;
        return false;
    }

    // MaterialSystem::get_material_sample (material-system.slang:64-72) -> DiffuseMaterial::get_sample
    // (diffuse-material.slang:19-34) or GltfMaterial::get_sample (gltf-material.slang:153-308).
    // emission_only mirrors MaterialModel::get_emission (material.slang:40-47, gltf-material.slang:142-151).
    MaterialSample get_material_sample(uint32_t material_id, const ShadingData& sd, float3& out_emission,
                                       Frame& out_frame, bool emission_only = false) const {
        const uint8_t* e = entry(material_id);
        const uint16_t type_id = rd16(e), alpha_tex = rd16(e + 2);
        const mpt_material_type mt = material_types[type_id];
        const uint8_t* p = e + 4;
        MaterialSample s;
        if (mt.model == MPT_MATERIAL_DIFFUSE) {
            out_emission = float3(0.f);
            if (emission_only) return s;
            out_frame = sd.frame.apply_normal_map(float3(0, 0, 1));
            float3 albedo = rdf3(p);
            if (alpha_tex != MPT_TEXTURE_NONE) albedo = albedo * tex(alpha_tex, sd).xyz();
            s.is_gltf = false;
            s.diffuse = LambertDiffuseBRDF(albedo);
            return s;
        }
        // ---- glTF payload, gltf_material.hpp:102-146
        float4 base(rdf(p), rdf(p + 4), rdf(p + 8), rdf(p + 12));
        const float3 emissive_factor = rdf3(p + 16);
        const float metallic_factor = rdf(p + 28), roughness_factor = rdf(p + 32), normal_scale = rdf(p + 36);
        const float ior = rdf(p + 44);
        const uint16_t mr_tex = rd16(p + 48), normal_tex = rd16(p + 50), emissive_tex = rd16(p + 54);
        const uint8_t* q = p + 56;
        const bool T = mt.features & MPT_GLTF_TRANSMISSION, V = mt.features & MPT_GLTF_VOLUME,
                   C = mt.features & MPT_GLTF_CLEARCOAT, S = mt.features & MPT_GLTF_SHEEN,
                   I = mt.features & MPT_GLTF_IRIDESCENCE, A = mt.features & MPT_GLTF_ANISOTROPY;
        float tw = 0.f;
        if (T) { tw = rdf(q); q += 4; }
        float3 vol_abs(0.f);
        if (V) { vol_abs = rdf3(q); q += 12; }
        float cc_w = 0, cc_r = 0; uint16_t cc_tex = MPT_TEXTURE_NONE, cc_rtex = MPT_TEXTURE_NONE;
        if (C) { cc_w = rdf(q); cc_r = rdf(q + 4); cc_tex = rd16(q + 8); cc_rtex = rd16(q + 10); q += 12; }
        float3 sh_col(0.f); float sh_r = 0; uint16_t sh_ctex = MPT_TEXTURE_NONE, sh_rtex = MPT_TEXTURE_NONE;
        if (S) { sh_col = rdf3(q); sh_r = rdf(q + 12); sh_ctex = rd16(q + 16); sh_rtex = rd16(q + 18); q += 20; }
        float ir_f = 0, ir_ior = 1.3f, ir_tmin = 100, ir_tmax = 400; uint16_t ir_tex = MPT_TEXTURE_NONE, ir_ttex = MPT_TEXTURE_NONE;
        if (I) { ir_f = rdf(q); ir_ior = rdf(q + 4); ir_tmin = rdf(q + 8); ir_tmax = rdf(q + 12); ir_tex = rd16(q + 16); ir_ttex = rd16(q + 18); q += 20; }
        float an_s = 0, an_rot = 0; uint16_t an_tex = MPT_TEXTURE_NONE;
        if (A) { an_s = rdf(q); an_rot = rdf(q + 4); an_tex = rd16(q + 8); q += 10; }

        float3 emission = emissive_factor;
        if (emissive_tex != MPT_TEXTURE_NONE) emission = emission * tex(emissive_tex, sd).xyz();
        out_emission = emission;
        if (emission_only) return s;

        if (alpha_tex != MPT_TEXTURE_NONE) base = base * tex(alpha_tex, sd);
        float metallic = metallic_factor, roughness = roughness_factor;
        if (mr_tex != MPT_TEXTURE_NONE) {
            const float4 mr = tex(mr_tex, sd);
            roughness *= mr.y;
            metallic *= mr.z;
        }
        Frame shading_frame = sd.frame;
        if (normal_tex != MPT_TEXTURE_NONE) {
            const float4 raw = tex(normal_tex, sd);
            const float2 nxy((raw.x * 2.0f - 1.0f) * normal_scale, (raw.y * 2.0f - 1.0f) * normal_scale);
            const float nz = std::sqrt(std::max(0.0f, 1.0f - dot(nxy, nxy)));
            const Frame mapped = sd.frame.apply_normal_map(float3(nxy.x, nxy.y, nz));
            // merian_hint_clamp_normals = true (material.slang:20)
            const float3 clamped_n = clamp_normals ? keller_clamp_normal(mapped.n, sd.face_normal, sd.wi) : mapped.n;
            shading_frame = Frame(clamped_n, mapped.t);
        }
        const float alpha = pow2(std::max(roughness, MIN_ROUGHNESS));
        float2 specular_alpha(alpha);
        if (A) {
            float strength = an_s;
            float2 dir(std::cos(an_rot), std::sin(an_rot));
            if (an_tex != MPT_TEXTURE_NONE) {
                const float4 a = tex(an_tex, sd);
                const float2 d = normalize(float2(a.x * 2.0f - 1.0f, a.y * 2.0f - 1.0f));
                dir = float2(dir.x * d.x - dir.y * d.y, dir.x * d.y + dir.y * d.x);
                strength *= a.z;
            }
            if (strength > 0.0f) {
                specular_alpha.x = lerp(alpha, 1.0f, strength * strength);
                const float3 t = dir.x * shading_frame.t + dir.y * shading_frame.b;
                shading_frame = Frame(shading_frame.n, t);
            }
        }
        const float ior_r = (ior - 1.0f) / (ior + 1.0f);
        const float dielectric_f0 = ior_r * ior_r;
        const float eta = sd.ior / (sd.front_facing ? ior : 1.0f);
        float iridescence_factor = 0.f, iri_ior = 1.0f, iri_thickness = 0.f;
        if (I) {
            iridescence_factor = ir_f;
            if (ir_tex != MPT_TEXTURE_NONE) iridescence_factor *= tex(ir_tex, sd).x;
            iri_ior = ir_ior;
            float t = 1.0f;
            if (ir_ttex != MPT_TEXTURE_NONE) t = tex(ir_ttex, sd).y;
            iri_thickness = lerp(ir_tmin, ir_tmax, t);
        }
        const float3 base_rgb = base.xyz();
        const GltfOpaqueDielectric opaque_dielectric(SpecularBRDF(specular_alpha), LambertDiffuseBRDF(base_rgb), dielectric_f0,
                                                     eta, iri_ior, iri_thickness, iridescence_factor);
        const GltfDielectric dielectric(opaque_dielectric,
                                        IridescentRoughDielectricBSDF(alpha, eta, base_rgb, float3(dielectric_f0), iri_ior,
                                                                      iri_thickness, iridescence_factor),
                                        tw);
        const GltfReflective reflective(dielectric,
                                        GltfConductor(SpecularBRDF(specular_alpha), base_rgb, iri_ior, iri_thickness, iridescence_factor),
                                        metallic);
        float coat_weight = 0.f, coat_alpha = pow2(MIN_ROUGHNESS);
        if (C) {
            coat_weight = cc_w;
            if (cc_tex != MPT_TEXTURE_NONE) coat_weight *= tex(cc_tex, sd).x;
            float coat_rough = cc_r;
            if (cc_rtex != MPT_TEXTURE_NONE) coat_rough *= tex(cc_rtex, sd).y;
            coat_alpha = pow2(std::max(coat_rough, MIN_ROUGHNESS));
        }
        float3 absorb(0.f);
        if (sd.front_facing && V) absorb = vol_abs;
        float3 sheen_col(0.f);
        float sheen_rough = 0.f;
        if (S) {
            sheen_col = sh_col;
            if (sh_ctex != MPT_TEXTURE_NONE) sheen_col = sheen_col * tex(sh_ctex, sd).xyz();
            sheen_rough = sh_r;
            if (sh_rtex != MPT_TEXTURE_NONE) sheen_rough *= tex(sh_rtex, sd).w;
        }
        const float sheen_alpha = pow2(sheen_rough);
        const float max_sheen = std::max(sheen_col.x, std::max(sheen_col.y, sheen_col.z));
        const GltfSheened sheened(SheenBRDF(sheen_col, sheen_alpha), reflective, sheen_alpha, max_sheen);
        s.is_gltf = true;
        s.gltf = GltfBSDF(sheened, GltfCoated(SpecularBRDF(coat_alpha), sheened, GLTF_DIELECTRIC_IOR_F0), coat_weight);
        out_frame = shading_frame;
        s.eta = eta;
        s.absorption = absorb;
        return s;
    }

    // ---------------------------------------------------------------- geometry access
    const mpt_geometry_desc& geometry(const Hit& h) const { return geometries[h.instance_id + h.geometry_index]; }
    float2 vertex_uv(const mpt_packed_vertex& v) const { return float2(half_to_float(v.uv[0]), half_to_float(v.uv[1])); }
    static float3 vpos(const mpt_packed_vertex& v) { return float3(v.position[0], v.position[1], v.position[2]); }

    float3 mul_point(const float* m, float3 p) const { // mul(float4x4, float4(p, 1)).xyz
        return float3(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
                      m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]);
    }
    float3 mul_vec(const float* m, float3 v) const { // mul((float3x3)M, v) / mul(M, float4(v, 0)).xyz
        return float3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z,
                      m[8] * v.x + m[9] * v.y + m[10] * v.z);
    }
    // scene.slang:236-259
    float3 to_world_position(const mpt_geometry_desc& gd, uint32_t ii, float3 p) const {
        if (gd.flags & MPT_GEOM_PRETRANSFORMED) return p;
        return mul_point(&instance_transforms[size_t(ii) * 16], p);
    }
    float3 to_world_normal(const mpt_geometry_desc& gd, uint32_t ii, float3 n) const {
        if (gd.flags & MPT_GEOM_PRETRANSFORMED) return n;
        return normalize(mul_vec(&inv_t_instance_transforms[size_t(ii) * 16], n));
    }
    float3 to_world_vector(const mpt_geometry_desc& gd, uint32_t ii, float3 v) const {
        if (gd.flags & MPT_GEOM_PRETRANSFORMED) return v;
        return mul_vec(&instance_transforms[size_t(ii) * 16], v);
    }

    // ---------------------------------------------------------------- traversal (Scene::trace_ray, scene.slang:62-76, 500-519)
    void intersect_blas(const Blas& b, const Ray& r, uint32_t instance_index, const mpt_instance_desc& inst, Hit& best,
                        float& best_t) const {
        const RayShear rs(r.d);
        const float3 inv_d(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
        const int cull = (inst.flags & MPT_INSTANCE_FACING_CULL_DISABLE) ? CULL_NONE
                         : ((inst.flags & MPT_INSTANCE_FLIP_FACING) ? CULL_FRONT : CULL_BACK);
        if (b.bvh.nodes.empty()) return;
        uint32_t stack[64];
        int sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const Bvh2::Node& n = b.bvh.nodes[stack[--sp]];
            if (!ray_box(n.box, r.o, inv_d, r.tmin, best_t)) continue;
            if (n.count == 0) {
                stack[sp++] = n.left;
                stack[sp++] = n.left + 1;
                continue;
            }
            for (uint32_t i = n.left; i < n.left + n.count; i++) {
                const Blas::Prim pr = b.prims[b.bvh.prim_index[i]];
                const Mesh& m = meshes[b.meshes[pr.geometry_index]];
                uint32_t ia, ib, ic;
                m.tri(pr.primitive_id, ia, ib, ic);
                float t, u, v;
                if (!intersect_triangle(rs, r.o, vpos(m.vertices[ia]), vpos(m.vertices[ib]), vpos(m.vertices[ic]), cull, r.tmin,
                                        best_t, t, u, v))
                    continue;
                if (!m.opaque && !(inst.flags & MPT_INSTANCE_FORCE_OPAQUE)) {
                    // CANDIDATE_NON_OPAQUE_TRIANGLE: alpha test, scene.slang:503-516
                    const mpt_geometry_desc& gd = geometries[inst.instance_custom_index + pr.geometry_index];
                    const float3 b3(1.0f - u - v, u, v);
                    const float2 uv = b3.x * vertex_uv(m.vertices[ia]) + b3.y * vertex_uv(m.vertices[ib]) + b3.z * vertex_uv(m.vertices[ic]);
                    if (alpha_test(gd.material_id, uv)) continue;
                }
                best_t = t;
                best.instance_id = inst.instance_custom_index;
                best.instance_index = instance_index;
                best.primitive_id = pr.primitive_id;
                best.geometry_index = pr.geometry_index;
                best.bary = float2(u, v);
            }
        }
    }

    Hit trace_ray(const Ray& ray, uint32_t mask, float* t_out = nullptr) const {
        Hit best;
        float best_t = ray.tmax;
        const float3 inv_d(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);
        auto visit = [&](uint32_t ii) {
            const mpt_instance_desc& inst = instances[ii];
            if ((inst.mask & mask & 0xFFu) == 0) return;
            if (!ray_box(instance_bounds[ii], ray.o, inv_d, ray.tmin, best_t)) return;
            const float* m = instance_inverse[ii].data();
            Ray r = ray;
            // world -> object; direction is not renormalised so t is preserved (Vulkan spec, ray transform)
            r.o = float3(m[0] * ray.o.x + m[1] * ray.o.y + m[2] * ray.o.z + m[3], m[4] * ray.o.x + m[5] * ray.o.y + m[6] * ray.o.z + m[7],
                         m[8] * ray.o.x + m[9] * ray.o.y + m[10] * ray.o.z + m[11]);
            r.d = float3(m[0] * ray.d.x + m[1] * ray.d.y + m[2] * ray.d.z, m[4] * ray.d.x + m[5] * ray.d.y + m[6] * ray.d.z,
                         m[8] * ray.d.x + m[9] * ray.d.y + m[10] * ray.d.z);
            intersect_blas(blas[inst.blas], r, ii, inst, best, best_t);
        };
        if (instances.size() <= 8) {
            for (uint32_t ii = 0; ii < instances.size(); ii++) visit(ii);
        } else {
            uint32_t stack[64];
            int sp = 0;
            stack[sp++] = 0;
            while (sp) {
                const Bvh2::Node& n = tlas.nodes[stack[--sp]];
                if (!ray_box(n.box, ray.o, inv_d, ray.tmin, best_t)) continue;
                if (n.count == 0) { stack[sp++] = n.left; stack[sp++] = n.left + 1; continue; }
                for (uint32_t i = n.left; i < n.left + n.count; i++) visit(tlas.prim_index[i]);
            }
        }
        if (t_out) *t_out = best_t;
        return best;
    }

    // ---------------------------------------------------------------- shading point (scene.slang:312-498)
    struct Surface { uint32_t material_id; float3 pos; float2 uv; float3 face_normal; Frame frame; };

    Surface get_surface(const Hit& hit, const mpt_geometry_desc& gd, const mpt_packed_vertex& v0, const mpt_packed_vertex& v1,
                        const mpt_packed_vertex& v2) const {
        const float3 bary = hit.bary3();
        const uint32_t ii = hit.instance_index;
        const float3 p_obj = bary.x * vpos(v0) + bary.y * vpos(v1) + bary.z * vpos(v2); // TriangleHit::get_position, hit.slang:74-77
        const float3 pos = to_world_position(gd, ii, p_obj);
        const float3 edge01 = to_world_vector(gd, ii, vpos(v1) - vpos(v0));
        const float3 edge02 = to_world_vector(gd, ii, vpos(v2) - vpos(v0));
        float3 face_normal = normalize(cross(edge01, edge02));
        if (gd.flags & MPT_GEOM_FLIP_FACING) face_normal = -face_normal;
        float3 vertex_normal;
        if (gd.flags & MPT_GEOM_FLAT_SHADING) {
            vertex_normal = face_normal;
        } else {
            const float3 n_obj = normalize(bary.x * decode_normal(v0.encoded_normal) + bary.y * decode_normal(v1.encoded_normal) +
                                           bary.z * decode_normal(v2.encoded_normal));
            vertex_normal = to_world_normal(gd, ii, n_obj);
        }
        float3 tangent;
        float handedness;
        bool tangent_valid;
        const float2 uv0 = vertex_uv(v0), uv1 = vertex_uv(v1), uv2 = vertex_uv(v2);
        if (gd.flags & MPT_GEOM_HAS_TANGENTS) {
            const float4 t0 = decode_tangent(v0.encoded_tangent), t1 = decode_tangent(v1.encoded_tangent), t2 = decode_tangent(v2.encoded_tangent);
            const float4 t = bary.x * t0 + bary.y * t1 + bary.z * t2;
            tangent = to_world_vector(gd, ii, normalize(t.xyz()));
            handedness = signf(t.w);
            tangent_valid = true;
        } else {
            const float2 duv1 = uv1 - uv0, duv2 = uv2 - uv0;
            const float det = duv1.x * duv2.y - duv2.x * duv1.y;
            if (std::fabs(det) > 1e-8f) {
                const float winding_sign = (gd.flags & MPT_GEOM_FLIP_FACING) ? 1.f : -1.f;
                tangent = (duv2.y * edge01 - duv1.y * edge02) / det;
                handedness = signf(det) * winding_sign;
                tangent_valid = true;
            } else {
                tangent = float3(1, 0, 0);
                handedness = 1.f;
                tangent_valid = false;
            }
        }
        Surface s;
        s.material_id = gd.material_id;
        s.pos = pos;
        s.uv = bary.x * uv0 + bary.y * uv1 + bary.z * uv2; // TriangleHit::get_uv, hit.slang:78-81
        s.face_normal = face_normal;
        s.frame = tangent_valid ? Frame(vertex_normal, tangent, handedness) : Frame(vertex_normal);
        return s;
    }

    // environment-map.slang:23-35
    float3 env_eval(float3 dir) const {
        if (env_type != MPT_ENV_LATLONG) return float3(0.f);
        const float3 d = mul_vec(env_to_local, dir);
        return env_intensity * sample_texture(env_tex, world_to_latlong_map(d)).xyz();
    }

    // scene.slang:482-498
    ShadingPoint make_environment_map_shading_point(const Ray& ray) const {
        ShadingPoint sp;
        sp.pos = ray.o + ray.d * ENV_MAP_DISTANCE;
        sp.enc_face_normal = encode_normal(-ray.d);
        sp.emission = env_eval(ray.d); // eval uses -sd.wi = ray.Direction
        sp.shading_frame = Frame(-ray.d);
        sp.has_surface_material = false;
        return sp;
    }

    // scene.slang:418-469
    ShadingPoint get_shading_point(const Ray& ray, const Hit& hit, uint16_t bounce = 0, float ior = 1.0f, bool with_material = true) const {
        if (!hit.valid()) return make_environment_map_shading_point(ray);
        const mpt_geometry_desc& gd = geometry(hit);
        if (gd.flags & MPT_GEOM_USE_ENV_MAP) return make_environment_map_shading_point(ray);
        const Mesh& m = meshes[gd.mesh];
        uint32_t ia, ib, ic;
        m.tri(hit.primitive_id, ia, ib, ic);
        const Surface surface = get_surface(hit, gd, m.vertices[ia], m.vertices[ib], m.vertices[ic]);
        Frame frame = surface.frame;
        float3 face_normal = surface.face_normal;
        bool front_facing = true;
        if (gd.flags & MPT_GEOM_TWO_SIDED) {
            front_facing = dot(-ray.d, surface.face_normal) >= 0.f;
            if (!front_facing) {
                frame.flip_normal();
                face_normal = -face_normal;
            }
        }
        ShadingData sd{-ray.d, surface.pos, surface.uv, frame, face_normal, bounce, ior, front_facing};
        ShadingPoint sp;
        sp.pos = surface.pos;
        sp.enc_face_normal = encode_normal(surface.face_normal);
        sp.has_surface_material = with_material;
        if (with_material) {
            sp.surface_material = get_material_sample(surface.material_id, sd, sp.emission, sp.shading_frame);
        } else {
            Frame unused;
            get_material_sample(surface.material_id, sd, sp.emission, unused, true);
            sp.shading_frame = frame;
        }
        return sp;
    }

    // hit.slang:134-147
    static std::optional<Ray> make_ray(const ShadingPoint& sp, float3 wi, float3 wo, bool transmission) {
        const float3 ng = sp.get_face_normal();
        const float wo_ng = dot(wo, ng);
        if (((dot(wi, ng) < 0.f) != (wo_ng < 0.f)) != transmission) return std::nullopt;
        Ray r;
        r.o = offset_ray_origin(sp.pos, wo_ng < 0.f ? -ng : ng);
        r.d = wo;
        r.tmin = SCENE_RAY_TMIN;
        r.tmax = SCENE_RAY_TMAX;
        return r;
    }

    // Scene::generate_camera_ray, scene.slang:91-101: pinhole (camera.slang:45-56) or thin lens (camera.slang:73-92) when the camera has an
    // aperture (merian_hint_enable_thin_lens follows the active camera, scene.cpp:2019-2021). frame = Scene::frame (scene.slang:48).
    static uint32_t xxhash32(uint32_t px, uint32_t py) { // hash.slang:68-77
        const uint32_t P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
        uint32_t h = py + P5 + px * P3;
        h = P4 * ((h << 17) | (h >> 15));
        h = P2 * (h ^ (h >> 15));
        h = P3 * (h ^ (h >> 13));
        return h ^ (h >> 16);
    }
    static float2 r2_sequence(uint32_t n) { // pseudorandom.slang:35-38
        return float2(frac(0.5f + 0.7548776662466927f * float(n)), frac(0.5f + 0.5698402909980532f * float(n)));
    }
    static Ray generate_camera_ray(const mpt_camera& c, uint32_t px, uint32_t py, uint32_t w, uint32_t h, uint32_t frame = 0) {
        const float2 p((float(px) + 0.5f + c.jitter[0]) / float(w), (float(py) + 0.5f + c.jitter[1]) / float(h));
        const float2 ndc(2.f * p.x + -1.f, -2.f * p.y + 1.f);
        const float3 U(c.U[0], c.U[1], c.U[2]), V(c.V[0], c.V[1], c.V[2]), W(c.W[0], c.W[1], c.W[2]);
        Ray r;
        r.o = float3(c.position[0], c.position[1], c.position[2]);
        float3 dir = ndc.x * U + ndc.y * V + W;
        if (c.aperture_radius > 0.f) {
            Rng rng = Rng::pcg(xxhash32(px, py));
            const float2 r2 = r2_sequence(frame), xi = rng.next_float2();
            const float2 random(frac(r2.x + xi.x), frac(r2.y + xi.y));
            const float angle = 6.28318530717958647692f * random.x, rad = std::sqrt(random.y); // sample_disk, sampling.slang:39-42
            const float2 a(std::cos(angle) * rad, std::sin(angle) * rad);
            const float3 aperture_offset = c.aperture_radius * (a.x * normalize(U) + a.y * normalize(V));
            r.o = r.o + aperture_offset;
            dir = dir - aperture_offset;
        }
        r.d = normalize(dir);
        const float inv_cos = 1.f / dot(normalize(W), r.d);
        r.tmin = c.near_plane * inv_cos;
        r.tmax = c.far_plane * inv_cos;
        return r;
    }

    // camera.slang:58-70
    static float2 project_pinhole(const mpt_camera& c, float3 world_pos, uint32_t w, uint32_t h) {
        const float3 U(c.U[0], c.U[1], c.U[2]), V(c.V[0], c.V[1], c.V[2]), W(c.W[0], c.W[1], c.W[2]);
        const float3 d = world_pos - float3(c.position[0], c.position[1], c.position[2]);
        const float t = dot(W, W) / dot(d, W);
        const float3 plane_pos = t * d - W;
        const float2 ndc(dot(plane_pos, U) / dot(U, U), dot(plane_pos, V) / dot(V, V));
        const float2 p((ndc.x - -1.f) / 2.f, (ndc.y - 1.f) / -2.f);
        return float2(p.x * float(w) - 0.5f, p.y * float(h) - 0.5f);
    }

    // One pixel of the GBuffer (Raytraced) node: gbuffer_rt.slang:18-94, stores of gbuffer.slang:152-172.
    // Textures are sampled at LOD 0 (mpt_texture_upload takes no mip chain), so the ray-differential UV gradients of
    // gbuffer_rt.slang:41-50 have no consumer and are not evaluated. prev_pos: scene.slang:392-401 without HasPrevVertices.
    struct GBufferTexel {
        uint32_t tex0[2];
        Hit hit; // packed into tex1 by the caller (gbuffer.slang:167-169)
        float tex2[4], tex3[4], emission[4];
    };
    uint32_t frame = 0; // Scene::frame (scene.slang:48) for the passes without render params (oracle_set_scene_frame)
    GBufferTexel gbuffer_pixel(const mpt_camera& cam, const mpt_camera& prev_cam, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint32_t mask) const {
        GBufferTexel g;
        const Ray ray = generate_camera_ray(cam, x, y, w, h, frame);
        const Hit hit = trace_ray(ray, mask);
        g.hit = hit;
        float3 normal, albedo, emission;
        float linear_z, delta_z = 0.f;
        float2 grad_z(0.f, 0.f), mv;
        bool surface_hit = false;
        if (hit.valid()) {
            const mpt_geometry_desc& gd = geometry(hit);
            if ((gd.flags & MPT_GEOM_USE_ENV_MAP) == 0) {
                surface_hit = true;
                const Mesh& m = meshes[gd.mesh];
                uint32_t ia, ib, ic;
                m.tri(hit.primitive_id, ia, ib, ic);
                const Surface surface = get_surface(hit, gd, m.vertices[ia], m.vertices[ib], m.vertices[ic]);
                // Scene::get_prev_world_position (scene.slang:392-401) without HasPrevVertices: the object-space hit point carried by the
                // previous frame's instance transform for animated geometry (to_prev_world_position, :261-270)
                float3 prev_pos = surface.pos;
                if (!prev_instance_transforms.empty() && (gd.flags & MPT_GEOM_IS_ANIMATED) && !(gd.flags & MPT_GEOM_PRETRANSFORMED)) {
                    const float3 bary = hit.bary3();
                    const float3 p_obj = bary.x * vpos(m.vertices[ia]) + bary.y * vpos(m.vertices[ib]) + bary.z * vpos(m.vertices[ic]);
                    prev_pos = mul_point(&prev_instance_transforms[size_t(hit.instance_index) * 16], p_obj);
                }
                linear_z = length(surface.pos - ray.o);
                mpt_camera pin = cam; // get_pinhole_direction (camera.slang:34-37), whatever the aperture
                pin.aperture_radius = 0.f;
                const float3 r_x = generate_camera_ray(pin, x + 1, y, w, h).d, r_y = generate_camera_ray(pin, x, y + 1, w, h).d;
                const float n_dot_oo = dot(surface.face_normal, surface.pos - ray.o);                          // hit.slang:99-103
                grad_z = float2(n_dot_oo / dot(surface.face_normal, r_x) - linear_z, n_dot_oo / dot(surface.face_normal, r_y) - linear_z);
                delta_z = length(prev_pos - float3(prev_cam.position[0], prev_cam.position[1], prev_cam.position[2])) - linear_z;
                const float2 pp = project_pinhole(prev_cam, prev_pos, w, h);                                     // scene.slang:404-414
                mv = float2(pp.x - float(x) - cam.jitter[0], pp.y - float(y) - cam.jitter[1]);
                Frame frame = surface.frame;
                float3 face_normal = surface.face_normal;
                bool front_facing = true;
                if (gd.flags & MPT_GEOM_TWO_SIDED) {
                    front_facing = dot(-ray.d, surface.face_normal) >= 0.f;
                    if (!front_facing) { frame.flip_normal(); face_normal = -face_normal; }
                }
                ShadingData sd{-ray.d, surface.pos, surface.uv, frame, face_normal, 0, 1.0f, front_facing};
                {   // screen-space UV gradients for the mip selection, gbuffer_rt.slang:41-50 + ray-differential.slang:13-62
                    // (RayDifferential::from_pinhole_camera, propagate: Igehy eq. 8, 10, 12; compute_uv_gradients)
                    const float3 e01 = to_world_vector(gd, hit.instance_index, vpos(m.vertices[ib]) - vpos(m.vertices[ia]));
                    const float3 e02 = to_world_vector(gd, hit.instance_index, vpos(m.vertices[ic]) - vpos(m.vertices[ia]));
                    const float3 N = surface.face_normal;
                    float3 dOdx = linear_z * (r_x - ray.d), dOdy = linear_z * (r_y - ray.d);
                    const float rcpDN = 1.0f / dot(ray.d, N);
                    dOdx = dOdx + ray.d * (-dot(dOdx, N) * rcpDN);
                    dOdy = dOdy + ray.d * (-dot(dOdy, N) * rcpDN);
                    const float3 Nu = cross(e02, N), Nv = cross(e01, N);
                    const float3 Lu = Nu / dot(Nu, e01), Lv = Nv / dot(Nv, e02);
                    const float2 d1 = vertex_uv(m.vertices[ib]) - vertex_uv(m.vertices[ia]), d2 = vertex_uv(m.vertices[ic]) - vertex_uv(m.vertices[ia]);
                    sd.ddx = dot(Lu, dOdx) * d1 + dot(Lv, dOdx) * d2;
                    sd.ddy = dot(Lu, dOdy) * d1 + dot(Lv, dOdy) * d2;
                    sd.use_grad = true;
                }
                Frame mat_frame;
                float3 mat_emission;
                const MaterialSample mat = get_material_sample(surface.material_id, sd, mat_emission, mat_frame);
                normal = mat_frame.n;
                albedo = mat.get_albedo(mat_frame.to_local(-ray.d));
                emission = mat_emission; // x exterior_volume.transmittance(linear_z) = 1 (Vacuum)
            }
        }
        if (!surface_hit) {
            const ShadingPoint sp = make_environment_map_shading_point(ray);
            linear_z = length(sp.pos - ray.o);
            const float2 pp = project_pinhole(prev_cam, sp.pos, w, h);
            mv = float2(pp.x - float(x) - cam.jitter[0], pp.y - float(y) - cam.jitter[1]);
            normal = sp.shading_frame.n;
            albedo = sp.emission;
            emission = sp.emission;
        }
        g.tex0[0] = encode_normal(normal);
        g.tex0[1] = uint32_t(float_to_half(grad_z.x)) | (uint32_t(float_to_half(grad_z.y)) << 16);
        g.tex2[0] = mv.x; g.tex2[1] = mv.y; g.tex2[2] = std::min(linear_z, 65504.0f); g.tex2[3] = delta_z;
        g.tex3[0] = albedo.x; g.tex3[1] = albedo.y; g.tex3[2] = albedo.z; g.tex3[3] = 1.f;
        g.emission[0] = emission.x; g.emission[1] = emission.y; g.emission[2] = emission.z; g.emission[3] = 0.f;
        return g;
    }

    // render_pt.slang:24-90
    float3 trace_path(Ray ray, ShadingPoint sp, Rng& rng, const mpt_render_params& P, uint64_t* ray_count) const {
        Medium medium; // scene.exterior_volume = Vacuum
        float current_ior = 1.0f;
        float3 throughput = medium.transmittance(distance(ray.o, sp.pos));
        float3 radiance = (P.flags & MPT_RENDER_EMISSION_ON_PRIMARY) ? throughput * sp.emission : float3(0.f);
        for (int bounce = 1; bounce < P.max_path_length; bounce++) {
            if (!sp.has_surface_material || all_lt(throughput, SCENE_THROUGHPUT_THRESHOLD)) break;
            const Frame frame = sp.shading_frame;
            const float3 wi_local = frame.to_local(-ray.d);
            const OptSample sample = sp.surface_material.sample_eval(wi_local, rng);
            if (!sample) break;
            const float3 wo = frame.to_world(sample->wo);
            throughput *= sample->weight;
            const bool transmission = wi_local.z * sample->wo.z < 0.f;
            if (transmission) {
                const Medium interior = sp.surface_material.get_interior_volume();
                current_ior /= interior.eta;
                if (dot(-ray.d, sp.get_face_normal()) >= 0.f) medium = interior;
                else medium = Medium();
            }
            const auto next = make_ray(sp, -ray.d, wo, transmission);
            if (!next) break;
            ray = *next;
            (*ray_count)++;
            const Hit hit = trace_ray(ray, P.instance_mask);
            sp = get_shading_point(ray, hit, uint16_t(bounce), current_ior, bounce + 1 < P.max_path_length);
            throughput *= medium.transmittance(distance(ray.o, sp.pos));
            radiance += throughput * sp.emission;
        }
        return radiance;
    }
};

// gbuffer.slang:40-61
struct PackedHitInfo { uint32_t instance_id, primitive_id; uint16_t geometry_index, instance_index; uint16_t bary[2]; };
static_assert(sizeof(PackedHitInfo) == 16, "PackedHitInfo must be one uint4 texel");

inline PackedHitInfo pack_hit(const Hit& h) {
    PackedHitInfo p;
    p.instance_id = h.instance_id;
    p.primitive_id = h.primitive_id;
    p.geometry_index = uint16_t(h.geometry_index);
    p.instance_index = uint16_t(h.instance_index);
    p.bary[0] = float_to_half(h.bary.x);
    p.bary[1] = float_to_half(h.bary.y);
    return p;
}
inline Hit unpack_hit(const PackedHitInfo& p) {
    Hit h;
    h.instance_id = p.instance_id;
    h.instance_index = uint32_t(p.instance_index);
    h.primitive_id = p.primitive_id;
    h.geometry_index = uint32_t(p.geometry_index);
    h.bary = float2(half_to_float(p.bary[0]), half_to_float(p.bary[1]));
    return h;
}

// 3x4 affine inverse (row-major)
inline std::array<float, 12> invert_affine(const float* m) {
    const double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    const double id = 1.0 / det;
    double r[9] = {(e * i - f * h) * id, (c * h - b * i) * id, (b * f - c * e) * id, (f * g - d * i) * id, (a * i - c * g) * id,
                   (c * d - a * f) * id, (d * h - e * g) * id, (b * g - a * h) * id, (a * e - b * d) * id};
    std::array<float, 12> o;
    for (int row = 0; row < 3; row++) {
        for (int col = 0; col < 3; col++) o[row * 4 + col] = float(r[row * 3 + col]);
        o[row * 4 + 3] = float(-(r[row * 3] * m[3] + r[row * 3 + 1] * m[7] + r[row * 3 + 2] * m[11]));
    }
    return o;
}

} // namespace orc

#include "oracle_bvh8.hpp" // walker over the product's BVH8 (counts nodes / triangles per ray for the roofline)

// ---- BSDF property suite: tests/bsdf/bsdf-checks.slang:30-135 restated ---------------------------
using namespace orc;
namespace {
inline float accumulate_err(float acc, float v) { return std::isfinite(v) ? std::max(acc, v) : 1e30f; }

template <class B> void bsdf_checks(const B& bsdf, float3 wi, uint32_t n_samples, uint32_t seed, float* out) {
    Rng rng = Rng::xorshift(seed | 1u);
    float3 mean_weight(0.f), m2_weight(0.f);
    float mean_cos = 0, m2_cos = 0;
    uint32_t valid = 0;
    float max_weight_err = 0, max_pdf_err = 0;
    for (uint32_t i = 0; i < n_samples; ++i) {
        float3 weight(0.f);
        float c = 0;
        if (auto s = bsdf.sample_eval(wi, rng)) {
            valid++;
            weight = s->weight;
            c = s->wo.z / s->pdf;
            const float pdf_ref = bsdf.pdf(wi, s->wo);
            const float3 eval_ref = bsdf.eval(wi, s->wo, rng);
            max_pdf_err = accumulate_err(max_pdf_err, std::fabs(s->pdf - pdf_ref) / (std::max(std::fabs(s->pdf), std::fabs(pdf_ref)) + 1e-6f));
            if (pdf_ref > 0) max_weight_err = accumulate_err(max_weight_err, length(s->weight - (eval_ref / pdf_ref)));
        }
        const float inv_i = 1.0f / float(i + 1);
        const float3 delta_weight = weight - mean_weight;
        mean_weight += delta_weight * inv_i;
        m2_weight += delta_weight * (weight - mean_weight);
        const float delta_cos = c - mean_cos;
        mean_cos += delta_cos * inv_i;
        m2_cos += delta_cos * (c - mean_cos);
    }
    const float inv_n = 1.0f / float(n_samples);
    const float3 furnace = mean_weight;
    const float3 furnace_stderr = sqrt3(m2_weight) * inv_n;
    const float stderr_cos = std::sqrt(m2_cos) * inv_n;
    const uint32_t NT = 256, NP = 512;
    const float dtheta = (PI * 0.5f) / float(NT), dphi = TWO_PI / float(NP);
    float pdf_integral = 0, max_recip_err = 0;
    for (uint32_t ti = 0; ti < NT; ++ti) {
        const float theta = (float(ti) + 0.5f) * dtheta;
        const float ct = std::cos(theta), st = std::sin(theta);
        for (uint32_t pi = 0; pi < NP; ++pi) {
            const float phi = (float(pi) + 0.5f) * dphi;
            const float3 wo(st * std::cos(phi), st * std::sin(phi), ct);
            pdf_integral += bsdf.pdf(wi, wo) * st * dtheta * dphi;
            const float3 f_io = bsdf.eval(wi, wo, rng) / wo.z;
            const float3 f_oi = bsdf.eval(wo, wi, rng) / wi.z;
            const float3 d = abs3(f_io - f_oi);
            const float3 m = max3v(abs3(f_io), abs3(f_oi)) + 1e-6f;
            max_recip_err = accumulate_err(max_recip_err, std::max(std::max(d.x / m.x, d.y / m.y), d.z / m.z));
        }
    }
    const float3 albedo = bsdf.get_albedo(wi);
    out[0] = furnace.x; out[1] = furnace.y; out[2] = furnace.z;
    out[3] = furnace_stderr.x; out[4] = furnace_stderr.y; out[5] = furnace_stderr.z;
    out[6] = mean_cos; out[7] = stderr_cos; out[8] = float(valid);
    out[9] = max_weight_err; out[10] = max_pdf_err; out[11] = pdf_integral; out[12] = max_recip_err;
    out[13] = albedo.x; out[14] = albedo.y; out[15] = albedo.z;
}
} // namespace


// ==================================================================================================
// C API (ctypes) — mirrors include/mpt.h entry for entry so tests read the same for both arms.
struct oracle_context {
    Scene scene;
    std::string error;
};

extern "C" {

oracle_context* oracle_create() { return new oracle_context(); }
void oracle_destroy(oracle_context* c) { delete c; }
const char* oracle_last_error(oracle_context* c) { return c->error.c_str(); }

int oracle_scene_upload(oracle_context* c, const mpt_scene_desc* d) {
    Scene& s = c->scene;
    s.meshes.clear(); s.blas.clear();
    s.meshes.resize(d->mesh_count);
    for (uint32_t i = 0; i < d->mesh_count; i++) {
        const mpt_mesh_desc& md = d->meshes[i];
        Mesh& m = s.meshes[i];
        m.vertices.assign(md.vertices, md.vertices + md.vertex_count);
        m.primitive_count = md.primitive_count;
        m.index_flag = md.index_flag;
        m.opaque = md.opaque != 0;
        if (md.index_flag) {
            m.indices.resize(size_t(md.primitive_count) * 3);
            for (size_t k = 0; k < m.indices.size(); k++) {
                if (md.index_flag == MPT_GEOM_INDEX_U32) m.indices[k] = ((const uint32_t*)md.indices)[k];
                else if (md.index_flag == MPT_GEOM_INDEX_U16) m.indices[k] = ((const uint16_t*)md.indices)[k];
                else m.indices[k] = ((const uint8_t*)md.indices)[k];
            }
        }
    }
    s.blas.resize(d->blas_count);
    for (uint32_t b = 0; b < d->blas_count; b++) {
        Blas& bl = s.blas[b];
        bl.meshes.assign(d->blas[b].meshes, d->blas[b].meshes + d->blas[b].mesh_count);
        std::vector<AABB> boxes;
        for (uint32_t g = 0; g < bl.meshes.size(); g++) {
            const Mesh& m = s.meshes[bl.meshes[g]];
            for (uint32_t p = 0; p < m.primitive_count; p++) {
                uint32_t ia, ib, ic;
                m.tri(p, ia, ib, ic);
                AABB bx;
                bx.grow(Scene::vpos(m.vertices[ia])); bx.grow(Scene::vpos(m.vertices[ib])); bx.grow(Scene::vpos(m.vertices[ic]));
                boxes.push_back(bx);
                bl.prims.push_back({g, p});
                bl.bounds.grow(bx);
            }
        }
        bl.bvh.build(boxes);
    }
    s.geometries.assign(d->geometries, d->geometries + d->geometry_count);
    s.instances.assign(d->instances, d->instances + d->instance_count);
    s.instance_transforms.clear(); s.inv_t_instance_transforms.clear();
    if (d->instance_transforms) s.instance_transforms.assign(d->instance_transforms, d->instance_transforms + size_t(d->instance_count) * 16);
    if (d->inverse_transposed_instance_transforms)
        s.inv_t_instance_transforms.assign(d->inverse_transposed_instance_transforms, d->inverse_transposed_instance_transforms + size_t(d->instance_count) * 16);
    s.instance_bounds.resize(d->instance_count);
    s.instance_inverse.resize(d->instance_count);
    for (uint32_t i = 0; i < d->instance_count; i++) {
        const mpt_instance_desc& in = s.instances[i];
        const AABB& ob = s.blas[in.blas].bounds;
        AABB wb;
        for (int k = 0; k < 8; k++) {
            const float3 p((k & 1) ? ob.hi.x : ob.lo.x, (k & 2) ? ob.hi.y : ob.lo.y, (k & 4) ? ob.hi.z : ob.lo.z);
            const float* m = in.transform;
            wb.grow(float3(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
                           m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]));
        }
        // pad: the object-space test is exact, the world box only prunes
        const float3 pad = (wb.hi - wb.lo) * 1e-5f + float3(1e-6f);
        wb.lo = wb.lo - pad; wb.hi = wb.hi + pad;
        s.instance_bounds[i] = wb;
        s.instance_inverse[i] = invert_affine(in.transform);
    }
    s.tlas.build(s.instance_bounds);
    return 0;
}

int oracle_materials_upload(oracle_context* c, const void* entries, uint32_t entry_size, uint32_t count, const mpt_material_type* types,
                            uint32_t type_count) {
    Scene& s = c->scene;
    s.material_entries.assign((const uint8_t*)entries, (const uint8_t*)entries + size_t(entry_size) * count);
    s.material_entry_size = entry_size;
    s.material_count = count;
    s.material_types.assign(types, types + type_count);
    return 0;
}

int oracle_materials_set_hints(oracle_context* c, float alpha_test_threshold, float min_roughness, uint32_t clamp_normals) {
    c->scene.ALPHA_TEST_THRESHOLD = alpha_test_threshold;
    c->scene.MIN_ROUGHNESS = min_roughness;
    c->scene.clamp_normals = clamp_normals != 0;
    return 0;
}

int oracle_texture_upload_ex(oracle_context* c, uint32_t id, uint32_t w, uint32_t h, uint32_t format, uint32_t wrap_u, uint32_t wrap_v,
                             uint32_t filter, uint32_t mipmap_mode, const void* pixels) {
    Scene& s = c->scene;
    if (s.textures.size() <= id) s.textures.resize(id + 1);
    Texture& t = s.textures[id];
    t = Texture();
    t.w = w; t.h = h; t.format = format; t.wrap_u = wrap_u; t.wrap_v = wrap_v; t.filter = filter; t.mipmap_mode = mipmap_mode;
    if (format == MPT_TEX_RGBA32F) t.rgba32f.assign((const float*)pixels, (const float*)pixels + size_t(w) * h * 4);
    else t.rgba8.assign((const uint8_t*)pixels, (const uint8_t*)pixels + size_t(w) * h * 4);
    if (mipmap_mode != MPT_MIPMAP_NONE) generate_mipmaps(t);
    return 0;
}
int oracle_texture_upload(oracle_context* c, uint32_t id, uint32_t w, uint32_t h, uint32_t format, uint32_t wrap_u, uint32_t wrap_v,
                          uint32_t filter, const void* pixels) {
    return oracle_texture_upload_ex(c, id, w, h, format, wrap_u, wrap_v, filter, MPT_MIPMAP_NONE, pixels);
}
// one texel of one mip level, decoded (test hook for the mip chain)
void oracle_texture_level_texel(oracle_context* c, uint32_t id, uint32_t level, uint32_t x, uint32_t y, float* out4, uint32_t* dims2) {
    const Texture& t = c->scene.textures[id].level(level);
    const float4 v = texel(t, int(x), int(y));
    out4[0] = v.x; out4[1] = v.y; out4[2] = v.z; out4[3] = v.w;
    if (dims2) { dims2[0] = t.w; dims2[1] = t.h; }
}
void oracle_sample_texture_grad(oracle_context* c, uint32_t id, float u, float v, const float* ddx, const float* ddy, float* out) {
    const float4 t = sample_texture_grad(c->scene.textures[id], float2(u, v), float2(ddx[0], ddx[1]), float2(ddy[0], ddy[1]));
    out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
}

int oracle_env_set(oracle_context* c, uint32_t type, const float* to_local, float intensity, uint32_t w, uint32_t h, const float* rgba32f) {
    Scene& s = c->scene;
    s.env_type = type;
    if (type == MPT_ENV_LATLONG) {
        std::memcpy(s.env_to_local, to_local, 64);
        s.env_intensity = intensity;
        s.env_tex.w = w; s.env_tex.h = h; s.env_tex.format = MPT_TEX_RGBA32F;
        s.env_tex.wrap_u = s.env_tex.wrap_v = MPT_WRAP_REPEAT; // scene.cpp:786-789
        s.env_tex.filter = MPT_FILTER_LINEAR;
        s.env_tex.rgba32f.assign(rgba32f, rgba32f + size_t(w) * h * 4);
    }
    return 0;
}

// gbuffer_rt.slang:25-26 + gbuffer.slang:167-169 (store_hit)
int oracle_primary_visibility(oracle_context* c, const mpt_camera* cam, uint32_t w, uint32_t h, uint32_t mask, void* tex1) {
    PackedHitInfo* out = (PackedHitInfo*)tex1;
    const Scene& s = c->scene;
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < int(h); y++)
        for (uint32_t x = 0; x < w; x++) {
            const Ray r = Scene::generate_camera_ray(*cam, x, uint32_t(y), w, h, s.frame);
            out[size_t(y) * w + x] = pack_hit(s.trace_ray(r, mask));
        }
    return 0;
}

// GBuffer (Raytraced) node, gbuffer_rt.slang. Any of the output images may be NULL.
int oracle_gbuffer(oracle_context* c, const mpt_camera* cam, const mpt_camera* prev_cam, uint32_t w, uint32_t h, uint32_t mask, uint32_t* tex0,
                   void* tex1, float* tex2, float* tex3, float* emission) {
    const Scene& s = c->scene;
    if (!prev_cam) prev_cam = cam;
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < int(h); y++)
        for (uint32_t x = 0; x < w; x++) {
            const Scene::GBufferTexel g = s.gbuffer_pixel(*cam, *prev_cam, x, uint32_t(y), w, h, mask);
            const size_t i = size_t(y) * w + x;
            if (tex0) { tex0[2 * i] = g.tex0[0]; tex0[2 * i + 1] = g.tex0[1]; }
            if (tex1) ((PackedHitInfo*)tex1)[i] = pack_hit(g.hit);
            if (tex2) std::memcpy(tex2 + 4 * i, g.tex2, 16);
            if (tex3) std::memcpy(tex3 + 4 * i, g.tex3, 16);
            if (emission) std::memcpy(emission + 4 * i, g.emission, 16);
        }
    return 0;
}

int oracle_trace(oracle_context* c, const mpt_ray* rays, uint64_t n, uint32_t mask, mpt_hit* hits, float* t_out) {
    const Scene& s = c->scene;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < int64_t(n); i++) {
        Ray r{float3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]), rays[i].tmin,
              float3(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]), rays[i].tmax};
        float t;
        const Hit h = s.trace_ray(r, mask, &t);
        hits[i] = {h.instance_id, h.instance_index, h.primitive_id, h.geometry_index, {h.bary.x, h.bary.y}};
        if (t_out) t_out[i] = h.valid() ? t : -1.f;
    }
    return 0;
}

// render_pt.slang:92-130. tex1 == NULL: primary hits are traced here and round-tripped through
// PackedHitInfo (half2 barycentrics) exactly as the gbuffer -> pt hand-off does.
int oracle_render(oracle_context* c, const mpt_camera* cam, const mpt_render_params* P, const void* tex1, const void* albedo,
                  float* irradiance, uint64_t* rays_out) {
    const Scene& s = c->scene;
    const uint32_t w = P->width, h = P->height;
    std::atomic<uint64_t> total_rays{0};
#pragma omp parallel for schedule(dynamic, 2)
    for (int y = 0; y < int(h); y++) {
        uint64_t rays = 0;
        for (uint32_t x = 0; x < w; x++) {
            const size_t pix = size_t(y) * w + x;
            const Ray primary_ray = Scene::generate_camera_ray(*cam, x, uint32_t(y), w, h, P->frame);
            Hit primary_hit;
            if (tex1) primary_hit = unpack_hit(((const PackedHitInfo*)tex1)[pix]);
            else { primary_hit = unpack_hit(pack_hit(s.trace_ray(primary_ray, P->instance_mask))); rays++; }
            const ShadingPoint primary_sp = s.get_shading_point(primary_ray, primary_hit);
            float3 demod(1.f);
            if ((P->flags & MPT_RENDER_DEMODULATE_ALBEDO) && albedo) {
                const float* a = (const float*)albedo + pix * 4;
                demod = max3v(float3(a[0], a[1], a[2]), 1e-2f); // gbuffer.slang:146-148
            }
            float3 irr(0.f);
            float second_moment = 0.f;
            for (int i = 0; i < P->spp; i++) {
                const uint32_t sidx = P->sample_offset + uint32_t(i) * P->sample_stride;
                Rng rng = Rng::pcg(pcg4d16(x, uint32_t(y), P->frame, sidx));
                const float3 contrib = s.trace_path(primary_ray, primary_sp, rng, *P, &rays) / demod;
                if (!any_nan(contrib) && !any_inf(contrib)) {
                    irr += contrib;
                    const float l = yuv_luminance(contrib);
                    second_moment += l * l;
                }
            }
            if (P->spp > 0 && !(P->flags & MPT_RENDER_OUTPUT_SUM)) {
                irr /= float(P->spp);
                second_moment /= float(P->spp);
            }
            irradiance[pix * 4 + 0] = irr.x;
            irradiance[pix * 4 + 1] = irr.y;
            irradiance[pix * 4 + 2] = irr.z;
            irradiance[pix * 4 + 3] = second_moment;
        }
        total_rays += rays;
    }
    if (rays_out) *rays_out = total_rays.load();
    return 0;
}

// Accumulate node, plain path (accumulate.slang:103-255 with USE_MOTION_VECTORS = false, firefly filter and adaptive alpha
// off, GBUFFER_CHECK_NEVER or GBUFFER_CHECK_ALWAYS without EXTENDED_SEARCH). Texel formats: accumulate.slang:47-61.
int oracle_accumulate(const mpt_accumulate_params* A, const float* src_img, const float* prev_out_img, const uint32_t* prev_history,
                      float* out_img, uint32_t* history, const uint32_t* tex0, const float* tex2) {
    const size_t npix = size_t(A->width) * A->height;
    const bool f32_history = (A->flags & MPT_ACC_HISTORY_F32) != 0;
    for (size_t i = 0; i < npix; i++) {
        float4 src(src_img[4 * i], src_img[4 * i + 1], src_img[4 * i + 2], src_img[4 * i + 3]);
        const uint32_t enc_normal = tex0 ? tex0[2 * i] : 0u;                                  // in_gbuffer.get_encoded_normal, :109
        const float linear_z = std::min(tex2 ? tex2[4 * i + 2] : 0.f, 65504.0f);              // min(gbuf.linear_z, MAX_HISTORY_Z), :111
        const float delta_z = tex2 ? tex2[4 * i + 3] : 0.f;
        auto encode_history = [&](float length, uint32_t* o) {                                // :49-52
            o[0] = enc_normal;
            o[1] = f32_history ? as_uint(length) : (uint32_t(float_to_half(linear_z)) | (uint32_t(float_to_half(length)) << 16));
        };
        // firefly clamping, :120-138 (filter disabled: max_l = pc.firefly_hard_clamp)
        const float max_l = A->hard_clamp;
        const bool bad = std::isnan(src.x) || std::isnan(src.y) || std::isnan(src.z) || std::isnan(src.w) || std::isinf(src.x) || std::isinf(src.y) ||
                         std::isinf(src.z) || std::isinf(src.w);
        if (bad) {
            src = float4(0.f, 0.f, 0.f, 0.f);
        } else {
            const float l = yuv_luminance(float3(src.x, src.y, src.z));
            if (l > max_l) {
                const float factor = max_l / l;
                src.x *= factor; src.y *= factor; src.z *= factor;
                src.w *= factor * factor;
            }
        }
        float* o = out_img + 4 * i;
        if (A->clear == 1 || A->alpha == 0.f) {                                               // :141-146
            o[0] = src.x; o[1] = src.y; o[2] = src.z; o[3] = src.w;
            encode_history(1.0f, history + 2 * i);
            continue;
        }
        // no motion vectors: prev_ipos = ipos, is_static = true, :176-180
        const uint32_t ph0 = prev_history[2 * i], ph1 = prev_history[2 * i + 1];
        const float hist_length = f32_history ? as_float(ph1) : half_to_float(uint16_t(ph1 >> 16));
        bool found = true;                                                                    // skip_gbuffer_check, :186-187
        if (A->flags & MPT_ACC_GBUFFER_CHECK) {                                               // reprojection_valid, reprojection.slang:56-71
            const float3 n = decode_normal(enc_normal), pn = decode_normal(ph0);
            const float prev_z = half_to_float(uint16_t(ph1 & 0xFFFFu));
            found = (A->normal_reject_cos <= dot(n, pn)) && (std::fabs((linear_z - prev_z) + delta_z) <= A->z_reject_percent * std::max(linear_z, prev_z));
        }
        float4 prev_out(0.f, 0.f, 0.f, 0.f);
        float prev_history_len = 0.f;
        if (found) {
            prev_history_len = hist_length;
            prev_out = float4(prev_out_img[4 * i], prev_out_img[4 * i + 1], prev_out_img[4 * i + 2], prev_out_img[4 * i + 3]);
        }
        const bool prev_bad = std::isnan(prev_out.x) || std::isnan(prev_out.y) || std::isnan(prev_out.z) || std::isnan(prev_out.w) ||
                              std::isinf(prev_out.x) || std::isinf(prev_out.y) || std::isinf(prev_out.z) || std::isinf(prev_out.w);
        if (found && !prev_bad) {                                                             // :228-251
            const float new_history = std::min(prev_history_len + 1.0f, A->max_history);
            const float alpha = std::max(1.0f - A->alpha, 1.0f / new_history);
            o[0] = lerp(prev_out.x, src.x, alpha); o[1] = lerp(prev_out.y, src.y, alpha);
            o[2] = lerp(prev_out.z, src.z, alpha); o[3] = lerp(prev_out.w, src.w, alpha);
            encode_history(new_history, history + 2 * i);
        } else {
            o[0] = src.x; o[1] = src.y; o[2] = src.z; o[3] = src.w;
            encode_history(1.0f, history + 2 * i);
        }
    }
    return 0;
}

// Closest hits of `n` rays over the acceleration structure libmpt.so built (mpt_host_accel_get), with per-ray work counters summed
// into counts5 = {rays, node tests, triangle tests, instance entries, deepest stack}. `c` may be NULL when every geometry is opaque.
// skip_records (nullable): per ray, the triangle record it leaves (not tested, like the device's origin-triangle rule).
int oracle_bvh8_trace(oracle_context* c, const mpt_host_accel_view* accel, const mpt_ray* rays, uint64_t n, uint32_t mask, const uint32_t* skip_records,
                      mpt_hit* hits, float* t_out, uint32_t* records_out, uint64_t* counts5) {
    const Bvh8Walker W(*accel, c ? &c->scene : nullptr);
    uint64_t nodes = 0, tris = 0, entries = 0, max_stack = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : nodes, tris, entries) reduction(max : max_stack)
    for (int64_t i = 0; i < int64_t(n); i++) {
        const Ray r{float3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]), rays[i].tmin,
                    float3(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]), rays[i].tmax};
        Bvh8Counts cnt;
        float t;
        Bvh8Walker::State st;
        st.best_t = r.tmax;
        const uint32_t skip = skip_records ? skip_records[i] : 0xFFFFFFFFu;
        if (accel->two_level) W.walk(accel->root, true, r.o, r.d, r.tmin, mask, 0, 0xFFFFFFFFu, st, 0);
        else if ((W.inst[0].mask_flags & 0xFFu) & mask) W.walk(accel->root, false, r.o, r.d, r.tmin, mask, 0, skip, st, 0);
        cnt = st.cnt;
        t = st.best.valid() ? st.best_t : -1.f;
        const Hit& h = st.best;
        if (hits) hits[i] = {h.instance_id, h.instance_index, h.primitive_id, h.geometry_index, {h.bary.x, h.bary.y}};
        if (t_out) t_out[i] = t;
        if (records_out) records_out[i] = st.best_record;
        nodes += cnt.nodes; tris += cnt.tris; entries += cnt.instance_entries;
        max_stack = std::max(max_stack, cnt.max_stack);
    }
    if (counts5) { counts5[0] = n; counts5[1] = nodes; counts5[2] = tris; counts5[3] = entries; counts5[4] = max_stack; }
    return 0;
}

// bench.py --impl reference sets the thread count itself: under torchrun every rank inherits OMP_NUM_THREADS=1
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// ---- small-function hooks for unit tests -------------------------------------------------------
uint32_t oracle_pcg4d16(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return pcg4d16(x, y, z, w); }
void oracle_pcg_stream(uint32_t seed, uint32_t n, uint32_t* out_u, float* out_f) {
    Rng r = Rng::pcg(seed);
    for (uint32_t i = 0; i < n; i++) { const uint32_t u = r.next(); out_u[i] = u; out_f[i] = uint_to_zero_one_float_rust(u); }
}
uint32_t oracle_encode_normal(const float* n) { return encode_normal(float3(n[0], n[1], n[2])); }
void oracle_decode_normal(uint32_t e, float* out) { const float3 n = decode_normal(e); out[0] = n.x; out[1] = n.y; out[2] = n.z; }
uint32_t oracle_encode_tangent(const float* t) { return encode_tangent(float4(t[0], t[1], t[2], t[3])); }
uint16_t oracle_float_to_half(float f) { return float_to_half(f); }
float oracle_half_to_float(uint16_t h) { return half_to_float(h); }
void oracle_offset_ray_origin(const float* p, const float* n, float* out) {
    const float3 o = offset_ray_origin(float3(p[0], p[1], p[2]), float3(n[0], n[1], n[2]));
    out[0] = o.x; out[1] = o.y; out[2] = o.z;
}
void oracle_set_scene_frame(oracle_context* c, uint32_t frame) { c->scene.frame = frame; }
// Scene::prev_instance_transforms (scene.slang:38): instance_count x 16 floats, NULL = static scene
void oracle_scene_set_prev_transforms(oracle_context* c, const float* t, uint32_t instance_count) {
    c->scene.prev_instance_transforms.clear();
    if (t) c->scene.prev_instance_transforms.assign(t, t + size_t(instance_count) * 16);
}
void oracle_camera_ray_frame(const mpt_camera* cam, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint32_t frame, mpt_ray* out) {
    const Ray r = Scene::generate_camera_ray(*cam, x, y, w, h, frame);
    *out = {{r.o.x, r.o.y, r.o.z}, r.tmin, {r.d.x, r.d.y, r.d.z}, r.tmax};
}
void oracle_camera_ray(const mpt_camera* cam, uint32_t x, uint32_t y, uint32_t w, uint32_t h, mpt_ray* out) {
    const Ray r = Scene::generate_camera_ray(*cam, x, y, w, h);
    *out = {{r.o.x, r.o.y, r.o.z}, r.tmin, {r.d.x, r.d.y, r.d.z}, r.tmax};
}
void oracle_sample_texture(oracle_context* c, uint32_t id, float u, float v, float* out) {
    const float4 t = sample_texture(c->scene.textures[id], float2(u, v));
    out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
}
void oracle_env_eval(oracle_context* c, const float* dir, float* out) {
    const float3 e = c->scene.env_eval(float3(dir[0], dir[1], dir[2]));
    out[0] = e.x; out[1] = e.y; out[2] = e.z;
}
// intersect one triangle (for the facing / watertightness unit tests). cull: 0 back, 1 front, 2 none
int oracle_intersect_triangle(const float* o, const float* d, const float* v0, const float* v1, const float* v2, int cull, float tmin,
                              float tmax, float* tuv) {
    const RayShear rs(float3(d[0], d[1], d[2]));
    return intersect_triangle(rs, float3(o[0], o[1], o[2]), float3(v0[0], v0[1], v0[2]), float3(v1[0], v1[1], v1[2]),
                              float3(v2[0], v2[1], v2[2]), cull, tmin, tmax, tuv[0], tuv[1], tuv[2])
               ? 1 : 0;
}

// tests/material_system/read_material.slang + sample_material.slang restated (tests/test_material_system.cpp:107-215):
// header words, first four payload floats, and MaterialSample::eval(wi, wo) in the frame N=(0,0,1), T=(1,0,0).
void oracle_material_read(oracle_context* c, uint32_t material_id, uint32_t* out7) {
    const uint8_t* e = c->scene.entry(material_id);
    out7[0] = Scene::rd16(e);
    out7[1] = Scene::rd16(e + 2);
    for (int i = 0; i < 4; i++) std::memcpy(&out7[2 + i], e + 4 + 4 * i, 4);
    out7[6] = c->scene.material_count;
}
void oracle_material_eval(oracle_context* c, uint32_t material_id, const float* wi, const float* wo, const float* uv, float* out3) {
    ShadingData sd{float3(wi[0], wi[1], wi[2]), float3(0.f), float2(uv[0], uv[1]), Frame(float3(0, 0, 1), float3(1, 0, 0)), float3(0, 0, 1), 0, 1.0f, true};
    float3 emission;
    Frame fr;
    const MaterialSample ms = c->scene.get_material_sample(material_id, sd, emission, fr);
    Rng rng = Rng::pcg(1u);
    const float3 f = ms.eval(fr.to_local(sd.wi), fr.to_local(float3(wo[0], wo[1], wo[2])), rng);
    out3[0] = f.x; out3[1] = f.y; out3[2] = f.z;
}

// BSDF ids as in mpt_bsdf_check_params; configs as in tests/test_bsdf.cpp:62-160.
int oracle_bsdf_check(const mpt_bsdf_check_params* p, float* out16) {
    const float3 wi(p->wi[0], p->wi[1], p->wi[2]);
    const float3 albedo(p->albedo[0], p->albedo[1], p->albedo[2]);
    const float3 f0(p->f0[0], p->f0[1], p->f0[2]);
    switch (p->bsdf) {
    case 0: bsdf_checks(LambertDiffuseBRDF(albedo), wi, p->n_samples, p->seed, out16); break;
    case 1: bsdf_checks(GGXBRDF(p->alpha), wi, p->n_samples, p->seed, out16); break;
    case 2: bsdf_checks(GGXBRDF(float2(p->alpha, p->alpha_bitangent)), wi, p->n_samples, p->seed, out16); break;
    case 3: bsdf_checks(FresnelMixBSDF<GGXBRDF, LambertDiffuseBRDF>(GGXBRDF(p->alpha), LambertDiffuseBRDF(albedo), p->ior_f0), wi, p->n_samples, p->seed, out16); break;
    case 4: bsdf_checks(IridescentFresnelMixBSDF<GGXBRDF, LambertDiffuseBRDF>(GGXBRDF(p->alpha), LambertDiffuseBRDF(albedo), p->ior_f0, 1.0f / 1.5f, 1.3f, 400.0f, p->iridescence), wi, p->n_samples, p->seed, out16); break;
    case 5: bsdf_checks(IridescentRoughDielectricBSDF(p->alpha, 1.0f / 1.5f, albedo, float3(0.04f), 1.3f, 400.0f, p->iridescence), wi, p->n_samples, p->seed, out16); break;
    case 6: bsdf_checks(SheenBRDF(albedo, p->alpha), wi, p->n_samples, p->seed, out16); break;
    case 8: bsdf_checks(ConductorFresnelBSDF<GGXBRDF>(GGXBRDF(p->alpha), f0), wi, p->n_samples, p->seed, out16); break;
    case 7:
    case 9: {
        // the nest GltfMaterial::get_sample builds (gltf-material.slang:18-27, 153-308), with the factors mpt.h states for these codes
        const bool full = p->bsdf == 7;
        const float eta = 1.0f / 1.5f, iri = full ? p->iridescence : 0.f, metallic = full ? 0.5f : 0.f, coat_w = full ? 0.5f : 0.f;
        const float2 spec_alpha(p->alpha, full && p->alpha_bitangent > 0.f ? p->alpha_bitangent : p->alpha);
        const float3 sheen_col = full ? albedo * 0.5f : albedo;
        const GltfOpaqueDielectric od(SpecularBRDF(spec_alpha), LambertDiffuseBRDF(albedo), p->ior_f0, eta, 1.3f, 400.0f, iri);
        const GltfDielectric diel(od, IridescentRoughDielectricBSDF(p->alpha, eta, albedo, float3(p->ior_f0), 1.3f, 400.0f, iri), 0.f);
        const GltfReflective refl(diel, GltfConductor(SpecularBRDF(spec_alpha), albedo, 1.3f, 400.0f, iri), metallic);
        const float max_sheen = std::max(sheen_col.x, std::max(sheen_col.y, sheen_col.z));
        const GltfSheened sheened(SheenBRDF(sheen_col, p->alpha), refl, p->alpha, max_sheen);
        const float coat_alpha = DEFAULT_MIN_ROUGHNESS * DEFAULT_MIN_ROUGHNESS;
        bsdf_checks(GltfBSDF(sheened, GltfCoated(SpecularBRDF(coat_alpha), sheened, GLTF_DIELECTRIC_IOR_F0), coat_w), wi, p->n_samples, p->seed, out16);
        break;
    }
    default: return 1;
    }
    return 0;
}

} // extern "C"
