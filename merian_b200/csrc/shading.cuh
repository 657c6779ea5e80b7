// shading.cuh — device scene tables and the shading-point / material-sample evaluation.
//
// Replaces Scene::get_triangle_hit / get_surface / get_shading_point / _make_environment_map_shading_point
// (include/merian-shaders/scene/scene.slang:312-382, 418-498), MaterialSystem::get_material_sample
// (shading/materials/material-system.slang:51-72) with DiffuseMaterial (diffuse-material.slang:19-34) and
// GltfMaterial (gltf-material.slang:153-308), LatLongMap::eval (scene/environment-map.slang:23-35),
// ShadingPoint::make_ray (scene/hit.slang:134-147).
//
// Data layout in HBM (all 16-byte aligned, read with 128-bit loads):
//   tris      3 x float4 per triangle record  v0|primitive_id, v1|geometry_index, v2|flags      (shared with traversal)
//   tri_attr  3 x uint4  per triangle record  per corner: oct normal, half2 uv, oct tangent     (9 words + 3 pad)
//   geometry  1 x uint4  per GeometryID       material_id|flags<<16, primitive_count, prim->tri offset, -
// so a shaded hit costs two coalesced 48-byte records instead of the reference's index triple + 3 scattered
// 24-byte PackedVertexData gathers (scene.slang:158-173,312-320).
#pragma once

#include "bsdf.cuh"
#include "bvh8_traverse.cuh"

namespace mpt {

enum : uint32_t { // GeometryDataFlags, scene-data.slangh:15-42
    G_PRETRANSFORMED = 0x1, G_IS_ANIMATED = 0x4, G_FLIP_FACING = 0x8, G_FLAT_SHADING = 0x80, G_HAS_TANGENTS = 0x100, G_USE_ENV_MAP = 0x200, G_TWO_SIDED = 0x400
};

struct TextureRec {
    cudaTextureObject_t obj;
    uint32_t valid;
    uint32_t pad;
};

struct SceneView {
    AccelView accel;
    const uint4* __restrict__ tri_attr;
    const uint4* __restrict__ geometries;
    const uint32_t* __restrict__ prim_to_tri;
    const float4* __restrict__ inst_xform;    // 4 float4 per instance: rows of object->world
    const float4* __restrict__ inst_xform_it; // rows of the inverse transpose
    const float4* __restrict__ inst_xform_prev; // object->world of the previous frame (mpt_scene_set_prev_transforms), nullptr = static scene
    const uint32_t* __restrict__ materials;   // material buffer as words
    uint32_t material_stride_words;
    const uint2* __restrict__ material_types; // (model, features)
    const TextureRec* __restrict__ textures;
    uint32_t texture_count;
    // environment
    uint32_t env_type;
    float env_m[9]; // (float3x3)to_local, row-major
    float env_intensity;
    const float4* __restrict__ env_pixels;
    uint32_t env_w, env_h;
    // environment importance sampling for MPT_RENDER_NEE (built lazily by libmpt; no reference counterpart)
    const uint2* __restrict__ env_alias;   // per texel: x = probability of keeping the texel (float bits), y = alias texel
    const float* __restrict__ env_pmf;     // per texel: probability mass (sums to 1)
    float env_inv[9];                      // inverse of env_m: local -> world
    // MaterialSystem hints (material_system.hpp:83-97)
    float alpha_threshold, min_roughness;
    uint32_t clamp_normals;
};

constexpr float kRayTMax = 1e7f;           // SCENE_RAY_TMAX, hit.slang:8
constexpr float kEnvDistance = 1e7f;       // ENV_MAP_DISTANCE, environment-map.slang:11
constexpr float kThroughputCut = 1e-5f;    // SCENE_THROUGHPUT_THRESHOLD, scene.slang:24
constexpr float kMinRoughness = 0.0316f;   // merian_hint_min_roughness, material.slang:23
constexpr float kAlphaThreshold = 0.5f;    // merian_alpha_test_threshold, material-system.slang:7
constexpr uint32_t kNoTexture = 0xFFFFu;

struct ShadingPoint { // hit.slang:105-147
    f3 pos;
    f3 face_normal_dec; // decode_normal(encode_normal(surface.face_normal)): what get_face_normal() returns
    f3 geo_normal;      // surface.face_normal itself (GBuffer grad_z, hit.slang:99-103); dead in the pt kernels
    f3 emission;
    Frame frame;
    bool has_material;
    Bsdf bsdf;
    f3 absorption; // MaterialSample interior volume (gltf-material.slang:322-324)
};

// ---------------------------------------------------------------- textures
MPT_DEV f4 sample_tex(const SceneView& S, uint32_t id, f2 uv) { // LOD0Sampler::sample, textures.slang:338-366
    const float4 t = tex2D<float4>(S.textures[id].obj, uv.x, uv.y);
    return mk4(t.x, t.y, t.z, t.w);
}

MPT_DEV int wrap_repeat(int i, int n) {
    int m = i % n;
    return m < 0 ? m + n : m;
}
// RGBA32F lat-long map, repeat addressing, bilinear with full fp32 weights (scene.cpp:786-789)
MPT_DEV f3 env_eval(const SceneView& S, f3 dir) {
    if (S.env_type != 1u) return mk3(0.f);
    const f3 d = mk3(S.env_m[0] * dir.x + S.env_m[1] * dir.y + S.env_m[2] * dir.z, S.env_m[3] * dir.x + S.env_m[4] * dir.y + S.env_m[5] * dir.z,
                     S.env_m[6] * dir.x + S.env_m[7] * dir.y + S.env_m[8] * dir.z);
    const f2 uv = world_to_latlong(d);
    const float fu = uv.x * float(S.env_w) - 0.5f, fv = uv.y * float(S.env_h) - 0.5f;
    const float flu = floorf(fu), flv = floorf(fv);
    const float a = fu - flu, b = fv - flv;
    const int x0 = wrap_repeat(int(flu), int(S.env_w)), x1 = wrap_repeat(int(flu) + 1, int(S.env_w));
    const int y0 = wrap_repeat(int(flv), int(S.env_h)), y1 = wrap_repeat(int(flv) + 1, int(S.env_h));
    const float4 t00 = __ldg(S.env_pixels + size_t(y0) * S.env_w + x0), t10 = __ldg(S.env_pixels + size_t(y0) * S.env_w + x1);
    const float4 t01 = __ldg(S.env_pixels + size_t(y1) * S.env_w + x0), t11 = __ldg(S.env_pixels + size_t(y1) * S.env_w + x1);
    const float w00 = (1 - a) * (1 - b), w10 = a * (1 - b), w01 = (1 - a) * b, w11 = a * b;
    return S.env_intensity * mk3(w00 * t00.x + w10 * t10.x + w01 * t01.x + w11 * t11.x, w00 * t00.y + w10 * t10.y + w01 * t01.y + w11 * t11.y,
                                 w00 * t00.z + w10 * t10.z + w01 * t01.z + w11 * t11.z);
}

// ---- environment importance sampling (MPT_RENDER_NEE). Texel (i, j) of the lat-long map covers u in [i, i+1)/W, v in [j, j+1)/H
// with uv = world_to_latlong(local dir); solid angle of d(u, v) = 2 pi^2 sin(theta) du dv.
MPT_DEV float env_pdf(const SceneView& S, f3 dir_world) {
    const f3 d = normalize(mk3(S.env_m[0] * dir_world.x + S.env_m[1] * dir_world.y + S.env_m[2] * dir_world.z,
                               S.env_m[3] * dir_world.x + S.env_m[4] * dir_world.y + S.env_m[5] * dir_world.z,
                               S.env_m[6] * dir_world.x + S.env_m[7] * dir_world.y + S.env_m[8] * dir_world.z));
    const float sin_theta = sqrtf(fmaxf(0.f, 1.0f - d.y * d.y));
    if (sin_theta < 1e-6f) return 0.f;
    const f2 uv = mk2(atan2f(d.x, -d.z) * kInvTwoPi + 0.5f, acosf(clampf(d.y, -1.f, 1.f)) * kInvPi);
    const uint32_t i = min(uint32_t(fmaxf(uv.x, 0.f) * float(S.env_w)), S.env_w - 1u), j = min(uint32_t(fmaxf(uv.y, 0.f) * float(S.env_h)), S.env_h - 1u);
    return __ldg(S.env_pmf + size_t(j) * S.env_w + i) * float(S.env_w) * float(S.env_h) / (2.0f * kPi * kPi * sin_theta);
}
// Alias-table draw of a texel, uniform position inside it. Returns false when the direction has no usable density.
MPT_DEV bool env_sample(const SceneView& S, float u0, float u1, float u2, f3& dir_world, float& pdf) {
    const uint32_t n = S.env_w * S.env_h;
    const float scaled = u0 * float(n);
    uint32_t t = min(uint32_t(scaled), n - 1u);
    const uint2 e = __ldg(S.env_alias + t);
    if (scaled - float(t) >= __uint_as_float(e.x)) t = e.y;
    const uint32_t j = t / S.env_w, i = t - j * S.env_w;
    const float u = (float(i) + u1) / float(S.env_w), v = (float(j) + u2) / float(S.env_h);
    const float phi = (u - 0.5f) * kTwoPi, theta = v * kPi;
    const float st = sinf(theta), ct = cosf(theta);
    if (st < 1e-6f) return false;
    const f3 l = mk3(st * sinf(phi), ct, -st * cosf(phi)); // inverse of world_to_latlong
    dir_world = normalize(mk3(S.env_inv[0] * l.x + S.env_inv[1] * l.y + S.env_inv[2] * l.z, S.env_inv[3] * l.x + S.env_inv[4] * l.y + S.env_inv[5] * l.z,
                              S.env_inv[6] * l.x + S.env_inv[7] * l.y + S.env_inv[8] * l.z));
    pdf = __ldg(S.env_pmf + t) * float(n) / (2.0f * kPi * kPi * st);
    return pdf > 0.f;
}

// ---------------------------------------------------------------- materials
MPT_DEV float wordf(const uint32_t* p, int i) { return __uint_as_float(__ldg(p + i)); }
MPT_DEV uint32_t lo16(uint32_t w) { return w & 0xFFFFu; }
MPT_DEV uint32_t hi16(uint32_t w) { return w >> 16; }

// keller_clamp_normal, utils/normal-map.slang:5-11
MPT_DEV f3 keller_clamp(f3 ns, f3 ng, f3 wi) {
    const f3 r = (-wi) - 2.0f * dot(ns, -wi) * ns;
    if (dot(r, ng) < 0.f) return normalize(wi + normalize(r - ng * dot(ng, r)));
    return ns;
}

struct ShadingData { // shading-data.slang:27-58 (only the members the supported models read)
    f3 wi;
    f2 uv;
    Frame frame;
    f3 face_normal;
    float ior;
    bool front_facing;
    // the LODSampler2D the caller passes to get_material_sample: LOD0Sampler (pt: use_grad = false, a compile-time constant there) or
    // GradLODSampler(ddx, ddy) (the GBuffer node, textures.slang:398-421)
    bool use_grad;
    f2 ddx, ddy;
};
MPT_DEV f4 sample_tex(const SceneView& S, uint32_t id, const ShadingData& sd) {
    if (sd.use_grad) { // Sampler2D::SampleGrad; a texture without a mip chain has one level, so the gradients change nothing
        const float4 t = tex2DGrad<float4>(S.textures[id].obj, sd.uv.x, sd.uv.y, make_float2(sd.ddx.x, sd.ddx.y), make_float2(sd.ddy.x, sd.ddy.y));
        return mk4(t.x, t.y, t.z, t.w);
    }
    return sample_tex(S, id, sd.uv);
}

// Emission only: MaterialModel::get_emission (material.slang:40-47; gltf-material.slang:142-151; diffuse: 0).
MPT_DEV f3 material_emission(const SceneView& S, uint32_t material_id, f2 uv) {
    const uint32_t* e = S.materials + size_t(material_id) * S.material_stride_words;
    const uint32_t hdr = __ldg(e);
    const uint2 mt = __ldg(S.material_types + lo16(hdr));
    if (mt.x == 0u) return mk3(0.f);
    const uint32_t* p = e + 1;
    f3 em = mk3(wordf(p, 4), wordf(p, 5), wordf(p, 6));
    const uint32_t em_tex = hi16(__ldg(p + 13));
    if (em_tex != kNoTexture) {
        const f4 t = sample_tex(S, em_tex, uv);
        em = em * mk3(t.x, t.y, t.z);
    }
    return em;
}

// Full material sample. FEAT as in bsdf.cuh.
template <uint32_t FEAT>
MPT_DEV void material_sample(const SceneView& S, uint32_t material_id, const ShadingData& sd, ShadingPoint& sp) {
    const uint32_t* e = S.materials + size_t(material_id) * S.material_stride_words;
    const uint32_t hdr = __ldg(e);
    const uint2 mt = __ldg(S.material_types + lo16(hdr));
    const uint32_t alpha_tex = hi16(hdr);
    const uint32_t* p = e + 1;
    Bsdf& b = sp.bsdf;
    sp.absorption = mk3(0.f);
    f4 base = mk4(wordf(p, 0), wordf(p, 1), wordf(p, 2), wordf(p, 3));
    if (mt.x == 0u) { // DiffuseMaterial, diffuse-material.slang:19-34
        sp.emission = mk3(0.f);
        sp.frame = frame_apply_normal_map(sd.frame, mk3(0.f, 0.f, 1.f));
        f3 albedo = mk3(base.x, base.y, base.z);
        if (alpha_tex != kNoTexture) {
            const f4 t = sample_tex(S, alpha_tex, sd);
            albedo = albedo * mk3(t.x, t.y, t.z);
        }
        b.gltf = 0u;
        b.base = albedo;
        return;
    }
    // ---- GltfMaterial::get_sample, gltf-material.slang:153-308; payload layout gltf_material.hpp:102-112
    const uint32_t feat = mt.y & FEAT;
    if (alpha_tex != kNoTexture) {
        const f4 t = sample_tex(S, alpha_tex, sd);
        base = mk4(base.x * t.x, base.y * t.y, base.z * t.z, base.w * t.w);
    }
    f3 emission = mk3(wordf(p, 4), wordf(p, 5), wordf(p, 6));
    float metallic = wordf(p, 7), roughness = wordf(p, 8);
    const float normal_scale = wordf(p, 9);
    const float ior = wordf(p, 11);
    const uint32_t t01 = __ldg(p + 12), t23 = __ldg(p + 13);
    const uint32_t mr_tex = lo16(t01), normal_tex = hi16(t01), em_tex = hi16(t23);
    if (mr_tex != kNoTexture) {
        const f4 mr = sample_tex(S, mr_tex, sd);
        roughness *= mr.y;
        metallic *= mr.z;
    }
    Frame fr = sd.frame;
    if (normal_tex != kNoTexture) {
        const f4 raw = sample_tex(S, normal_tex, sd);
        const float nx = (raw.x * 2.0f - 1.0f) * normal_scale, ny = (raw.y * 2.0f - 1.0f) * normal_scale;
        const float nz = sqrtf(fmaxf(0.0f, 1.0f - (nx * nx + ny * ny)));
        const Frame mapped = frame_apply_normal_map(sd.frame, mk3(nx, ny, nz));
        const f3 cn = S.clamp_normals ? keller_clamp(mapped.n, sd.face_normal, sd.wi) : mapped.n; // merian_hint_clamp_normals
        fr = frame_from_tangent(cn, mapped.t, 1.0f);
    }
    if (em_tex != kNoTexture) {
        const f4 t = sample_tex(S, em_tex, sd);
        emission = emission * mk3(t.x, t.y, t.z);
    }
    const float alpha = sq(fmaxf(roughness, S.min_roughness));
    b.gltf = 1u;
    b.base = mk3(base.x, base.y, base.z);
    b.ax = b.ay = b.alpha = alpha;
    b.metallic = metallic;
    const float ior_r = (ior - 1.0f) / (ior + 1.0f);
    b.f0 = ior_r * ior_r;
    b.eta = sd.ior / (sd.front_facing ? ior : 1.0f);
    b.tw = 0.f;
    b.coat_w = 0.f;
    b.coat_alpha = sq(S.min_roughness);
    b.sheen_col = mk3(0.f);
    b.sheen_alpha = 0.f;
    b.max_sheen = 0.f;
    b.iri = 0.f;
    b.iri_ior = 1.0f;
    b.iri_thick = 0.f;
    if (FEAT != 0u) {
        // optional extension payloads follow the core fields in declaration order (gltf_material.hpp:114-145);
        // byte offsets are multiples of 2, so read 16-bit granules.
        const uint16_t* q16 = reinterpret_cast<const uint16_t*>(p) + 28;
        auto rdf = [&](int byte_off) {
            const uint32_t lo = q16[byte_off / 2], hi = q16[byte_off / 2 + 1];
            return __uint_as_float(lo | (hi << 16));
        };
        auto rd16 = [&](int byte_off) { return uint32_t(q16[byte_off / 2]); };
        int off = 0;
        if (mt.y & F_TRANSMISSION) { if (feat & F_TRANSMISSION) b.tw = rdf(off); off += 4; }
        if (mt.y & F_VOLUME) {
            if ((feat & F_VOLUME) && sd.front_facing) sp.absorption = mk3(rdf(off), rdf(off + 4), rdf(off + 8));
            off += 12;
        }
        if (mt.y & F_CLEARCOAT) {
            if (feat & F_CLEARCOAT) {
                b.coat_w = rdf(off);
                float cr = rdf(off + 4);
                const uint32_t ct = rd16(off + 8), crt = rd16(off + 10);
                if (ct != kNoTexture) b.coat_w *= sample_tex(S, ct, sd).x;
                if (crt != kNoTexture) cr *= sample_tex(S, crt, sd).y;
                b.coat_alpha = sq(fmaxf(cr, S.min_roughness));
            }
            off += 12;
        }
        if (mt.y & F_SHEEN) {
            if (feat & F_SHEEN) {
                f3 col = mk3(rdf(off), rdf(off + 4), rdf(off + 8));
                float sr = rdf(off + 12);
                const uint32_t sct = rd16(off + 16), srt = rd16(off + 18);
                if (sct != kNoTexture) { const f4 t = sample_tex(S, sct, sd); col = col * mk3(t.x, t.y, t.z); }
                if (srt != kNoTexture) sr *= sample_tex(S, srt, sd).w;
                b.sheen_col = col;
                b.sheen_alpha = sq(sr);
                b.max_sheen = max3(col);
            }
            off += 20;
        }
        if (mt.y & F_IRIDESCENCE) {
            if (feat & F_IRIDESCENCE) {
                b.iri = rdf(off);
                b.iri_ior = rdf(off + 4);
                const float tmin = rdf(off + 8), tmax = rdf(off + 12);
                const uint32_t it = rd16(off + 16), itt = rd16(off + 18);
                if (it != kNoTexture) b.iri *= sample_tex(S, it, sd).x;
                float t = 1.0f;
                if (itt != kNoTexture) t = sample_tex(S, itt, sd).y;
                b.iri_thick = mix(tmin, tmax, t);
            }
            off += 20;
        }
        if (mt.y & F_ANISOTROPY) {
            if (feat & F_ANISOTROPY) {
                float strength = rdf(off);
                const float rot = rdf(off + 4);
                const uint32_t at = rd16(off + 8);
                float dx = cosf(rot), dy = sinf(rot);
                if (at != kNoTexture) {
                    const f4 a = sample_tex(S, at, sd);
                    float ex = a.x * 2.0f - 1.0f, ey = a.y * 2.0f - 1.0f;
                    const float l = sqrtf(ex * ex + ey * ey);
                    ex /= l;
                    ey /= l;
                    const float ndx = dx * ex - dy * ey, ndy = dx * ey + dy * ex;
                    dx = ndx;
                    dy = ndy;
                    strength *= a.z;
                }
                if (strength > 0.0f) {
                    b.ax = mix(alpha, 1.0f, strength * strength);
                    const f3 t = dx * fr.t + dy * fr.b;
                    fr = frame_from_tangent(fr.n, t, 1.0f);
                }
            }
            off += 10;
        }
    }
    sp.frame = fr;
    sp.emission = emission;
}

// ---------------------------------------------------------------- shading point
MPT_DEV f3 xf_point(const float4* m, f3 p) {
    const float4 r0 = __ldg(m), r1 = __ldg(m + 1), r2 = __ldg(m + 2);
    return mk3(r0.x * p.x + r0.y * p.y + r0.z * p.z + r0.w, r1.x * p.x + r1.y * p.y + r1.z * p.z + r1.w, r2.x * p.x + r2.y * p.y + r2.z * p.z + r2.w);
}
MPT_DEV f3 xf_vec(const float4* m, f3 v) {
    const float4 r0 = __ldg(m), r1 = __ldg(m + 1), r2 = __ldg(m + 2);
    return mk3(r0.x * v.x + r0.y * v.y + r0.z * v.z, r1.x * v.x + r1.y * v.y + r1.z * v.z, r2.x * v.x + r2.y * v.y + r2.z * v.z);
}

// scene.slang:482-498
MPT_DEV void env_shading_point(const SceneView& S, f3 ray_o, f3 ray_d, ShadingPoint& sp) {
    sp.pos = ray_o + ray_d * kEnvDistance;
    sp.face_normal_dec = -ray_d; // never read: has_material is false
    sp.geo_normal = -ray_d;
    sp.frame = frame_from_normal(-ray_d); // OrthonormalFrame(-ray.Direction), scene.slang:492
    sp.emission = env_eval(S, ray_d);
    sp.has_material = false;
}

// scene.slang:418-469 with get_surface (:322-382) inlined. `hit.tri` must be valid.
template <uint32_t FEAT>
// diff_dirs (GBuffer node only): the camera ray directions of the pixels to the right and below; the material is then sampled with the
// screen-space UV gradients of the ray differential (gbuffer_rt.slang:41-50, ray-differential.slang) instead of at LOD 0.
MPT_DEV void surface_shading_point(const SceneView& S, f3 ray_o, f3 ray_d, const HitRec& hit, float ctx_ior, bool with_material, ShadingPoint& sp,
                                   const f3* diff_dirs = nullptr) {
    const float4* tp = S.accel.tris + size_t(hit.tri) * 3;
    const float4 ta = __ldg(tp), tb = __ldg(tp + 1), tc = __ldg(tp + 2);
    const uint32_t geometry_index = __float_as_uint(tb.w);
    const uint32_t instance_id = S.accel.instances[hit.instance].custom_index;
    const uint4 gd = __ldg(S.geometries + instance_id + geometry_index);
    const uint32_t gflags = gd.x >> 16, material_id = gd.x & 0xFFFFu;
    if (gflags & G_USE_ENV_MAP) { env_shading_point(S, ray_o, ray_d, sp); return; }

    const uint4* ap = S.tri_attr + size_t(hit.tri) * 3;
    const uint4 a0 = __ldg(ap), a1 = __ldg(ap + 1), a2 = __ldg(ap + 2);
    const f3 v0 = xyz(ta), v1 = xyz(tb), v2 = xyz(tc);
    const float bx = 1.0f - hit.u - hit.v, by = hit.u, bz = hit.v; // Hit::get_barycentrics3, hit.slang:58-60

    const bool pre = gflags & G_PRETRANSFORMED;
    const float4* M = S.inst_xform + size_t(hit.instance) * 4;
    const f3 p_obj = bx * v0 + by * v1 + bz * v2;
    const f3 pos = pre ? p_obj : xf_point(M, p_obj);
    const f3 e01 = pre ? (v1 - v0) : xf_vec(M, v1 - v0);
    const f3 e02 = pre ? (v2 - v0) : xf_vec(M, v2 - v0);
    f3 face_normal = normalize(cross(e01, e02));
    if (gflags & G_FLIP_FACING) face_normal = -face_normal;

    f3 vertex_normal;
    if (gflags & G_FLAT_SHADING) {
        vertex_normal = face_normal;
    } else {
        const f3 n_obj = normalize(bx * decode_normal(a0.x) + by * decode_normal(a0.w) + bz * decode_normal(a1.z));
        vertex_normal = pre ? n_obj : normalize(xf_vec(S.inst_xform_it + size_t(hit.instance) * 4, n_obj));
    }
    const f2 uv0 = half2_bits_to_f2(a0.y), uv1 = half2_bits_to_f2(a1.x), uv2 = half2_bits_to_f2(a1.w);
    Frame frame;
    if (gflags & G_HAS_TANGENTS) {
        const f4 t0 = decode_tangent(a0.z), t1 = decode_tangent(a1.y), t2 = decode_tangent(a2.x);
        const f3 t = mk3(bx * t0.x + by * t1.x + bz * t2.x, bx * t0.y + by * t1.y + bz * t2.y, bx * t0.z + by * t1.z + bz * t2.z);
        const float tw = bx * t0.w + by * t1.w + bz * t2.w;
        const f3 tn = normalize(t);
        frame = frame_from_tangent(vertex_normal, pre ? tn : xf_vec(M, tn), sgn(tw));
    } else {
        const f2 duv1 = uv1 - uv0, duv2 = uv2 - uv0;
        const float det = duv1.x * duv2.y - duv2.x * duv1.y;
        if (fabsf(det) > 1e-8f) {
            const float winding = (gflags & G_FLIP_FACING) ? 1.f : -1.f;
            const f3 tangent = (duv2.y * e01 - duv1.y * e02) / det;
            frame = frame_from_tangent(vertex_normal, tangent, sgn(det) * winding);
        } else {
            frame = frame_from_normal(vertex_normal);
        }
    }
    const f2 uv = bx * uv0 + by * uv1 + bz * uv2;

    f3 fn = face_normal;
    bool front_facing = true;
    if (gflags & G_TWO_SIDED) {
        front_facing = dot(-ray_d, face_normal) >= 0.f;
        if (!front_facing) {
            frame.n = -frame.n; // OrthonormalFrame::flip_normal, frame.slang:122-125
            fn = -fn;
        }
    }
    sp.pos = pos;
    sp.geo_normal = face_normal;
    sp.face_normal_dec = decode_normal(encode_normal(face_normal)); // sp.enc_face_normal round trip, scene.slang:443
    sp.has_material = with_material;
    if (with_material) {
        ShadingData sd;
        sd.wi = -ray_d;
        sd.uv = uv;
        sd.frame = frame;
        sd.face_normal = fn;
        sd.ior = ctx_ior;
        sd.front_facing = front_facing;
        sd.use_grad = false;
        sd.ddx = sd.ddy = mk2(0.f, 0.f);
        if (diff_dirs) {
            // RayDifferential::from_pinhole_camera + propagate (Igehy eq. 8, 10, 12) + compute_uv_gradients, ray-differential.slang:13-62
            const float t = length(pos - ray_o);
            f3 dOdx = t * (diff_dirs[0] - ray_d), dOdy = t * (diff_dirs[1] - ray_d);
            const float rcp_dn = 1.0f / dot(ray_d, face_normal);
            dOdx = dOdx + ray_d * (-dot(dOdx, face_normal) * rcp_dn);
            dOdy = dOdy + ray_d * (-dot(dOdy, face_normal) * rcp_dn);
            const f3 Nu = cross(e02, face_normal), Nv = cross(e01, face_normal);
            const f3 Lu = Nu / dot(Nu, e01), Lv = Nv / dot(Nv, e02);
            const f2 d1 = uv1 - uv0, d2 = uv2 - uv0;
            sd.ddx = dot(Lu, dOdx) * d1 + dot(Lv, dOdx) * d2;
            sd.ddy = dot(Lu, dOdy) * d1 + dot(Lv, dOdy) * d2;
            sd.use_grad = true;
        }
        material_sample<FEAT>(S, material_id, sd, sp);
    } else {
        sp.emission = material_emission(S, material_id, uv);
        sp.frame = frame;
    }
}

// Alpha test for non-opaque candidates (scene.slang:503-516, material-system.slang:28-38).
struct MaterialAlphaTest {
    const SceneView* S;
    MPT_DEV bool operator()(uint32_t tri, uint32_t instance, float u, float v) const {
        const float4 tb = __ldg(S->accel.tris + size_t(tri) * 3 + 1);
        const uint32_t gid = S->accel.instances[instance].custom_index + __float_as_uint(tb.w);
        const uint32_t material_id = __ldg(S->geometries + gid).x & 0xFFFFu;
        const uint32_t hdr = __ldg(S->materials + size_t(material_id) * S->material_stride_words);
        const uint32_t alpha_tex = hi16(hdr);
        if (alpha_tex == kNoTexture) return false;
        const uint4* ap = S->tri_attr + size_t(tri) * 3;
        const uint4 a0 = __ldg(ap), a1 = __ldg(ap + 1);
        const f2 uv = (1.0f - u - v) * half2_bits_to_f2(a0.y) + u * half2_bits_to_f2(a1.x) + v * half2_bits_to_f2(a1.w);
        return sample_tex(*S, alpha_tex, uv).w < S->alpha_threshold;
    }
};

// camera.slang:29-56 (pinhole)
struct CameraParams {
    f3 position, U, V, W;
    float near_plane, far_plane;
    float jitter_x, jitter_y;
    float aperture_radius; // > 0: thin lens (camera.slang:73-92), the hint merian_hint_enable_thin_lens follows the camera (scene.cpp:2019-2021)
    uint32_t frame;        // Scene::frame: the lens sample depends on (pixel, frame) only (scene.slang:91-101)
};
// camera.slang:58-70
MPT_DEV f2 project_pinhole(const CameraParams& c, f3 world_pos, uint32_t w, uint32_t h) {
    const f3 d = world_pos - c.position;
    const float t = dot(c.W, c.W) / dot(d, c.W);
    const f3 plane_pos = t * d - c.W;
    const f2 ndc = mk2(dot(plane_pos, c.U) / dot(c.U, c.U), dot(plane_pos, c.V) / dot(c.V, c.V));
    return mk2((ndc.x + 1.f) / 2.f * float(w) - 0.5f, (ndc.y - 1.f) / -2.f * float(h) - 0.5f);
}
// Every operation is individually rounded (no FMA contraction) so that the primary rays are bit-identical to a plain
// IEEE evaluation of camera.slang:29-56 on the host; primary-visibility parity is then exact up to true ties.
MPT_DEV f3 normalize_rn(f3 v) {
    const float l = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)), __fmul_rn(v.z, v.z)));
    return mk3(__fdiv_rn(v.x, l), __fdiv_rn(v.y, l), __fdiv_rn(v.z, l));
}
// xxhash32(uint2), hash.slang:68-77
MPT_DEV uint32_t xxhash32_2(uint32_t px, uint32_t py) {
    const uint32_t P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
    uint32_t h = py + P5 + px * P3;
    h = P4 * ((h << 17) | (h >> 15));
    h = P2 * (h ^ (h >> 15));
    h = P3 * (h ^ (h >> 13));
    return h ^ (h >> 16);
}
// Scene::generate_camera_ray (scene.slang:94-101): pinhole (camera.slang:45-56) or, with aperture_radius > 0, thin lens
// (camera.slang:73-92) whose lens sample is the R2 sequence over frames (pseudorandom.slang:35-38) rotated per pixel by PCG(xxhash32(pixel)).
// The pinhole path is evaluated with individually rounded operations (see normalize_rn); the lens sample goes through sin / cos, whose
// last bits differ between libraries, so thin-lens rays agree with a host evaluation to a few ulp only.
MPT_DEV void camera_ray(const CameraParams& c, uint32_t px, uint32_t py, uint32_t w, uint32_t h, f3& o, f3& d, float& tmin, float& tmax) {
    const float pxn = __fdiv_rn(__fadd_rn(__fadd_rn(float(px), 0.5f), c.jitter_x), float(w));
    const float pyn = __fdiv_rn(__fadd_rn(__fadd_rn(float(py), 0.5f), c.jitter_y), float(h));
    const float nx = __fadd_rn(__fmul_rn(2.f, pxn), -1.f), ny = __fadd_rn(__fmul_rn(-2.f, pyn), 1.f);
    o = c.position;
    f3 dv = mk3(__fadd_rn(__fadd_rn(__fmul_rn(nx, c.U.x), __fmul_rn(ny, c.V.x)), c.W.x),
                __fadd_rn(__fadd_rn(__fmul_rn(nx, c.U.y), __fmul_rn(ny, c.V.y)), c.W.y),
                __fadd_rn(__fadd_rn(__fmul_rn(nx, c.U.z), __fmul_rn(ny, c.V.z)), c.W.z));
    if (c.aperture_radius > 0.f) {
        Rng rng = make_pcg(xxhash32_2(px, py));
        const float fn = float(c.frame);
        const float r2x = __fadd_rn(0.5f, __fmul_rn(0.7548776662466927f, fn)), r2y = __fadd_rn(0.5f, __fmul_rn(0.5698402909980532f, fn));
        const float sx = __fadd_rn(r2x - floorf(r2x), rng.next_float()), sy = __fadd_rn(r2y - floorf(r2y), rng.next_float());
        const float rx = sx - floorf(sx), ry = sy - floorf(sy);
        const float angle = __fmul_rn(kTwoPi, rx), rad = __fsqrt_rn(ry);
        const float dx = __fmul_rn(cosf(angle), rad), dy = __fmul_rn(sinf(angle), rad); // sample_disk, sampling.slang:39-42
        const f3 un = normalize_rn(c.U), vn = normalize_rn(c.V);
        const f3 off = mk3(__fmul_rn(c.aperture_radius, __fadd_rn(__fmul_rn(dx, un.x), __fmul_rn(dy, vn.x))),
                           __fmul_rn(c.aperture_radius, __fadd_rn(__fmul_rn(dx, un.y), __fmul_rn(dy, vn.y))),
                           __fmul_rn(c.aperture_radius, __fadd_rn(__fmul_rn(dx, un.z), __fmul_rn(dy, vn.z))));
        o = mk3(__fadd_rn(c.position.x, off.x), __fadd_rn(c.position.y, off.y), __fadd_rn(c.position.z, off.z));
        dv = mk3(__fsub_rn(dv.x, off.x), __fsub_rn(dv.y, off.y), __fsub_rn(dv.z, off.z));
    }
    d = normalize_rn(dv);
    const f3 wn = normalize_rn(c.W);
    const float inv_cos = __fdiv_rn(1.f, __fadd_rn(__fadd_rn(__fmul_rn(wn.x, d.x), __fmul_rn(wn.y, d.y)), __fmul_rn(wn.z, d.z)));
    tmin = __fmul_rn(c.near_plane, inv_cos);
    tmax = __fmul_rn(c.far_plane, inv_cos);
}

} // namespace mpt
