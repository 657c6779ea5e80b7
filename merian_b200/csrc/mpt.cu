// mpt.cu — libmpt.so: wavefront path tracer kernels for sm_100a and the C-ABI of include/mpt.h.
//
// Pipeline per pass (K samples per pixel in flight, slots = K * pixels):
//   k_shade_primary   one thread per pixel: primary shading point from the GBuffer hit (render_pt.slang:99-103),
//                     then for each of the K samples: seed RNG, BSDF sample, continuation ray -> queue
//   for bounce = 1 .. L-1:
//     k_extend        persistent warps, per-lane dynamic fetch from the ray queue (atomic counter), BVH8 closest hit;
//                     misses are shaded in place (environment lookup, path ends), hits are compacted into the hit
//                     queue with warp ballot + prefix popcount + one atomicAdd per warp
//     k_shade_bounce  hit queue only: shading point, Beer-Lambert, emission accumulate, next BSDF sample; survivors
//                     are compacted into the next ray queue the same way
//   k_resolve         per pixel: NaN/Inf rejection, sum of the K sample radiances and luminance^2 (render_pt.slang:116-129)
// Queues are double-buffered SoA (origin|slot, direction|tmax); per-path state lives in fixed slots.
#include "../../include/mpt.h"
#include "bvh8.h"
#include "float_image_files.h"
#include "scene_build.h"
#include "texture_mips.h"
#include "mpt_internal.h"
#include "shading.cuh"
#include "bvh8_traverse_pooled.cuh"

#include <cctype>
#include <chrono>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace mpt;

// Ray / hit / shadow queue records are written once and read once per bounce (gigabytes per frame), while the BVH nodes and triangle
// records (tens of MB) are re-read by every ray. Queue READS are marked streaming (ld.global.cs: the line is dead after the read).
// Measured on C3 (profiles/r02a_ab.jsonl, r02b_ab.jsonl) with streaming marks on reads AND writes of queues AND path state: extend
// -1.5 % but shade +10 % (5.10 -> 5.65 ms: what a kernel writes evict-first is gone from L2 when the next kernel wants it), so writes and
// the read-modify-write path state stay plain. The persisting-L2 window (MPT_L2_PERSIST=1) is opt-in: it left extend unchanged and
// doubled the shade time (the 64 MB carve-out is taken from the L2 that also holds attribute records, environment texels, path state).
#ifndef MPT_STREAM_HINTS
#define MPT_STREAM_HINTS 1
#endif
template <class T> __device__ __forceinline__ T ldq(const T* p) {
#if MPT_STREAM_HINTS
    return __ldcs(p);
#else
    return *p;
#endif
}

// ======================================================================================= device side
namespace {

constexpr int kMaxBounces = 64;
constexpr int kThreads = 128;

struct Counters {
    uint32_t qcount[kMaxBounces + 2]; // qcount[b] = rays queued for the extend of bounce b
    uint32_t hcount[kMaxBounces + 2]; // hcount[b] = surface hits found by the extend of bounce b
    uint32_t fetch[kMaxBounces + 2];  // persistent-thread fetch cursors
    uint32_t scount[kMaxBounces + 2]; // scount[b] = shadow rays queued by the shade of bounce b - 1 (MPT_RENDER_NEE)
    uint32_t sfetch[kMaxBounces + 2]; // fetch cursors of the persistent shadow kernel
    unsigned long long rays_continuation, rays_primary, shaded_hits, nodes_visited, tris_tested, rays_shadow;
};

struct PackedHitInfo { // gbuffer.slang:40-61
    uint32_t instance_id, primitive_id;
    uint16_t geometry_index, instance_index;
    __half2 bary;
};
static_assert(sizeof(PackedHitInfo) == 16, "PackedHitInfo is one uint4 texel");

struct PathBuffers {
    float4* q_o[2];   // origin.xyz | slot (bits)
    float4* q_d[2];   // direction.xyz | tmax
    uint4* hq_hit;    // hit queue: tri | instance | u | v
    float4* hq_dir;   // hit queue: ray direction.xyz | slot (bits)
    float4* hq_org;   // hit queue: ray origin.xyz | -   (only written when the scene has transmissive materials)
    float4* thr;      // throughput.rgb | current_ior
    float4* rad;      // radiance.rgb | rng state (bits)
    float4* med;      // medium mu_t.rgb | -   (only touched when the scene has transmissive materials)
    float4* sq_o;     // shadow queue (MPT_RENDER_NEE): origin.xyz | -
    float4* sq_d;     //                                direction.xyz | triangle record the ray leaves (bits)
    float4* sq_c;     //                                MIS-weighted contribution.rgb if the ray reaches the environment | slot (bits)
};

struct RenderConst {
    uint32_t width, height, npix;
    uint32_t frame;
    int32_t max_path_length;
    uint32_t instance_mask;
    uint32_t flags;
    uint32_t sample_offset, sample_stride;
    uint32_t pass_first; // index (within this call) of the first sample of the pass
    uint32_t pass_k;     // samples per pixel in this pass
    int32_t spp;
    uint32_t first_pass, last_pass;
};

MPT_DEV uint32_t lane_id() { return threadIdx.x & 31u; }

// Warp-aggregated queue append: one atomicAdd per warp. All 32 lanes must call.
MPT_DEV uint32_t queue_push_index(bool alive, uint32_t* counter) {
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, alive);
    uint32_t base = 0;
    if (lane_id() == 0 && m) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    return base + __popc(m & ((1u << lane_id()) - 1u));
}

// One scattering step of trace_path (render_pt.slang:35-76): BSDF sample, throughput, medium bookkeeping, make_ray.
template <uint32_t FEAT>
MPT_DEV bool scatter(const ShadingPoint& sp, f3 ray_d, Rng& rng, f3& throughput, float& current_ior, f3& medium, f3& out_o, f3& out_d, float& out_pdf) {
    if (!sp.has_material || all_below(throughput, kThroughputCut)) return false;
    const f3 wi_world = -ray_d;
    const f3 wi = sp.frame.to_local(wi_world);
    f3 wo_l, weight;
    float pdf;
    if (!bsdf_sample_eval<FEAT>(sp.bsdf, wi, rng, wo_l, pdf, weight)) return false;
    out_pdf = pdf;
    const f3 wo = sp.frame.to_world(wo_l);
    throughput *= weight;
    const bool transmission = wi.z * wo_l.z < 0.f;
    const f3 ng = sp.face_normal_dec;
    if (transmission) {
        const float eta = sp.bsdf.gltf ? sp.bsdf.eta : 1.0f; // MaterialSample::get_interior_volume
        current_ior /= eta;
        medium = (dot(wi_world, ng) >= 0.f) ? sp.absorption : mk3(0.f);
    }
    // ShadingPoint::make_ray, hit.slang:134-147
    const float wo_ng = dot(wo, ng);
    if (((dot(wi_world, ng) < 0.f) != (wo_ng < 0.f)) != transmission) return false;
    out_o = offset_ray_origin(sp.pos, wo_ng < 0.f ? -ng : ng);
    out_d = wo;
    return true;
}

// Next event estimation towards the lat-long environment (MPT_RENDER_NEE; an addition, the reference's pt has none: SURVEY §0.1).
// One light sample per path vertex from the alias table, combined with the BSDF-sampled continuation by the balance heuristic:
//   light sample:  T * f*cos * L / (p_light + p_bsdf)   if the shadow ray reaches the environment
//   BSDF sample that misses all geometry (k_extend):  weight p_bsdf / (p_bsdf + p_light)
// so the expectation equals the reference estimator's. Emissive surfaces stay BSDF-sampled only. The light sample draws from
// its own RNG stream (keyed on the path's RNG state), so the BSDF-sampled paths are the same with and without NEE.
template <uint32_t FEAT>
MPT_DEV bool nee_environment(const SceneView& S, const ShadingPoint& sp, f3 ray_d, uint32_t rng_state, f3 throughput, f3 medium, f3& so, f3& sd, f3& contrib) {
    if (S.env_type != 1u || !sp.has_material || all_below(throughput, kThroughputCut)) return false;
    Rng r = make_pcg(pcg4d16(rng_state, 0x9E3779B9u, 0x7F4A7C15u, 1u));
    const float u0 = r.next_float(), u1 = r.next_float(), u2 = r.next_float();
    f3 wo;
    float p_light;
    if (!env_sample(S, u0, u1, u2, wo, p_light)) return false;
    const f3 wi_world = -ray_d;
    const f3 wi = sp.frame.to_local(wi_world), wo_l = sp.frame.to_local(wo);
    f3 e;
    float p_bsdf;
    bsdf_eval_pdf<FEAT>(sp.bsdf, wi, wo_l, e, p_bsdf);
    if (!(e.x > 0.f || e.y > 0.f || e.z > 0.f)) return false;
    const bool transmission = wi.z * wo_l.z < 0.f;
    const f3 ng = sp.face_normal_dec;
    const float wo_ng = dot(wo, ng);
    if (((dot(wi_world, ng) < 0.f) != (wo_ng < 0.f)) != transmission) return false; // ShadingPoint::make_ray, hit.slang:134-147
    // the medium the shadow segment would cross (same bookkeeping as scatter): an absorbing interior over ENV_MAP_DISTANCE
    // transmits nothing, exactly like the BSDF-sampled ray that escapes from inside it
    f3 med = medium;
    if (transmission) med = (dot(wi_world, ng) >= 0.f) ? sp.absorption : mk3(0.f);
    if (med.x != 0.f || med.y != 0.f || med.z != 0.f) return false;
    contrib = throughput * e * env_eval(S, wo) / (p_light + p_bsdf);
    if (!finite3(contrib) || !(contrib.x > 0.f || contrib.y > 0.f || contrib.z > 0.f)) return false;
    so = offset_ray_origin(sp.pos, wo_ng < 0.f ? -ng : ng);
    sd = wo;
    return true;
}

// --------------------------------------------------------------------------------------- primary visibility
template <bool COUNT, bool TWO_LEVEL>
__global__ void __launch_bounds__(kThreads) k_primary(const __grid_constant__ SceneView S, CameraParams cam, uint32_t w, uint32_t h, uint32_t mask, PackedHitInfo* out,
                                                      Counters* ctr) {
    // 8x4 pixel tiles per warp keep primary rays of a warp spatially coherent
    const uint32_t tiles_x = (w + 7) / 8;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t tx = warp % tiles_x, ty = warp / tiles_x;
    const uint32_t x = tx * 8 + (lane_id() & 7u), y = ty * 4 + (lane_id() >> 3);
    if (x >= w || y >= h) return;
    f3 o, d;
    float tmin, tmax;
    camera_ray(cam, x, y, w, h, o, d, tmin, tmax);
    uint32_t nn = 0, nt = 0;
    MaterialAlphaTest at{&S};
    const HitRec hr = trace_closest<COUNT, TWO_LEVEL>(S.accel, o, d, tmin, tmax, mask, at, nn, nt);
    PackedHitInfo p;
    if (hr.tri == kMiss) {
        p.instance_id = kMiss;
        p.primitive_id = kMiss;
        p.geometry_index = 0xFFFF;
        p.instance_index = 0xFFFF;
        p.bary = __floats2half2_rn(0.f, 0.f);
    } else {
        const float4* tp = S.accel.tris + size_t(hr.tri) * 3;
        p.instance_id = S.accel.instances[hr.instance].custom_index;
        p.primitive_id = __float_as_uint(__ldg(tp).w);
        p.geometry_index = uint16_t(__float_as_uint(__ldg(tp + 1).w));
        p.instance_index = uint16_t(hr.instance);
        p.bary = __floats2half2_rn(hr.u, hr.v);
    }
    out[size_t(y) * w + x] = p;
    if (COUNT) {
        atomicAdd(&ctr->nodes_visited, (unsigned long long)nn);
        atomicAdd(&ctr->tris_tested, (unsigned long long)nt);
    }
}

// --------------------------------------------------------------------------------------- GBuffer (Raytraced) node
// gbuffer_rt.slang:18-94 with the stores of gbuffer.slang:152-172: primary hit, material shading normal, linear depth and its
// screen-space gradient, depth change and motion vector against the previous camera, directional albedo, emission.
// prev_pos: the hit point carried by the previous frame's instance transform for animated geometry (Scene::get_prev_world_position,
// scene.slang:392-401, to_prev_world_position :261-270), else the hit point itself. Material textures are sampled with the
// ray-differential UV gradients of gbuffer_rt.slang:41-50.
struct GBufferOut {
    uint2* tex0;           // enc_normal | half2 grad_z
    PackedHitInfo* tex1;
    float4* tex2;          // motion vector.xy | linear_z | delta_z
    float4* tex3;          // albedo.rgb | 1
    float4* emission;
};
template <uint32_t FEAT, bool TWO_LEVEL>
__global__ void __launch_bounds__(kThreads) k_gbuffer(const __grid_constant__ SceneView S, CameraParams cam, CameraParams prev_cam, uint32_t w, uint32_t h, uint32_t mask,
                                                      GBufferOut out) {
    const uint32_t tiles_x = (w + 7) / 8;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t tx = warp % tiles_x, ty = warp / tiles_x;
    const uint32_t x = tx * 8 + (lane_id() & 7u), y = ty * 4 + (lane_id() >> 3);
    if (x >= w || y >= h) return;
    f3 o, d;
    float tmin, tmax;
    camera_ray(cam, x, y, w, h, o, d, tmin, tmax);
    uint32_t nn = 0, nt = 0;
    MaterialAlphaTest at{&S};
    const HitRec hr = trace_closest<false, TWO_LEVEL>(S.accel, o, d, tmin, tmax, mask, at, nn, nt);
    const size_t pix = size_t(y) * w + x;
    PackedHitInfo p;
    ShadingPoint sp;
    bool surface = false;
    if (hr.tri == kMiss) {
        p.instance_id = kMiss; p.primitive_id = kMiss; p.geometry_index = 0xFFFF; p.instance_index = 0xFFFF;
        p.bary = __floats2half2_rn(0.f, 0.f);
    } else {
        const float4* tp = S.accel.tris + size_t(hr.tri) * 3;
        p.instance_id = S.accel.instances[hr.instance].custom_index;
        p.primitive_id = __float_as_uint(__ldg(tp).w);
        p.geometry_index = uint16_t(__float_as_uint(__ldg(tp + 1).w));
        p.instance_index = uint16_t(hr.instance);
        p.bary = __floats2half2_rn(hr.u, hr.v);
        f3 ox, oy, diff[2]; // r_x, r_y of gbuffer_rt.slang:36-37: the camera rays of the neighbouring pixels
        float t0, t1;
        CameraParams pin = cam; // get_pinhole_direction (camera.slang:34-37), whatever the aperture
        pin.aperture_radius = 0.f;
        camera_ray(pin, x + 1, y, w, h, ox, diff[0], t0, t1);
        camera_ray(pin, x, y + 1, w, h, oy, diff[1], t0, t1);
        surface_shading_point<FEAT>(S, o, d, hr, 1.0f, true, sp, diff); // falls back to the environment for UseEnvMap geometry
        surface = sp.has_material;
    }
    if (!surface) env_shading_point(S, o, d, sp);
    if (out.tex1) out.tex1[pix] = p;
    const float linear_z = length(sp.pos - o);
    f2 grad_z = mk2(0.f, 0.f);
    float delta_z = 0.f;
    f3 albedo = sp.emission; // miss: store_albedo(Albedo(sp.emission)), gbuffer_rt.slang:90
    f3 prev_pos = sp.pos;
    if (surface) {
        f3 ox, rx, oy, ry;
        float t0, t1;
        CameraParams pin = cam;
        pin.aperture_radius = 0.f;
        camera_ray(pin, x + 1, y, w, h, ox, rx, t0, t1);
        camera_ray(pin, x, y + 1, w, h, oy, ry, t0, t1);
        const float n_dot_oo = dot(sp.geo_normal, sp.pos - o); // Surface::compute_grad_z, hit.slang:99-103
        grad_z = mk2(n_dot_oo / dot(sp.geo_normal, rx) - linear_z, n_dot_oo / dot(sp.geo_normal, ry) - linear_z);
        if (S.inst_xform_prev) {
            const float4* tp = S.accel.tris + size_t(hr.tri) * 3;
            const float4 ta = __ldg(tp), tb = __ldg(tp + 1), tc = __ldg(tp + 2);
            const uint32_t gflags = __ldg(S.geometries + S.accel.instances[hr.instance].custom_index + __float_as_uint(tb.w)).x >> 16;
            if ((gflags & G_IS_ANIMATED) && !(gflags & G_PRETRANSFORMED)) {
                const f3 p_obj = (1.0f - hr.u - hr.v) * xyz(ta) + hr.u * xyz(tb) + hr.v * xyz(tc); // TriangleHit::get_position, as surface_shading_point
                prev_pos = xf_point(S.inst_xform_prev + size_t(hr.instance) * 4, p_obj);
            }
        }
        delta_z = length(prev_pos - prev_cam.position) - linear_z;
        albedo = bsdf_albedo<FEAT>(sp.bsdf, sp.frame.to_local(-d));
    }
    const f2 pp = project_pinhole(prev_cam, prev_pos, w, h); // Scene::compute_motion_vector, scene.slang:404-414
    const f2 mv = mk2(pp.x - float(x) - cam.jitter_x, pp.y - float(y) - cam.jitter_y);
    if (out.tex0) {
        const __half2 g = __floats2half2_rn(grad_z.x, grad_z.y);
        out.tex0[pix] = make_uint2(encode_normal(sp.frame.n), *reinterpret_cast<const uint32_t*>(&g));
    }
    if (out.tex2) out.tex2[pix] = make_float4(mv.x, mv.y, fminf(linear_z, 65504.0f), delta_z);
    if (out.tex3) out.tex3[pix] = make_float4(albedo.x, albedo.y, albedo.z, 1.f);
    if (out.emission) out.emission[pix] = make_float4(sp.emission.x, sp.emission.y, sp.emission.z, 0.f); // x transmittance of Vacuum = 1
}

// --------------------------------------------------------------------------------------- arbitrary rays
template <bool COUNT, bool TWO_LEVEL>
__global__ void __launch_bounds__(kThreads) k_trace(const __grid_constant__ SceneView S, const float4* rays, uint64_t n, uint32_t mask, uint32_t* hits6, Counters* ctr) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 ro = __ldg(rays + 2 * i), rd = __ldg(rays + 2 * i + 1);
    uint32_t nn = 0, nt = 0;
    MaterialAlphaTest at{&S};
    const HitRec hr = trace_closest<COUNT, TWO_LEVEL>(S.accel, xyz(ro), xyz(rd), ro.w, rd.w, mask, at, nn, nt);
    uint32_t* o = hits6 + 6 * i;
    if (hr.tri == kMiss) {
        o[0] = o[1] = o[2] = o[3] = kMiss;
        o[4] = o[5] = 0u;
    } else {
        const float4* tp = S.accel.tris + size_t(hr.tri) * 3;
        o[0] = S.accel.instances[hr.instance].custom_index;
        o[1] = hr.instance;
        o[2] = __float_as_uint(__ldg(tp).w);
        o[3] = __float_as_uint(__ldg(tp + 1).w);
        o[4] = __float_as_uint(hr.u);
        o[5] = __float_as_uint(hr.v);
    }
    if (COUNT) {
        atomicAdd(&ctr->nodes_visited, (unsigned long long)nn);
        atomicAdd(&ctr->tris_tested, (unsigned long long)nt);
    }
}

// --------------------------------------------------------------------------------------- wavefront: primary shade
// Core-glTF shading (FEAT = 0) is latency-bound, not register-bound: 8 resident blocks/SM at 64 registers (a few spilled words) beat
// 5 blocks at 86 registers by 12 % of the shade time (5.84 -> 5.13 ms/frame on C3; 6 blocks: +-0, 10 blocks/48 regs: -3 %, 12 blocks: +3 %
// slower; profiles/r01f_extend_experiments.md). The extension variants keep the compiler's choice.
#ifndef MPT_SHADE_MIN_BLOCKS
#define MPT_SHADE_MIN_BLOCKS 8
#endif
template <uint32_t FEAT, bool NEE>
__global__ void __launch_bounds__(kThreads, FEAT == 0u ? MPT_SHADE_MIN_BLOCKS : 1) k_shade_primary(const __grid_constant__ SceneView S, CameraParams cam, RenderConst rc, const PackedHitInfo* tex1, PathBuffers pb,
                                                            Counters* ctr) {
    const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = pix < rc.npix;
    const uint32_t x = valid ? pix % rc.width : 0, y = valid ? pix / rc.width : 0;
    f3 o, d;
    float tmin, tmax;
    camera_ray(cam, x, y, rc.width, rc.height, o, d, tmin, tmax);
    ShadingPoint sp;
    sp.has_material = false;
    sp.emission = mk3(0.f);
    sp.pos = o;
    bool shaded = false;
    uint32_t origin_tri = kMiss; // the triangle record the continuation rays leave (k_extend need not test it)
    if (valid) {
        const PackedHitInfo ph = tex1[pix];
        if (ph.instance_id != kMiss) { // Hit::is_valid, hit.slang:53-55
            const uint4 gd = __ldg(S.geometries + ph.instance_id + uint32_t(ph.geometry_index));
            HitRec hr;
            hr.tri = __ldg(S.prim_to_tri + gd.z + ph.primitive_id);
            origin_tri = hr.tri;
            hr.instance = ph.instance_index;
            const float2 b = __half22float2(ph.bary);
            hr.u = b.x;
            hr.v = b.y;
            surface_shading_point<FEAT>(S, o, d, hr, 1.0f, true, sp);
            shaded = sp.has_material;
        } else {
            env_shading_point(S, o, d, sp);
        }
    }
    // trace_path prologue, render_pt.slang:29-33 (exterior volume = Vacuum: transmittance 1)
    const f3 radiance0 = (rc.flags & MPT_RENDER_EMISSION_ON_PRIMARY) ? sp.emission : mk3(0.f);
    const uint32_t ms = __ballot_sync(0xFFFFFFFFu, shaded);
    if (lane_id() == 0 && ms) atomicAdd(&ctr->shaded_hits, (unsigned long long)__popc(ms));

    for (uint32_t k = 0; k < rc.pass_k; k++) {
        const uint32_t slot = k * rc.npix + pix;
        const uint32_t sidx = rc.sample_offset + (rc.pass_first + k) * rc.sample_stride;
        Rng rng = make_pcg(pcg4d16(x, y, rc.frame, sidx)); // render_pt.slang:114
        f3 throughput = mk3(1.f), medium = mk3(0.f), no, nd;
        float ior = 1.0f, pdf = 0.f;
        bool alive = false;
        if (NEE) {
            f3 so, sd, sc;
            const bool lit = valid && rc.max_path_length > 1 && nee_environment<FEAT>(S, sp, d, rng.state, throughput, medium, so, sd, sc);
            const uint32_t si = queue_push_index(lit, &ctr->scount[1]);
            if (lit) {
                pb.sq_o[si] = make_float4(so.x, so.y, so.z, 0.f);
                pb.sq_d[si] = make_float4(sd.x, sd.y, sd.z, __uint_as_float(origin_tri)); // the triangle the shadow ray leaves
                pb.sq_c[si] = make_float4(sc.x, sc.y, sc.z, __uint_as_float(slot));
            }
        }
        if (valid && rc.max_path_length > 1) alive = scatter<FEAT>(sp, d, rng, throughput, ior, medium, no, nd, pdf);
        const uint32_t qi = queue_push_index(alive, &ctr->qcount[1]);
        if (alive) {
            pb.q_o[0][qi] = make_float4(no.x, no.y, no.z, __uint_as_float(slot));
            pb.q_d[0][qi] = make_float4(nd.x, nd.y, nd.z, NEE ? pdf : __uint_as_float(origin_tri));
        }
        if (valid) {
            pb.rad[slot] = make_float4(radiance0.x, radiance0.y, radiance0.z, __uint_as_float(rng.state));
            if (alive) {
                pb.thr[slot] = make_float4(throughput.x, throughput.y, throughput.z, ior);
                if (FEAT & F_TRANSMISSION) pb.med[slot] = make_float4(medium.x, medium.y, medium.z, 0.f);
            }
        }
    }
}

// --------------------------------------------------------------------------------------- wavefront: extend
// Persistent warps with per-lane dynamic fetch: a lane whose ray has finished does not wait for the slowest ray of
// its batch; when fewer than tune.refill (14) lanes of the warp are still traversing, the warp re-converges, all idle
// lanes claim new rays with ONE atomicAdd (ballot + prefix popcount), and traversal resumes.
struct ExtendTune {
    uint32_t refill;   // re-converge and refill when fewer than this many lanes are still traversing
};
// 7 resident blocks of 128 threads per SM (<= 72 registers, no spills): measured optimum on B200 (6 blocks/80 regs: -5 %,
// 8 blocks/64 regs with spills: -2 %, DESIGN.md "extend: occupancy").
#ifndef MPT_EXTEND_MIN_BLOCKS
#define MPT_EXTEND_MIN_BLOCKS 7
#endif
// The plain variant (single-level scene, no medium, no NEE, no counters) fits 64 registers without spills since the traversal lost a
// register to the origin-triangle trick: 8 blocks/SM, -0.9 % extend time on C3. The other variants keep 7 blocks / 72 registers.
#ifndef MPT_EXTEND_PLAIN_MIN_BLOCKS
#define MPT_EXTEND_PLAIN_MIN_BLOCKS 8
#endif
template <bool COUNT, bool MEDIUM, bool TWO_LEVEL, bool NEE>
__global__ void __launch_bounds__(kThreads, (!COUNT && !MEDIUM && !TWO_LEVEL && !NEE) ? MPT_EXTEND_PLAIN_MIN_BLOCKS : MPT_EXTEND_MIN_BLOCKS) k_extend(const __grid_constant__ SceneView S, const float4* __restrict__ q_o, const float4* __restrict__ q_d,
                                                     PathBuffers pb, uint32_t bounce, uint32_t mask, ExtendTune tune, Counters* ctr) {
    const uint32_t n = ctr->qcount[bounce];
    MaterialAlphaTest at{&S};
    uint2 stack_mem[kStackSize];
    Traversal<COUNT, TWO_LEVEL, MaterialAlphaTest> T(S.accel, at, stack_mem);
    if (TWO_LEVEL) { // world-space ray setup of every thread, word-major: leaving an instance reloads it (8 LDS) instead of redoing 6 divisions
        __shared__ float s_world[8 * kThreads];
        T.wcache = s_world + threadIdx.x;
        T.wcache_stride = kThreads;
    }

    bool active = false, exhausted = false, finished = false;
    uint32_t slot = 0;
    float pdf_bsdf = 0.f; // NEE only: the density the ray's direction was sampled with (rides in q_d.w)
    while (true) {
        // ---- retire finished rays (the warp is converged here)
        const bool is_hit = finished && T.found();
        const uint32_t hm = __ballot_sync(0xFFFFFFFFu, is_hit);
        if (hm) {
            const int leader = __ffs(hm) - 1;
            uint32_t base = 0;
            if (int(lane_id()) == leader) base = atomicAdd(&ctr->hcount[bounce], uint32_t(__popc(hm)));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (is_hit) {
                const uint32_t hi = base + __popc(hm & ((1u << lane_id()) - 1u));
                pb.hq_hit[hi] = make_uint4(T.hit.tri, T.hit.instance, __float_as_uint(T.hit.u), __float_as_uint(T.hit.v));
                const f3 rd = T.dir();
                pb.hq_dir[hi] = make_float4(rd.x, rd.y, rd.z, __uint_as_float(slot));
                if (MEDIUM) { const f3 ro = T.org(); pb.hq_org[hi] = make_float4(ro.x, ro.y, ro.z, 0.f); }
            }
        }
        if (finished && !is_hit && S.env_type != 0u) {
            // miss: _make_environment_map_shading_point (scene.slang:482-498) + render_pt.slang:84-86; the path ends
            const float4 th = pb.thr[slot];
            f3 throughput = mk3(th.x, th.y, th.z);
            if (MEDIUM) {
                const float4 m = pb.med[slot];
                if (m.x != 0.f || m.y != 0.f || m.z != 0.f) throughput *= vexp(-(kEnvDistance * length(T.dir())) * mk3(m.x, m.y, m.z));
            }
            f3 e = throughput * env_eval(S, T.dir());
            if (NEE) e = e * (pdf_bsdf / (pdf_bsdf + env_pdf(S, T.dir()))); // balance heuristic against nee_environment
            float4 ra = pb.rad[slot];
            ra.x += e.x;
            ra.y += e.y;
            ra.z += e.z;
            pb.rad[slot] = ra;
        }
        finished = false;
        // ---- refill idle lanes
        const uint32_t need = __ballot_sync(0xFFFFFFFFu, !active && !exhausted);
        if (need) {
            const int leader = __ffs(need) - 1;
            uint32_t base = 0;
            if (int(lane_id()) == leader) base = atomicAdd(&ctr->fetch[bounce], uint32_t(__popc(need)));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (!active && !exhausted) {
                const uint32_t j = base + __popc(need & ((1u << lane_id()) - 1u));
                if (j < n) {
                    const float4 ro = ldq(q_o + j), rd = ldq(q_d + j);
                    slot = __float_as_uint(ro.w);
                    // continuation rays all have tmax = SCENE_RAY_TMAX; the lane carries the BSDF pdf (NEE) or the triangle the ray leaves
                    if (NEE) pdf_bsdf = rd.w;
                    active = T.init(xyz(ro), xyz(rd), 0.f /* SCENE_RAY_TMIN */, kRayTMax, mask, NEE ? kMiss : __float_as_uint(rd.w));
                    finished = !active; // masked out: a miss
                } else {
                    exhausted = true;
                }
            }
        }
        const bool queue_empty = __any_sync(0xFFFFFFFFu, exhausted);
        if (!__any_sync(0xFFFFFFFFu, active || finished)) break;
        while (active) {
            if (T.step()) {
                active = false;
                finished = true;
                break;
            }
            if (!queue_empty && uint32_t(__popc(__activemask())) < tune.refill) break;
        }
        __syncwarp();
    }
    if (COUNT) {
        atomicAdd(&ctr->nodes_visited, (unsigned long long)T.n_nodes);
        atomicAdd(&ctr->tris_tested, (unsigned long long)T.n_tris);
    }
}

// --------------------------------------------------------------------------------------- wavefront: extend, pooled form
// The organisation described in bvh8_traverse_pooled.cuh: per-warp ray pool in shared memory, converged retire + refill + ray setup,
// divergent traversal in which a finished lane swaps its record for a READY one with two shared-memory atomics.
// Measured on C3 (profiles/r02a_*): node test 27.0 of 32 lanes (round 1: 22.8), extend 36.6 -> 32.9 ms per frame with the stack in
// local memory; shared memory is taken from L1, and every KB costs: READY 32 instead of 16: +1.7 %, 4 / 8 stack entries per lane in
// shared memory: +4.5 % / +4.1 %; READY 8: +6 % (the pool runs dry), READY 24: +3 %. 48-byte records (direction and path slot re-read
// from the queue when the ray retires) were slower than 64-byte ones: the converged phase then chains two dependent DRAM latencies.
// Postponing the primitive section to every 2nd / 4th loop iteration: +7 % / +34 % (the hit distance lags, rays take more iterations).
#ifndef MPT_POOL_READY
#define MPT_POOL_READY 32 // READY records per warp on top of the 32 a warp can be traversing
#endif
#ifndef MPT_POOL_SMEM_STACK
#define MPT_POOL_SMEM_STACK 0 // traversal-stack entries per lane kept in shared memory (the rest: local memory, i.e. L1)
#endif
// MPT_POOL_TRIGGER = T > 0: the warp re-converges as soon as T records are DONE (polled once per step) instead of waiting until the READY
// list is empty and the busy lanes have dwindled below tune.refill: the converged phase then always finds (a multiple of) 32 records
// to retire and refill — it runs at full width — and a finishing lane (almost) always finds a READY record, so the traversal loop
// does not ramp down between refills. Needs T <= MPT_POOL_READY (otherwise lanes could starve below the trigger). 0 = the ramp-down form.
// Measured (profiles/r02v_ab.md): READY 32 / trigger 32 against the ramp-down form with READY 16: C3 +0.7 %, C5 +3.0 %; trigger 16 (half-empty
// converged rounds) -4 %, READY 24 / trigger 24 +-0.
#ifndef MPT_POOL_TRIGGER
#define MPT_POOL_TRIGGER 32
#endif
template <bool COUNT, bool MEDIUM, bool TWO_LEVEL, bool NEE>
__global__ void __launch_bounds__(kThreads, (!COUNT && !MEDIUM && !TWO_LEVEL && !NEE) ? MPT_EXTEND_PLAIN_MIN_BLOCKS : MPT_EXTEND_MIN_BLOCKS)
k_extend_pooled(const __grid_constant__ SceneView S, const float4* __restrict__ q_o, const float4* __restrict__ q_d, PathBuffers pb, uint32_t bounce, uint32_t mask,
                ExtendTune tune, Counters* ctr) {
    constexpr int NE = 32 + MPT_POOL_READY, SS = MPT_POOL_SMEM_STACK, kWarps = kThreads / 32, TRIG = MPT_POOL_TRIGGER;
    static_assert(TRIG >= 0 && TRIG <= MPT_POOL_READY && MPT_POOL_READY <= 32, "MPT_POOL_TRIGGER <= MPT_POOL_READY <= 32");
    using Trav = TraversalP<COUNT, TWO_LEVEL, MaterialAlphaTest, SS, kThreads>;
    using Entry = PoolEntry;
    struct Pool {
        Entry e[NE];
        uint8_t ready[NE], done[NE];
        int ready_head, ready_n, done_n, pad;
    };
    __shared__ Pool s_pool[kWarps];
    __shared__ uint2 s_stack[(SS > 0 ? SS : 1) * (SS > 0 ? kThreads : 1)];
    const uint32_t n = ctr->qcount[bounce];
    const uint32_t lane = lane_id();
    const uint32_t lt = (1u << lane) - 1u;
    Pool& P = s_pool[threadIdx.x >> 5];
    uint2 lstack[kStackSize - SS];
    uint4 pstack[kPrimStackSize];
    MaterialAlphaTest at{&S};
    Trav T(S.accel, at, s_stack + (SS > 0 ? threadIdx.x : 0), lstack, pstack);
    const uint32_t scene_flags = TWO_LEVEL ? 0u : S.accel.instances[0].mask_flags;
    const bool scene_masked = !TWO_LEVEL && ((scene_flags & 0xFFu) & mask) == 0u; // the only instance is masked out: every ray misses
    constexpr uint32_t kNoRay = 0xFFFFFFFFu;

    // every record starts on the DONE list, marked empty (path slot kNoRay)
    for (int i = int(lane); i < NE; i += 32) {
        P.e[i].a.w = __uint_as_float(kNoRay);
        P.done[i] = uint8_t(i);
    }
    if (lane == 0) { P.done_n = NE; P.ready_head = 0; P.ready_n = 0; }
    __syncwarp();

    bool active = false, exhausted = false;
    uint32_t my = 0;
    int ready_n = 0;
    while (true) {
        // ======================================================= converged: retire DONE records, refill them, hand READY records to idle lanes
        const int nd_all = P.done_n;
        // trigger form: whole rounds of 32 only (what is left over stays DONE for the next phase), unless the queue is used up
        const int nd = (TRIG > 0 && !exhausted && nd_all > 32) ? (nd_all & ~31) : nd_all;
        uint32_t base = 0;
        if (!exhausted && nd > 0) {
            if (lane == 0) base = atomicAdd(&ctr->fetch[bounce], uint32_t(nd));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
        }
        int nready = 0;
        if (TRIG > 0) { // READY records nobody has taken yet move to the front of the list
            const int rh = min(P.ready_head, ready_n), left = ready_n - rh;
            const uint8_t v = int(lane) < left ? P.ready[rh + int(lane)] : uint8_t(0);
            __syncwarp();
            if (int(lane) < left) P.ready[lane] = v;
            nready = left;
        }
#pragma unroll 1
        for (int r = 0; r * 32 < nd; r++) {
            const int i = r * 32 + int(lane);
            const bool have = i < nd;
            const uint32_t e = have ? uint32_t(P.done[i]) : 0u;
            Entry& E = P.e[e];
            // ---- retire
            const uint32_t slot = have ? __float_as_uint(E.a.w) : kNoRay;
            const bool is_ray = slot != kNoRay;
            uint4 h = make_uint4(kMiss, kMiss, 0u, 0u);
            float4 dir = make_float4(0.f, 0.f, 0.f, 0.f);
            if (is_ray) {
                const float4 hb = E.b;
                h = make_uint4(__float_as_uint(hb.x), __float_as_uint(hb.y), __float_as_uint(hb.z), __float_as_uint(hb.w));
                dir = E.d; // direction | the BSDF pdf (NEE)
            }
            const bool is_hit = is_ray && h.y != kMiss;
            const uint32_t hm = __ballot_sync(0xFFFFFFFFu, is_hit);
            if (hm) {
                uint32_t hb = 0;
                if (lane == 0) hb = atomicAdd(&ctr->hcount[bounce], uint32_t(__popc(hm)));
                hb = __shfl_sync(0xFFFFFFFFu, hb, 0);
                if (is_hit) {
                    const uint32_t hi = hb + __popc(hm & lt);
                    pb.hq_hit[hi] = h;
                    pb.hq_dir[hi] = make_float4(dir.x, dir.y, dir.z, __uint_as_float(slot));
                    if (MEDIUM) { const float4 oa = E.a; pb.hq_org[hi] = make_float4(oa.x, oa.y, oa.z, 0.f); }
                }
            }
            if (is_ray && !is_hit && S.env_type != 0u) {
                // miss: _make_environment_map_shading_point (scene.slang:482-498) + render_pt.slang:84-86; the path ends
                const f3 rd = mk3(dir.x, dir.y, dir.z);
                const float4 th = pb.thr[slot];
                f3 throughput = mk3(th.x, th.y, th.z);
                if (MEDIUM) {
                    const float4 m = pb.med[slot];
                    if (m.x != 0.f || m.y != 0.f || m.z != 0.f) throughput *= vexp(-(kEnvDistance * length(rd)) * mk3(m.x, m.y, m.z));
                }
                f3 ev = throughput * env_eval(S, rd);
                if (NEE) ev = ev * (dir.w / (dir.w + env_pdf(S, rd))); // balance heuristic against nee_environment
                float4 ra = pb.rad[slot];
                ra.x += ev.x;
                ra.y += ev.y;
                ra.z += ev.z;
                pb.rad[slot] = ra;
            }
            // ---- refill: the i-th DONE record takes the i-th ray of this warp's claim
            bool filled = false;
            if (have) {
                const uint32_t jn = base + uint32_t(i);
                if (!exhausted && jn < n) {
                    const float4 ro = ldq(q_o + jn), rd = ldq(q_d + jn);
                    if (scene_masked) {
                        if (S.env_type != 0u) { // same miss shading as above (no medium can be active: nothing was ever hit)
                            const uint32_t sl = __float_as_uint(ro.w);
                            const float4 th = pb.thr[sl];
                            f3 ev = mk3(th.x, th.y, th.z) * env_eval(S, xyz(rd));
                            if (NEE) ev = ev * (rd.w / (rd.w + env_pdf(S, xyz(rd))));
                            float4 ra = pb.rad[sl];
                            ra.x += ev.x;
                            ra.y += ev.y;
                            ra.z += ev.z;
                            pb.rad[sl] = ra;
                        }
                    } else {
                        RaySetup R;
                        setup_ray(R, xyz(ro), xyz(rd));
                        const uint32_t oct = 7u - (R.perm >> 24);
                        E.a = ro; // origin | path slot
                        E.b = make_float4(R.idir.x, R.idir.y, R.idir.z, rd.w); // rd.w: the triangle the ray leaves (the BSDF pdf with NEE)
                        E.c = make_float4(R.shear.Sx, R.shear.Sy, R.shear.Sz,
                                          __uint_as_float(uint32_t(R.shear.kx) | (uint32_t(R.shear.ky) << 2) | (uint32_t(R.shear.kz) << 4) | (oct << 8)));
                        E.d = rd;
                        filled = true;
                    }
                }
                if (!filled) E.a.w = __uint_as_float(kNoRay);
            }
            const uint32_t fm = __ballot_sync(0xFFFFFFFFu, filled);
            if (filled) P.ready[nready + __popc(fm & lt)] = uint8_t(e);
            nready += __popc(fm);
        }
        if (!exhausted && nd > 0 && base + uint32_t(nd) >= n) exhausted = true;
        __syncwarp();
        // ---- idle lanes take READY records
        const uint32_t idle = __ballot_sync(0xFFFFFFFFu, !active);
        const int k = __popc(idle & lt);
        if (!active && k < nready) {
            my = P.ready[k];
            T.start(&P.e[my], mask, scene_flags, !NEE);
            active = true;
        }
        const int rem = nd_all - nd; // < 32
        if (TRIG > 0 && rem > 0) {
            const uint8_t v = int(lane) < rem ? P.done[nd + int(lane)] : uint8_t(0);
            __syncwarp();
            if (int(lane) < rem) P.done[lane] = v;
        }
        if (lane == 0) {
            P.done_n = rem;
            P.ready_n = nready;
            P.ready_head = min(__popc(idle), nready);
        }
        ready_n = nready;
        __syncwarp();
        if (!__any_sync(0xFFFFFFFFu, active)) {
            if (rem == 0) break;
            continue; // the queue ran out in this phase: the next one retires the rest
        }
        // ======================================================= divergent: traverse; a finished lane swaps its record for a READY one
        uint32_t tick = 0u; // the same in every lane of the warp (TraversalP::step, MPT_TLAS_GATE); unused and compiled out for single-level scenes
        while (active) {
            if (T.step(TWO_LEVEL ? tick++ : 0u)) {
                Entry& E = P.e[my];
                const bool f = T.found();
                E.b = make_float4(__uint_as_float(f ? T.hit.tri : kMiss), __uint_as_float(T.hit.instance), T.hit.u, T.hit.v);
                P.done[atomicAdd(&P.done_n, 1)] = uint8_t(my);
                const int kk = atomicAdd(&P.ready_head, 1);
                if (kk < ready_n) {
                    my = P.ready[kk];
                    T.start(&P.e[my], mask, scene_flags, !NEE);
                } else {
                    active = false;
                    break;
                }
            }
            // re-converge for a refill when the READY list is used up and too few lanes are left (never once the queue is exhausted:
            // the remaining rays just run to completion)
            if (TRIG > 0) {
                if (!exhausted && *(volatile int*)&P.done_n >= TRIG) break;
            } else if (!exhausted && uint32_t(__popc(__activemask())) < tune.refill && *(volatile int*)&P.ready_head >= ready_n) break;
        }
        __syncwarp();
    }
    if (COUNT) {
        atomicAdd(&ctr->nodes_visited, (unsigned long long)T.n_nodes);
        atomicAdd(&ctr->tris_tested, (unsigned long long)T.n_tris);
    }
}

// --------------------------------------------------------------------------------------- wavefront: shadow rays (MPT_RENDER_NEE)
// Visibility towards the environment with the SAME intersection semantics as the continuation rays (back-face culling, masks,
// alpha test), so that light samples and BSDF samples see the same scene. Same persistent structure as k_extend (per-lane dynamic
// fetch from the shadow queue); a ray ends at the first accepted hit. Unoccluded: add the MIS-weighted contribution to the path.
template <bool TWO_LEVEL>
__global__ void __launch_bounds__(kThreads, MPT_EXTEND_MIN_BLOCKS) k_shadow(const __grid_constant__ SceneView S, PathBuffers pb, uint32_t bounce, uint32_t mask,
                                                                            ExtendTune tune, Counters* ctr) {
    const uint32_t n = ctr->scount[bounce];
    MaterialAlphaTest at{&S};
    uint2 stack_mem[kStackSize];
    Traversal<false, TWO_LEVEL, MaterialAlphaTest> T(S.accel, at, stack_mem);
    bool active = false, exhausted = false, finished = false;
    uint32_t j = 0;
    while (true) {
        if (finished && !T.found()) {
            const float4 c = ldq(pb.sq_c + j);
            const uint32_t slot = __float_as_uint(c.w);
            float4 ra = pb.rad[slot]; // one shadow ray per slot and bounce: no race
            ra.x += c.x;
            ra.y += c.y;
            ra.z += c.z;
            pb.rad[slot] = ra;
        }
        finished = false;
        const uint32_t need = __ballot_sync(0xFFFFFFFFu, !active && !exhausted);
        if (need) {
            const int leader = __ffs(need) - 1;
            uint32_t base = 0;
            if (int(lane_id()) == leader) base = atomicAdd(&ctr->sfetch[bounce], uint32_t(__popc(need)));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (!active && !exhausted) {
                j = base + __popc(need & ((1u << lane_id()) - 1u));
                if (j < n) {
                    const float4 ro = ldq(pb.sq_o + j), rd = ldq(pb.sq_d + j);
                    active = T.init(xyz(ro), xyz(rd), 0.f, kRayTMax, mask, __float_as_uint(rd.w));
                    finished = !active; // masked out: nothing can occlude
                } else {
                    exhausted = true;
                }
            }
        }
        const bool queue_empty = __any_sync(0xFFFFFFFFu, exhausted);
        if (!__any_sync(0xFFFFFFFFu, active || finished)) break;
        while (active) {
            if (T.step() || T.found()) {
                active = false;
                finished = true;
                break;
            }
            if (!queue_empty && uint32_t(__popc(__activemask())) < tune.refill) break;
        }
        __syncwarp();
    }
}

// --------------------------------------------------------------------------------------- arbitrary rays, persistent form
// mpt_trace for incoherent rays: the structure of k_extend (per-lane dynamic fetch: a lane does not wait for the slowest ray of its
// warp) over the caller's ray array; results land at the ray's own index.
template <bool COUNT, bool TWO_LEVEL>
__global__ void __launch_bounds__(kThreads, MPT_EXTEND_MIN_BLOCKS) k_trace_persistent(const __grid_constant__ SceneView S, const float4* __restrict__ rays, uint32_t n,
                                                                                      uint32_t mask, uint32_t* hits6, ExtendTune tune, Counters* ctr, uint32_t* fetch) {
    MaterialAlphaTest at{&S};
    uint2 stack_mem[kStackSize];
    Traversal<COUNT, TWO_LEVEL, MaterialAlphaTest> T(S.accel, at, stack_mem);
    bool active = false, exhausted = false, finished = false;
    uint32_t j = 0;
    while (true) {
        if (finished) {
            uint32_t* o = hits6 + 6 * size_t(j);
            if (!T.found()) {
                o[0] = o[1] = o[2] = o[3] = kMiss;
                o[4] = o[5] = 0u;
            } else {
                const float4* tp = S.accel.tris + size_t(T.hit.tri) * 3;
                o[0] = S.accel.instances[T.hit.instance].custom_index;
                o[1] = T.hit.instance;
                o[2] = __float_as_uint(__ldg(tp).w);
                o[3] = __float_as_uint(__ldg(tp + 1).w);
                o[4] = __float_as_uint(T.hit.u);
                o[5] = __float_as_uint(T.hit.v);
            }
        }
        finished = false;
        const uint32_t need = __ballot_sync(0xFFFFFFFFu, !active && !exhausted);
        if (need) {
            const int leader = __ffs(need) - 1;
            uint32_t base = 0;
            if (int(lane_id()) == leader) base = atomicAdd(fetch, uint32_t(__popc(need)));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (!active && !exhausted) {
                j = base + __popc(need & ((1u << lane_id()) - 1u));
                if (j < n) {
                    const float4 ro = __ldg(rays + 2 * size_t(j)), rd = __ldg(rays + 2 * size_t(j) + 1);
                    active = T.init(xyz(ro), xyz(rd), ro.w, rd.w, mask);
                    finished = !active;
                } else {
                    exhausted = true;
                }
            }
        }
        const bool queue_empty = __any_sync(0xFFFFFFFFu, exhausted);
        if (!__any_sync(0xFFFFFFFFu, active || finished)) break;
        while (active) {
            if (T.step()) {
                active = false;
                finished = true;
                break;
            }
            if (!queue_empty && uint32_t(__popc(__activemask())) < tune.refill) break;
        }
        __syncwarp();
    }
    if (COUNT) {
        atomicAdd(&ctr->nodes_visited, (unsigned long long)T.n_nodes);
        atomicAdd(&ctr->tris_tested, (unsigned long long)T.n_tris);
    }
}

// --------------------------------------------------------------------------------------- primary visibility, persistent form
// Camera rays in 8x4-pixel tile order (index j -> tile j / 32, pixel j % 32 inside it) through the per-lane dynamic fetch of k_extend.
template <bool COUNT, bool TWO_LEVEL>
__global__ void __launch_bounds__(kThreads, MPT_EXTEND_MIN_BLOCKS) k_primary_persistent(const __grid_constant__ SceneView S, CameraParams cam, uint32_t w, uint32_t h, uint32_t mask,
                                                                                        PackedHitInfo* out, ExtendTune tune, Counters* ctr, uint32_t* fetch) {
    const uint32_t tiles_x = (w + 7) / 8, n = tiles_x * ((h + 3) / 4) * 32u;
    MaterialAlphaTest at{&S};
    uint2 stack_mem[kStackSize];
    Traversal<COUNT, TWO_LEVEL, MaterialAlphaTest> T(S.accel, at, stack_mem);
    bool active = false, exhausted = false, finished = false;
    uint32_t pix = 0;
    while (true) {
        if (finished) {
            PackedHitInfo p;
            if (!T.found()) {
                p.instance_id = kMiss; p.primitive_id = kMiss; p.geometry_index = 0xFFFF; p.instance_index = 0xFFFF;
                p.bary = __floats2half2_rn(0.f, 0.f);
            } else {
                const float4* tp = S.accel.tris + size_t(T.hit.tri) * 3;
                p.instance_id = S.accel.instances[T.hit.instance].custom_index;
                p.primitive_id = __float_as_uint(__ldg(tp).w);
                p.geometry_index = uint16_t(__float_as_uint(__ldg(tp + 1).w));
                p.instance_index = uint16_t(T.hit.instance);
                p.bary = __floats2half2_rn(T.hit.u, T.hit.v);
            }
            out[pix] = p;
        }
        finished = false;
        const uint32_t need = __ballot_sync(0xFFFFFFFFu, !active && !exhausted);
        if (need) {
            const int leader = __ffs(need) - 1;
            uint32_t base = 0;
            if (int(lane_id()) == leader) base = atomicAdd(fetch, uint32_t(__popc(need)));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (!active && !exhausted) {
                const uint32_t j = base + __popc(need & ((1u << lane_id()) - 1u));
                if (j >= n) {
                    exhausted = true;
                } else {
                    const uint32_t tile = j >> 5, in = j & 31u;
                    const uint32_t x = (tile % tiles_x) * 8 + (in & 7u), y = (tile / tiles_x) * 4 + (in >> 3);
                    if (x < w && y < h) { // else: padding of a partial tile, the lane asks again at the next refill
                        pix = y * w + x;
                        f3 o, d;
                        float tmin, tmax;
                        camera_ray(cam, x, y, w, h, o, d, tmin, tmax);
                        active = T.init(o, d, tmin, tmax, mask);
                        finished = !active;
                    }
                }
            }
        }
        const bool queue_empty = __any_sync(0xFFFFFFFFu, exhausted);
        if (!__any_sync(0xFFFFFFFFu, active || finished || !exhausted)) break; // idle lanes that drew padding pixels fetch again
        while (active) {
            if (T.step()) {
                active = false;
                finished = true;
                break;
            }
            if (!queue_empty && uint32_t(__popc(__activemask())) < tune.refill) break;
        }
        __syncwarp();
    }
    if (COUNT) {
        atomicAdd(&ctr->nodes_visited, (unsigned long long)T.n_nodes);
        atomicAdd(&ctr->tris_tested, (unsigned long long)T.n_tris);
    }
}

// --------------------------------------------------------------------------------------- closest-hit queries, pooled form
// mpt_trace (MODE 0: the caller's ray array, results at the ray's own index), mpt_primary_visibility (MODE 1: camera rays in 8x4-pixel
// tile order, PackedHitInfo per pixel) and the shadow rays of MPT_RENDER_NEE (MODE 2: the shadow queue of the bounce, a ray ends at the
// first accepted hit, an unoccluded one adds its contribution to the path) through the per-warp ray pool of k_extend_pooled: converged
// retire / refill / ray setup at full width, a finishing lane swaps its record for a READY one, re-convergence on 32 DONE records.
// Modes 0 and 1: the records carry the ray's own interval (TraversalP<..., TMIN_ZERO = false>): b.w = tmin, d.w = tmax, a.w = the ray /
// pixel index. Mode 2: continuation-ray records (b.w = the triangle the ray leaves, a.w = the shadow-queue index).
constexpr int kQueryTrace = 0, kQueryPrimary = 1, kQueryShadow = 2;
// Single-level scenes without counters: 8 blocks per SM / 64 registers like the plain extend kernel (primary visibility +1.7 %, incoherent
// rays +0.4 %, NEE frame +0.7 % against 7 blocks / 72 registers, `profiles/r02z_q8_ab.txt`); the other variants keep 7.
#ifndef MPT_QUERY_PLAIN_MIN_BLOCKS
#define MPT_QUERY_PLAIN_MIN_BLOCKS 8
#endif
template <bool COUNT, bool TWO_LEVEL, int MODE>
__global__ void __launch_bounds__(kThreads, (!COUNT && !TWO_LEVEL) ? MPT_QUERY_PLAIN_MIN_BLOCKS : MPT_EXTEND_MIN_BLOCKS)
k_query_pooled(const __grid_constant__ SceneView S, const float4* __restrict__ rays, const float4* __restrict__ rays_d, const float4* __restrict__ contrib,
               CameraParams cam, uint32_t w, uint32_t h, uint32_t n_host, const uint32_t* __restrict__ n_dev, uint32_t mask, void* __restrict__ out, Counters* ctr,
               uint32_t* fetch) {
    constexpr bool PRIMARY = MODE == kQueryPrimary, SHADOW = MODE == kQueryShadow;
    const uint32_t n = n_dev ? *n_dev : n_host; // the shadow queue's length only exists on the device
    constexpr int NE = 32 + MPT_POOL_READY, kWarps = kThreads / 32, TRIG = MPT_POOL_TRIGGER > 0 ? MPT_POOL_TRIGGER : MPT_POOL_READY;
    using Trav = TraversalP<COUNT, TWO_LEVEL, MaterialAlphaTest, 0, kThreads, SHADOW>;
    using Entry = PoolEntry;
    struct Pool {
        Entry e[NE];
        uint8_t ready[NE], done[NE];
        int ready_head, ready_n, done_n, pad;
    };
    __shared__ Pool s_pool[kWarps];
    __shared__ uint2 s_stack[1];
    const uint32_t lane = lane_id();
    const uint32_t lt = (1u << lane) - 1u;
    Pool& P = s_pool[threadIdx.x >> 5];
    uint2 lstack[kStackSize];
    uint4 pstack[kPrimStackSize];
    MaterialAlphaTest at{&S};
    Trav T(S.accel, at, s_stack, lstack, pstack);
    const uint32_t scene_flags = TWO_LEVEL ? 0u : S.accel.instances[0].mask_flags;
    const bool scene_masked = !TWO_LEVEL && ((scene_flags & 0xFFu) & mask) == 0u; // the only instance is masked out: every ray misses
    constexpr uint32_t kNoRay = 0xFFFFFFFFu;
    const uint32_t tiles_x = (w + 7) / 8;

    // the result of one ray: hits6 of mpt_trace (include/mpt.h) or the PackedHitInfo of gbuffer.slang:40-61
    auto store = [&](uint32_t tag, uint32_t tri, uint32_t instance, float u, float v) {
        if (SHADOW) { // unoccluded: the MIS-weighted light sample reaches the path (one shadow ray per slot and bounce: no race)
            if (tri == kMiss) {
                const float4 c = ldq(contrib + tag);
                float4* rad = static_cast<float4*>(out);
                const uint32_t slot = __float_as_uint(c.w);
                float4 ra = rad[slot];
                ra.x += c.x;
                ra.y += c.y;
                ra.z += c.z;
                rad[slot] = ra;
            }
            return;
        }
        uint32_t custom = kMiss, prim = kMiss, geom = kMiss;
        if (tri != kMiss) {
            const float4* tp = S.accel.tris + size_t(tri) * 3;
            custom = S.accel.instances[instance].custom_index;
            prim = __float_as_uint(__ldg(tp).w);
            geom = __float_as_uint(__ldg(tp + 1).w);
        }
        if (PRIMARY) {
            PackedHitInfo p;
            p.instance_id = custom; p.primitive_id = prim;
            p.geometry_index = uint16_t(tri != kMiss ? geom : 0xFFFFu);
            p.instance_index = uint16_t(tri != kMiss ? instance : 0xFFFFu);
            p.bary = tri != kMiss ? __floats2half2_rn(u, v) : __floats2half2_rn(0.f, 0.f);
            static_cast<PackedHitInfo*>(out)[tag] = p;
        } else {
            uint32_t* o = static_cast<uint32_t*>(out) + 6 * size_t(tag);
            o[0] = custom; o[1] = tri != kMiss ? instance : kMiss; o[2] = prim; o[3] = geom;
            o[4] = tri != kMiss ? __float_as_uint(u) : 0u;
            o[5] = tri != kMiss ? __float_as_uint(v) : 0u;
        }
    };

    for (int i = int(lane); i < NE; i += 32) {
        P.e[i].a.w = __uint_as_float(kNoRay);
        P.done[i] = uint8_t(i);
    }
    if (lane == 0) { P.done_n = NE; P.ready_head = 0; P.ready_n = 0; }
    __syncwarp();

    bool active = false, exhausted = false;
    uint32_t my = 0;
    int ready_n = 0;
    while (true) {
        // ======================================================= converged: retire DONE records, refill them, hand READY records to idle lanes
        const int nd_all = P.done_n;
        const int nd = (!exhausted && nd_all > 32) ? (nd_all & ~31) : nd_all; // whole rounds of 32 (the rest stays DONE) unless the rays ran out
        uint32_t base = 0;
        if (!exhausted && nd > 0) {
            if (lane == 0) base = atomicAdd(fetch, uint32_t(nd));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
        }
        int nready = 0;
        {   // READY records nobody has taken yet move to the front of the list
            const int rh = min(P.ready_head, ready_n), left = ready_n - rh;
            const uint8_t v = int(lane) < left ? P.ready[rh + int(lane)] : uint8_t(0);
            __syncwarp();
            if (int(lane) < left) P.ready[lane] = v;
            nready = left;
        }
        int nrq = 0; // records whose ray needed no traversal (padding pixel of a partial tile, masked-out scene): straight back onto the DONE list
#pragma unroll 1
        for (int r = 0; r * 32 < nd; r++) {
            const int i = r * 32 + int(lane);
            const bool have = i < nd;
            const uint32_t e = have ? uint32_t(P.done[i]) : 0u;
            Entry& E = P.e[e];
            const uint32_t tag = have ? __float_as_uint(E.a.w) : kNoRay;
            if (tag != kNoRay) {
                const float4 hb = E.b;
                store(tag, __float_as_uint(hb.x), __float_as_uint(hb.y), hb.z, hb.w);
            }
            bool filled = false, consumed = false;
            if (have) {
                const uint32_t jn = base + uint32_t(i);
                if (!exhausted && jn < n) {
                    consumed = true;
                    float4 ro, rd;
                    uint32_t rtag = jn;
                    bool valid = true;
                    if (PRIMARY) {
                        const uint32_t tile = jn >> 5, in = jn & 31u;
                        const uint32_t x = (tile % tiles_x) * 8 + (in & 7u), y = (tile / tiles_x) * 4 + (in >> 3);
                        valid = x < w && y < h; // else: padding of a partial tile
                        rtag = y * w + x;
                        f3 o = mk3(0.f), d = mk3(0.f, 0.f, 1.f);
                        float tmin = 0.f, tmax = 0.f;
                        if (valid) camera_ray(cam, x, y, w, h, o, d, tmin, tmax);
                        ro = make_float4(o.x, o.y, o.z, tmin);
                        rd = make_float4(d.x, d.y, d.z, tmax);
                    } else if (SHADOW) {
                        ro = ldq(rays + jn);
                        rd = ldq(rays_d + jn);
                    } else {
                        ro = __ldg(rays + 2 * size_t(jn));
                        rd = __ldg(rays + 2 * size_t(jn) + 1);
                    }
                    if (valid && scene_masked) {
                        store(rtag, kMiss, kMiss, 0.f, 0.f);
                    } else if (valid) {
                        RaySetup R;
                        setup_ray(R, xyz(ro), xyz(rd));
                        const uint32_t oct = 7u - (R.perm >> 24);
                        E.a = make_float4(ro.x, ro.y, ro.z, __uint_as_float(rtag));
                        E.b = make_float4(R.idir.x, R.idir.y, R.idir.z, SHADOW ? rd.w : ro.w); // the triangle the shadow ray leaves | tmin
                        E.c = make_float4(R.shear.Sx, R.shear.Sy, R.shear.Sz,
                                          __uint_as_float(uint32_t(R.shear.kx) | (uint32_t(R.shear.ky) << 2) | (uint32_t(R.shear.kz) << 4) | (oct << 8)));
                        E.d = rd;                                               // direction | tmax
                        filled = true;
                    }
                }
                if (!filled) E.a.w = __uint_as_float(kNoRay);
            }
            const uint32_t fm = __ballot_sync(0xFFFFFFFFu, filled);
            if (filled) P.ready[nready + __popc(fm & lt)] = uint8_t(e);
            nready += __popc(fm);
            // a record that consumed a ray without being filled goes round again (positions below the ones this and later rounds read)
            const bool rq = consumed && !filled;
            const uint32_t qm = __ballot_sync(0xFFFFFFFFu, rq);
            if (rq) P.done[nrq + __popc(qm & lt)] = uint8_t(e);
            nrq += __popc(qm);
        }
        if (!exhausted && nd > 0 && base + uint32_t(nd) >= n) exhausted = true;
        __syncwarp();
        const uint32_t idle = __ballot_sync(0xFFFFFFFFu, !active);
        const int k = __popc(idle & lt);
        if (!active && k < nready) {
            my = P.ready[k];
            T.start(&P.e[my], mask, scene_flags, SHADOW);
            active = true;
        }
        const int rem = nd_all - nd; // < 32: DONE records beyond the whole rounds, kept for the next phase behind the re-queued ones
        {
            const uint8_t v = int(lane) < rem ? P.done[nd + int(lane)] : uint8_t(0);
            __syncwarp();
            if (int(lane) < rem) P.done[nrq + int(lane)] = v;
        }
        if (lane == 0) {
            P.done_n = nrq + rem;
            P.ready_n = nready;
            P.ready_head = min(__popc(idle), nready);
        }
        ready_n = nready;
        __syncwarp();
        if (!__any_sync(0xFFFFFFFFu, active)) {
            if (exhausted && nrq + rem == 0) break; // the rays have run out and every record is retired
            continue;
        }
        // ======================================================= divergent: traverse; a finished lane swaps its record for a READY one
        uint32_t tick = 0u;
        while (active) {
            if (T.step(TWO_LEVEL ? tick++ : 0u) || (SHADOW && T.found())) { // a shadow ray ends at the first accepted hit
                Entry& E = P.e[my];
                const bool f = T.found();
                E.b = make_float4(__uint_as_float(f ? T.hit.tri : kMiss), __uint_as_float(T.hit.instance), T.hit.u, T.hit.v);
                P.done[atomicAdd(&P.done_n, 1)] = uint8_t(my);
                const int kk = atomicAdd(&P.ready_head, 1);
                if (kk < ready_n) {
                    my = P.ready[kk];
                    T.start(&P.e[my], mask, scene_flags, SHADOW);
                } else {
                    active = false;
                    break;
                }
            }
            if (!exhausted && *(volatile int*)&P.done_n >= TRIG) break;
        }
        __syncwarp();
    }
    if (COUNT) {
        atomicAdd(&ctr->nodes_visited, (unsigned long long)T.n_nodes);
        atomicAdd(&ctr->tris_tested, (unsigned long long)T.n_tris);
    }
}


// --------------------------------------------------------------------------------------- wavefront: shade after extend
template <uint32_t FEAT, bool NEE>
__global__ void __launch_bounds__(kThreads, FEAT == 0u ? MPT_SHADE_MIN_BLOCKS : 1) k_shade_bounce(const __grid_constant__ SceneView S, RenderConst rc, PathBuffers pb, uint32_t bounce, Counters* ctr) {
    const uint32_t n = ctr->hcount[bounce];
    const uint32_t out = bounce & 1u;
    const bool more = int32_t(bounce) + 1 < rc.max_path_length; // with_material, render_pt.slang:80
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (uint32_t base = warp * 32u; base < n; base += warps_total * 32u) {
        const uint32_t j = base + lane_id();
        const bool valid = j < n;
        bool alive = false, lit = false;
        f3 no, nd, throughput, medium = mk3(0.f), so, sd, sc;
        float ior = 1.f, pdf = 0.f;
        uint32_t slot = 0, origin_tri = kMiss;
        if (valid) {
            const uint4 hw = ldq(pb.hq_hit + j);
            const float4 rd = ldq(pb.hq_dir + j);
            HitRec hr;
            hr.tri = hw.x;
            origin_tri = hw.x;
            hr.instance = hw.y;
            hr.u = __uint_as_float(hw.z);
            hr.v = __uint_as_float(hw.w);
            slot = __float_as_uint(rd.w);
            const float4 th = pb.thr[slot];
            const float4 ra = pb.rad[slot];
            throughput = mk3(th.x, th.y, th.z);
            ior = th.w;
            const f3 d = xyz(rd);
            f3 o = mk3(0.f);
            if (FEAT & F_TRANSMISSION) {
                const float4 m = pb.med[slot];
                medium = mk3(m.x, m.y, m.z);
                o = xyz(ldq(pb.hq_org + j));
            }
            ShadingPoint sp;
            surface_shading_point<FEAT>(S, o, d, hr, ior, more, sp);
            if (!sp.has_material && !more) {} // emission-only vertex
            // Beer-Lambert over the segment just travelled, render_pt.slang:84 (Vacuum: mu_t = 0 -> exactly 1)
            if (FEAT & F_TRANSMISSION) {
                if (medium.x != 0.f || medium.y != 0.f || medium.z != 0.f) throughput *= vexp(-length(o - sp.pos) * medium);
            }
            const f3 radiance = mk3(ra.x, ra.y, ra.z) + throughput * sp.emission; // :86
            Rng rng = make_pcg(__float_as_uint(ra.w));
            if (NEE && more) lit = nee_environment<FEAT>(S, sp, d, rng.state, throughput, medium, so, sd, sc);
            if (more) alive = scatter<FEAT>(sp, d, rng, throughput, ior, medium, no, nd, pdf);
            pb.rad[slot] = make_float4(radiance.x, radiance.y, radiance.z, __uint_as_float(rng.state));
            if (alive) {
                pb.thr[slot] = make_float4(throughput.x, throughput.y, throughput.z, ior);
                if (FEAT & F_TRANSMISSION) pb.med[slot] = make_float4(medium.x, medium.y, medium.z, 0.f);
            }
        }
        const uint32_t qi = queue_push_index(alive, &ctr->qcount[bounce + 1]);
        if (alive) {
            pb.q_o[out][qi] = make_float4(no.x, no.y, no.z, __uint_as_float(slot));
            pb.q_d[out][qi] = make_float4(nd.x, nd.y, nd.z, NEE ? pdf : __uint_as_float(origin_tri));
        }
        if (NEE) {
            const uint32_t si = queue_push_index(lit, &ctr->scount[bounce + 1]);
            if (lit) {
                pb.sq_o[si] = make_float4(so.x, so.y, so.z, 0.f);
                pb.sq_d[si] = make_float4(sd.x, sd.y, sd.z, __uint_as_float(origin_tri));
                pb.sq_c[si] = make_float4(sc.x, sc.y, sc.z, __uint_as_float(slot));
            }
        }
    }
}

// --------------------------------------------------------------------------------------- accumulate
// The plain path of the Accumulate node (include/merian-graph/nodes/accumulate/accumulate.slang:103-255 with
// USE_MOTION_VECTORS = false, firefly filter and adaptive alpha off): same pixel of the previous frame, optional GBuffer
// validation of that pixel (GBUFFER_CHECK_ALWAYS without the stochastic neighbourhood search), history length counter,
// result = lerp(prev, src, max(1 - alpha, 1 / history)). Every product and sum is rounded separately so the image
// is bit-identical to the oracle's.
struct AccumArgs {
    const float4* prev_out;
    const uint2* prev_history; // .x = enc_normal, .y = f16 linear_z | f16 length << 16 (accumulate.slang:49-52)
    float4* out;
    uint2* history;
    const uint2* gbuf_tex0;    // .x = enc_normal (gbuffer.slang:79), or NULL
    const float4* gbuf_tex2;   // .z = linear_z, .w = delta_z (gbuffer.slang:81), or NULL
    float alpha, max_history, hard_clamp, normal_reject_cos, z_reject_percent;
    uint32_t clear, flags, enabled;
};
MPT_DEV float lerp_rn(float a, float b, float s) { return __fadd_rn(__fmul_rn(a, __fsub_rn(1.0f, s)), __fmul_rn(b, s)); }
MPT_DEV float luminance_rn(float r, float g, float b) { return __fadd_rn(__fadd_rn(__fmul_rn(r, 0.299f), __fmul_rn(g, 0.587f)), __fmul_rn(b, 0.114f)); }
MPT_DEV void accumulate_pixel(const AccumArgs& a, uint32_t pix, float4 src) {
    uint32_t enc_normal = 0u;
    float linear_z = 0.f, delta_z = 0.f;
    if (a.gbuf_tex0) enc_normal = a.gbuf_tex0[pix].x;
    if (a.gbuf_tex2) { const float4 m = a.gbuf_tex2[pix]; linear_z = fminf(m.z, 65504.0f); delta_z = m.w; } // MAX_HISTORY_Z, :45,110
    // firefly hard clamp, :120-138 (filter off: max_l = pc.firefly_hard_clamp)
    if (!(isfinite(src.x) && isfinite(src.y) && isfinite(src.z) && isfinite(src.w))) {
        src = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        const float l = luminance_rn(src.x, src.y, src.z);
        if (l > a.hard_clamp) {
            const float factor = __fdiv_rn(a.hard_clamp, l);
            src.x = __fmul_rn(src.x, factor); src.y = __fmul_rn(src.y, factor); src.z = __fmul_rn(src.z, factor);
            src.w = __fmul_rn(src.w, __fmul_rn(factor, factor));
        }
    }
    const bool f32_history = (a.flags & MPT_ACC_HISTORY_F32) != 0u;
    auto encode_history = [&](float length) {
        if (f32_history) return make_uint2(enc_normal, __float_as_uint(length));
        return make_uint2(enc_normal, uint32_t(__half_as_ushort(__float2half_rn(linear_z))) | (uint32_t(__half_as_ushort(__float2half_rn(length))) << 16));
    };
    if (a.clear || a.alpha == 0.f) { // :141-146
        a.out[pix] = src;
        a.history[pix] = encode_history(1.0f);
        return;
    }
    const uint2 ph = a.prev_history[pix];
    const float prev_length = f32_history ? __uint_as_float(ph.y) : __half2float(__ushort_as_half((unsigned short)(ph.y >> 16)));
    bool found = true;
    if (a.flags & MPT_ACC_GBUFFER_CHECK) { // reprojection_valid, reprojection.slang:56-71
        const f3 n = decode_normal(enc_normal), pn = decode_normal(ph.x);
        const float prev_z = __half2float(__ushort_as_half((unsigned short)(ph.y & 0xFFFFu)));
        const float d = __fadd_rn(__fadd_rn(__fmul_rn(n.x, pn.x), __fmul_rn(n.y, pn.y)), __fmul_rn(n.z, pn.z));
        found = a.normal_reject_cos <= d && fabsf(__fadd_rn(__fsub_rn(linear_z, prev_z), delta_z)) <= __fmul_rn(a.z_reject_percent, fmaxf(linear_z, prev_z));
    }
    const float4 po = found ? a.prev_out[pix] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (found && isfinite(po.x) && isfinite(po.y) && isfinite(po.z) && isfinite(po.w)) { // :228-251
        const float new_history = fminf(__fadd_rn(prev_length, 1.0f), a.max_history);
        const float alpha = fmaxf(__fsub_rn(1.0f, a.alpha), __fdiv_rn(1.0f, new_history));
        a.out[pix] = make_float4(lerp_rn(po.x, src.x, alpha), lerp_rn(po.y, src.y, alpha), lerp_rn(po.z, src.z, alpha), lerp_rn(po.w, src.w, alpha));
        a.history[pix] = encode_history(new_history);
    } else {
        a.out[pix] = src;
        a.history[pix] = encode_history(1.0f);
    }
}
__global__ void __launch_bounds__(kThreads) k_accumulate(AccumArgs a, const float4* src, uint32_t npix) {
    const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix < npix) accumulate_pixel(a, pix, src[pix]);
}

// --------------------------------------------------------------------------------------- resolve
// `sum` carries the running sums between the passes of one call; on the last pass the result goes to `out`, or through
// accumulate_pixel into the accumulation buffers when `acc.enabled` (fused Accumulate: no irradiance image round trip).
__global__ void __launch_bounds__(kThreads) k_resolve(RenderConst rc, PathBuffers pb, const float4* albedo, float4* out, AccumArgs acc, Counters* ctr) {
    const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix == 0) {
        unsigned long long rays = 0, shaded = 0, shadow = 0;
        for (int b = 1; b < rc.max_path_length && b < kMaxBounces; b++) { rays += ctr->qcount[b]; shaded += ctr->hcount[b]; shadow += ctr->scount[b]; }
        atomicAdd(&ctr->rays_continuation, rays);
        atomicAdd(&ctr->rays_shadow, shadow);
        atomicAdd(&ctr->shaded_hits, shaded);
    }
    if (pix >= rc.npix) return;
    f3 demod = mk3(1.f);
    if ((rc.flags & MPT_RENDER_DEMODULATE_ALBEDO) && albedo) {
        const float4 a = albedo[pix];
        demod = vmax(mk3(a.x, a.y, a.z), 1e-2f); // gbuffer.slang:146-148
    }
    f3 irr = mk3(0.f);
    float m2 = 0.f;
    if (!rc.first_pass) {
        const float4 prev = out[pix];
        irr = mk3(prev.x, prev.y, prev.z);
        m2 = prev.w;
    }
    for (uint32_t k = 0; k < rc.pass_k; k++) {
        const float4 r = pb.rad[k * rc.npix + pix];
        const f3 contrib = mk3(r.x, r.y, r.z) / demod;
        if (finite3(contrib)) { // !any(isnan) && !any(isinf), render_pt.slang:117
            irr += contrib;
            const float l = yuv_luminance(contrib);
            m2 += l * l;
        }
    }
    if (rc.last_pass && rc.spp > 0 && !(rc.flags & MPT_RENDER_OUTPUT_SUM)) {
        irr = irr / float(rc.spp);
        m2 /= float(rc.spp);
    }
    if (rc.last_pass && acc.enabled) accumulate_pixel(acc, pix, make_float4(irr.x, irr.y, irr.z, m2));
    else out[pix] = make_float4(irr.x, irr.y, irr.z, m2);
}

// --------------------------------------------------------------------------------------- ray setup, for tests
// setup_ray of one direction per thread: the shear constants of the watertight test (which must be IEEE quotients: they define the
// barycentrics bit for bit), kx | ky << 2 | kz << 4, the visiting-order word and the two forms of the inverse direction.
__global__ void k_debug_ray_setup(const float* __restrict__ dirs, uint64_t n, float* __restrict__ shear, uint32_t* __restrict__ k_perm, float* __restrict__ idir) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const f3 d = mk3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
    RaySetup R, A;
    setup_ray<false>(R, mk3(0.f), d);
    setup_ray<true>(A, mk3(0.f), d);
    shear[3 * i] = R.shear.Sx; shear[3 * i + 1] = R.shear.Sy; shear[3 * i + 2] = R.shear.Sz;
    k_perm[2 * i] = uint32_t(R.shear.kx) | (uint32_t(R.shear.ky) << 2) | (uint32_t(R.shear.kz) << 4);
    k_perm[2 * i + 1] = R.perm;
    idir[6 * i] = R.idir.x; idir[6 * i + 1] = R.idir.y; idir[6 * i + 2] = R.idir.z;
    idir[6 * i + 3] = A.idir.x; idir[6 * i + 4] = A.idir.y; idir[6 * i + 5] = A.idir.z;
}

// --------------------------------------------------------------------------------------- BSDF property suite on the device
// tests/bsdf/bsdf-checks.slang:30-135 restated: one thread, XorShift32 stream, Welford means, 256x512 quadrature.
template <uint32_t FEAT> MPT_DEV bool check_sample(const Bsdf& b, f3 wi, Rng& rng, f3& wo, float& pdf, f3& w) {
    return bsdf_sample_eval<FEAT>(b, wi, rng, wo, pdf, w);
}
template <uint32_t FEAT> __global__ void k_bsdf_check(mpt_bsdf_check_params P, float* out) {
    Bsdf b;
    b.gltf = 1u;
    b.base = mk3(P.albedo[0], P.albedo[1], P.albedo[2]);
    b.ax = P.alpha; b.ay = P.alpha; b.alpha = P.alpha;
    b.metallic = 0.f; b.f0 = P.ior_f0; b.eta = 1.0f / 1.5f; b.tw = 0.f;
    b.coat_w = 0.f; b.coat_alpha = sq(kMinRoughness);
    b.sheen_col = mk3(0.f); b.sheen_alpha = 0.f; b.max_sheen = 0.f;
    b.iri = 0.f; b.iri_ior = 1.3f; b.iri_thick = 400.0f;
    switch (P.bsdf) {
    case 0: b.gltf = 0u; break;                                                    // Lambert
    case 1: b.metallic = 1.f; b.base = mk3(1.f); break;                            // GGX = conductor with F0 = 1
    case 2: b.metallic = 1.f; b.base = mk3(1.f); b.ay = P.alpha_bitangent; break;  // anisotropic GGX
    case 4: b.iri = P.iridescence; break;                                          // IridescentFresnelMix<GGX, Lambert>, eta 1/1.5
    case 5: b.tw = 1.f; b.f0 = 0.04f; b.iri = P.iridescence; break;                // IridescentRoughDielectric
    case 8: b.metallic = 1.f; b.base = mk3(P.f0[0], P.f0[1], P.f0[2]); break;      // Conductor: Schlick(f0) * GGX
    case 3:                                                                        // FresnelMix<GGX, Lambert> is the clearcoat construction of the nest
        b.coat_w = 1.f; b.coat_alpha = P.alpha; b.eta = 1.0f;                      // (GltfCoated, f0 = 0.04): full coat over a dielectric whose own
        break;                                                                     // Fresnel term is 0 (eta = 1), i.e. over plain Lambert
    case 9:                                                                        // SheenLayer<dielectric>: sheen colour = albedo over the same base
        b.sheen_col = b.base; b.sheen_alpha = P.alpha; b.max_sheen = fmaxf(b.base.x, fmaxf(b.base.y, b.base.z));
        break;
    case 7:                                                                        // the whole GltfBSDF nest (no transmission: reciprocity holds)
        b.metallic = 0.5f; b.ay = P.alpha_bitangent > 0.f ? P.alpha_bitangent : P.alpha; b.coat_w = 0.5f; b.iri = P.iridescence;
        b.sheen_col = 0.5f * b.base; b.sheen_alpha = P.alpha; b.max_sheen = 0.5f * fmaxf(b.base.x, fmaxf(b.base.y, b.base.z));
        break;
    default: break;
    }
    const f3 wi = mk3(P.wi[0], P.wi[1], P.wi[2]);
    Rng rng;
    rng.state = P.seed | 1u;
    rng.xorshift = true;
    f3 mean_w = mk3(0.f), m2_w = mk3(0.f);
    float mean_c = 0.f, m2_c = 0.f, max_w_err = 0.f, max_p_err = 0.f;
    uint32_t valid = 0;
    for (uint32_t i = 0; i < P.n_samples; ++i) {
        f3 weight = mk3(0.f), wo;
        float c = 0.f, pdf;
        f3 w;
        if (check_sample<FEAT>(b, wi, rng, wo, pdf, w)) {
            valid++;
            weight = w;
            c = wo.z / pdf;
            f3 e;
            float pr;
            bsdf_eval_pdf<FEAT>(b, wi, wo, e, pr);
            const float pe = fabsf(pdf - pr) / (fmaxf(fabsf(pdf), fabsf(pr)) + 1e-6f);
            max_p_err = isfinite(pe) ? fmaxf(max_p_err, pe) : 1e30f;
            if (pr > 0.f) {
                const float we = length(w - e / pr);
                max_w_err = isfinite(we) ? fmaxf(max_w_err, we) : 1e30f;
            }
        }
        const float inv_i = 1.0f / float(i + 1);
        const f3 dw = weight - mean_w;
        mean_w += dw * inv_i;
        m2_w += dw * (weight - mean_w);
        const float dc = c - mean_c;
        mean_c += dc * inv_i;
        m2_c += dc * (c - mean_c);
    }
    const float inv_n = 1.0f / float(P.n_samples);
    const uint32_t NT = 256, NP = 512;
    const float dtheta = (kPi * 0.5f) / float(NT), dphi = kTwoPi / float(NP);
    float pdf_int = 0.f, max_recip = 0.f;
    for (uint32_t ti = 0; ti < NT; ++ti) {
        const float theta = (float(ti) + 0.5f) * dtheta, ct = cosf(theta), st = sinf(theta);
        for (uint32_t pi = 0; pi < NP; ++pi) {
            const float phi = (float(pi) + 0.5f) * dphi;
            const f3 wo = mk3(st * cosf(phi), st * sinf(phi), ct);
            f3 e1, e2;
            float p1, p2;
            bsdf_eval_pdf<FEAT>(b, wi, wo, e1, p1);
            bsdf_eval_pdf<FEAT>(b, wo, wi, e2, p2);
            pdf_int += p1 * st * dtheta * dphi;
            const f3 fio = e1 / wo.z, foi = e2 / wi.z;
            const f3 dd = mk3(fabsf(fio.x - foi.x), fabsf(fio.y - foi.y), fabsf(fio.z - foi.z));
            const f3 mm = mk3(fmaxf(fabsf(fio.x), fabsf(foi.x)), fmaxf(fabsf(fio.y), fabsf(foi.y)), fmaxf(fabsf(fio.z), fabsf(foi.z))) + 1e-6f;
            const float r = fmaxf(fmaxf(dd.x / mm.x, dd.y / mm.y), dd.z / mm.z);
            max_recip = isfinite(r) ? fmaxf(max_recip, r) : 1e30f;
        }
    }
    out[0] = mean_w.x; out[1] = mean_w.y; out[2] = mean_w.z;
    out[3] = sqrtf(m2_w.x) * inv_n; out[4] = sqrtf(m2_w.y) * inv_n; out[5] = sqrtf(m2_w.z) * inv_n;
    out[6] = mean_c; out[7] = sqrtf(m2_c) * inv_n; out[8] = float(valid);
    out[9] = max_w_err; out[10] = max_p_err; out[11] = pdf_int; out[12] = max_recip;
    out[13] = out[14] = out[15] = 0.f; // get_albedo is a denoiser guide, not on the pt path
}

} // namespace

// ======================================================================================= host side
namespace {

std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    cudaError_t ensure(size_t n) {
        if (n <= bytes) return cudaSuccess;
        release();
        if (n == 0) return cudaSuccess;
        const cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
    cudaError_t upload(const void* src, size_t n) {
        const cudaError_t e = ensure(n ? n : 16);
        if (e != cudaSuccess || n == 0) return e;
        return cudaMemcpy(p, src, n, cudaMemcpyHostToDevice);
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct TextureSlot {
    cudaArray_t array = nullptr;
    cudaMipmappedArray_t mips = nullptr; // instead of `array` when the texture has a mip chain
    cudaTextureObject_t obj = 0;
};

} // namespace

struct mpt_context {
    int device = 0;
    uint32_t scene_frame = 0;     // Scene::frame for the entry points without mpt_render_params (mpt_set_scene_frame)
    int sm_count = 148;
    uint32_t extend_blocks_per_sm = 7;       // resident blocks of the persistent kernels in general
    uint32_t extend_plain_blocks_per_sm = 8; // ... of the plain k_extend variant
    uint32_t shade_blocks_cap = 64; // MPT_SHADE_BLOCKS_PER_SM (tuning)
    bool extend_pooled = true;    // k_extend_pooled (per-warp ray pool in shared memory); MPT_EXTEND_LEGACY=1 selects the round-1 kernel for A/B runs
    ExtendTune tune_pooled{24u};  // pooled form: re-converge when the READY list is empty and fewer than this many lanes are traversing
    uint32_t pooled_blocks_per_sm = 8, pooled_plain_blocks_per_sm = 8;
    ExtendTune tune{14u}; // 12-14 is the optimum with the current traversal (6: -13 %, 8: -5 %, 10: -0.8 %, 16: -1.1 %, 20: -2.6 %, 24: -3.7 %)
    // Path slots in flight, in Mi (128 B of queues + state each, 160 B with a medium). Launches of a few hundred thousand rays
    // are latency-bound (one batch per warp, time = the slowest ray), so bigger passes amortise the tails of the late
    // bounces: 8 / 16 / 32 / 64 Mi slots -> 2379 / 2632 / 2780 / 2859 Msamples/s on C3. 64 Mi slots = 8.6 GB of 180 GB.
    uint32_t pool_mslots = 64;
    std::string error;
    // scene
    DevBuf accel_mem; // BVH8 nodes, then (256-byte aligned) the triangle records
    size_t accel_bytes = 0;
    size_t l2_window_bytes = 0; // > 0: persisting-L2 access-policy window over accel_mem while a frame's kernels run (MPT_L2_PERSIST)
    DevBuf tri_attr, geometries, prim_to_tri, instances, tlas_prims, tlas_boxes, xform, xform_it, xform_prev;
    uint32_t instance_count = 0; // of the uploaded scene
    uint32_t single_instance_mask = 0xFFu; // single-level scenes: the 8-bit mask of the one instance (a masked-out scene takes the legacy extend kernel)
    uint32_t query_blocks_per_sm = 0, query_plain_blocks_per_sm = 0; // resident blocks of k_query_pooled (0 = not asked yet)
    bool query_pooled = true;              // MPT_QUERY_LEGACY=1: the per-lane-fetch forms of mpt_trace / mpt_primary_visibility (A/B runs)
    DevBuf materials, material_types, texture_table, env_pixels, env_alias, env_pmf;
    bool env_sampling_ready = false;
    std::vector<TextureSlot> textures;
    SceneView view{};
    uint32_t feature_union = 0;
    bool scene_ready = false, materials_ready = false;
    mpt_bvh_info bvh_info{};
    // wavefront pool
    DevBuf q_o[2], q_d[2], hq_hit, hq_dir, hq_org, thr, rad, med, sq_o, sq_d, sq_c, counters, tex1_internal, stage_tex1, stage_albedo, stage_out;
    size_t pool_slots = 0;
    // frames in flight of mpt_render_host_begin / _end
    struct InFlight {
        cudaStream_t copy = nullptr;
        cudaEvent_t h2d = nullptr, done = nullptr, d2h = nullptr;
        DevBuf tex1, albedo, out;
        bool pending = false;
    };
    static constexpr uint32_t kMaxInFlight = 4;
    InFlight in_flight[kMaxInFlight];
    cudaStream_t compute = nullptr;
    // pageable host -> CUDA array uploads go through two pinned chunks on their own stream (staged_upload_to_array)
    cudaStream_t upload_stream = nullptr;
    void* upload_pinned[2] = {nullptr, nullptr};
    cudaEvent_t upload_done[2] = {nullptr, nullptr};
    // fetch cursors of the persistent mpt_trace / mpt_primary_visibility kernels: their own words (a ring of 64, one per launch), so
    // that a launch on a caller's stream never shares a cursor with a frame rendering on ctx->compute (whose passes clear `counters`)
    DevBuf cursors;
    uint32_t cursor_next = 0;
    // stats / profiling
    uint32_t profile_flags = 0;
    mpt_stats stats{};
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_extend, ev_shade;
    size_t ev_extend_used = 0, ev_shade_used = 0;
};

namespace {

mpt_status fail(mpt_context* c, mpt_status s, const std::string& msg) {
    if (c) c->error = msg;
    else g_create_error = msg;
    return s;
}
#define CU(call)                                                                                              \
    do {                                                                                                      \
        const cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) return fail(ctx, MPT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

// Upload / set entry points copy on the legacy stream, which does not order against ctx->compute (a non-blocking stream) or a caller's
// stream: wait for every frame this context still has on the device before any table is overwritten or freed.
cudaError_t quiesce(mpt_context* ctx) {
    if (ctx->compute) {
        const cudaError_t e = cudaStreamSynchronize(ctx->compute);
        if (e != cudaSuccess) return e;
    }
    for (auto& f : ctx->in_flight)
        if (f.copy) {
            const cudaError_t e = cudaStreamSynchronize(f.copy);
            if (e != cudaSuccess) return e;
        }
    return cudaDeviceSynchronize(); // renders issued on caller-provided streams
}

CameraParams to_device_camera(const mpt_camera& c, uint32_t frame) {
    CameraParams p;
    p.aperture_radius = c.aperture_radius;
    p.frame = frame;
    p.position = {c.position[0], c.position[1], c.position[2]};
    p.U = {c.U[0], c.U[1], c.U[2]};
    p.V = {c.V[0], c.V[1], c.V[2]};
    p.W = {c.W[0], c.W[1], c.W[2]};
    p.near_plane = c.near_plane;
    p.far_plane = c.far_plane;
    p.jitter_x = c.jitter[0];
    p.jitter_y = c.jitter[1];
    return p;
}

inline uint32_t fbits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

template <class F> void time_launch(mpt_context* ctx, bool is_extend, cudaStream_t st, F&& launch) {
    if (!(ctx->profile_flags & 2u)) { launch(); return; }
    auto& pool = is_extend ? ctx->ev_extend : ctx->ev_shade;
    size_t& used = is_extend ? ctx->ev_extend_used : ctx->ev_shade_used;
    if (used == pool.size()) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        pool.emplace_back(a, b);
    }
    cudaEventRecord(pool[used].first, st);
    launch();
    cudaEventRecord(pool[used].second, st);
    used++;
}

void harvest_events(mpt_context* ctx) {
    for (int which = 0; which < 2; which++) {
        auto& pool = which ? ctx->ev_shade : ctx->ev_extend;
        size_t& used = which ? ctx->ev_shade_used : ctx->ev_extend_used;
        for (size_t i = 0; i < used; i++) {
            cudaEventSynchronize(pool[i].second);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, pool[i].first, pool[i].second);
            if (which) { ctx->stats.shade_ms += ms; ctx->stats.shade_launches++; }
            else { ctx->stats.extend_ms += ms; ctx->stats.extend_launches++; }
        }
        used = 0;
    }
}

mpt_status ensure_pool(mpt_context* ctx, size_t slots, bool need_medium, bool need_shadow) {
    if (slots > ctx->pool_slots) {
        for (int i = 0; i < 2; i++) {
            CU(ctx->q_o[i].ensure(slots * sizeof(float4)));
            CU(ctx->q_d[i].ensure(slots * sizeof(float4)));
        }
        CU(ctx->hq_hit.ensure(slots * sizeof(uint4)));
        CU(ctx->hq_dir.ensure(slots * sizeof(float4)));
        CU(ctx->thr.ensure(slots * sizeof(float4)));
        CU(ctx->rad.ensure(slots * sizeof(float4)));
        ctx->pool_slots = slots;
    }
    if (need_medium) {
        CU(ctx->med.ensure(ctx->pool_slots * sizeof(float4)));
        CU(ctx->hq_org.ensure(ctx->pool_slots * sizeof(float4)));
    }
    if (need_shadow) {
        CU(ctx->sq_o.ensure(ctx->pool_slots * sizeof(float4)));
        CU(ctx->sq_d.ensure(ctx->pool_slots * sizeof(float4)));
        CU(ctx->sq_c.ensure(ctx->pool_slots * sizeof(float4)));
    }
    if (!ctx->counters.p) {
        CU(ctx->counters.ensure(sizeof(Counters)));
        CU(cudaMemset(ctx->counters.p, 0, sizeof(Counters)));
    }
    return MPT_OK;
}

// resident blocks per SM of k_query_pooled: the plain variants (single-level scene, no counters) and the others have different launch bounds
uint32_t query_blocks_per_sm(mpt_context* ctx, bool plain = false) {
    if (!ctx->query_blocks_per_sm) {
        int occ = 0;
        ctx->query_blocks_per_sm = (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_query_pooled<true, true, kQueryTrace>, kThreads, 0) == cudaSuccess && occ > 0)
                                       ? uint32_t(occ) : 7u;
        ctx->query_plain_blocks_per_sm = (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_query_pooled<false, false, kQueryTrace>, kThreads, 0) == cudaSuccess && occ > 0)
                                             ? uint32_t(occ) : ctx->query_blocks_per_sm;
    }
    return plain ? ctx->query_plain_blocks_per_sm : ctx->query_blocks_per_sm;
}

mpt_status next_cursor(mpt_context* ctx, cudaStream_t st, uint32_t** out) {
    constexpr uint32_t kRing = 64;
    if (!ctx->cursors.p) CU(ctx->cursors.ensure(kRing * sizeof(uint32_t)));
    uint32_t* c = ctx->cursors.as<uint32_t>() + (ctx->cursor_next++ % kRing);
    CU(cudaMemsetAsync(c, 0, sizeof(uint32_t), st));
    *out = c;
    return MPT_OK;
}

// Importance table of the lat-long environment for MPT_RENDER_NEE: texel weight = sin(theta) * (luminance, half of it taken
// from the 3x3 neighbourhood because env_eval interpolates bilinearly between texel centres) + a small uniform floor, turned
// into an alias table (Vose) so that a draw costs one 8-byte load. Built on first use from the pixels already in HBM.
mpt_status ensure_env_sampling(mpt_context* ctx) {
    if (ctx->env_sampling_ready) return MPT_OK;
    const uint32_t w = ctx->view.env_w, h = ctx->view.env_h;
    const size_t n = size_t(w) * h;
    if (ctx->view.env_type != 1u || n == 0) return MPT_OK; // no environment: nee_environment returns false
    if (n > 0xFFFFFFFFull) return fail(ctx, MPT_ERR_LIMIT, "environment map too large for the alias table");
    std::vector<float> px(n * 4);
    CU(cudaMemcpy(px.data(), ctx->env_pixels.p, n * 16, cudaMemcpyDeviceToHost));
    std::vector<float> lum(n);
    for (size_t i = 0; i < n; i++) {
        const float l = 0.299f * px[4 * i] + 0.587f * px[4 * i + 1] + 0.114f * px[4 * i + 2];
        lum[i] = (l > 0.f && std::isfinite(l)) ? l : 0.f;
    }
    std::vector<double> weight(n);
    double total = 0.0, total_sin = 0.0;
    for (uint32_t y = 0; y < h; y++) {
        const double st = std::sin((double(y) + 0.5) / double(h) * 3.14159265358979323846);
        total_sin += st * w;
        for (uint32_t x = 0; x < w; x++) {
            double nb = 0.0;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) nb += lum[size_t((y + h + dy) % h) * w + (x + w + dx) % w]; // repeat addressing, scene.cpp:786-789
            const double v = st * (0.5 * lum[size_t(y) * w + x] + 0.5 * nb / 9.0);
            weight[size_t(y) * w + x] = v;
            total += v;
        }
    }
    std::vector<float> pmf(n);
    const double floor_share = total > 0.0 ? 0.05 : 1.0; // 5 % of the draws are uniform over the sphere
    for (uint32_t y = 0; y < h; y++) {
        const double st = std::sin((double(y) + 0.5) / double(h) * 3.14159265358979323846);
        for (uint32_t x = 0; x < w; x++) {
            const size_t i = size_t(y) * w + x;
            pmf[i] = float((1.0 - floor_share) * (total > 0.0 ? weight[i] / total : 0.0) + floor_share * st / total_sin);
        }
    }
    // Vose's alias method
    std::vector<uint2> alias(n);
    std::vector<double> scaled(n);
    std::vector<uint32_t> small, large;
    small.reserve(n);
    large.reserve(n);
    double sum = 0.0;
    for (size_t i = 0; i < n; i++) sum += pmf[i];
    for (size_t i = 0; i < n; i++) {
        scaled[i] = double(pmf[i]) / sum * double(n);
        (scaled[i] < 1.0 ? small : large).push_back(uint32_t(i));
    }
    while (!small.empty() && !large.empty()) {
        const uint32_t s_i = small.back(), l_i = large.back();
        small.pop_back();
        alias[s_i] = make_uint2(fbits(float(scaled[s_i])), l_i);
        scaled[l_i] = (scaled[l_i] + scaled[s_i]) - 1.0;
        if (scaled[l_i] < 1.0) { large.pop_back(); small.push_back(l_i); }
    }
    for (uint32_t i : large) alias[i] = make_uint2(fbits(1.0f), i);
    for (uint32_t i : small) alias[i] = make_uint2(fbits(1.0f), i);
    for (size_t i = 0; i < n; i++) pmf[i] = float(double(pmf[i]) / sum);
    CU(ctx->env_alias.upload(alias.data(), n * sizeof(uint2)));
    CU(ctx->env_pmf.upload(pmf.data(), n * sizeof(float)));
    ctx->view.env_alias = ctx->env_alias.as<uint2>();
    ctx->view.env_pmf = ctx->env_pmf.as<float>();
    ctx->env_sampling_ready = true;
    return MPT_OK;
}

template <uint32_t FEAT, bool NEE>
mpt_status run_passes(mpt_context* ctx, const CameraParams& cam, const mpt_render_params& P, const PackedHitInfo* tex1, const float4* albedo,
                      float4* out, const AccumArgs& acc, cudaStream_t st) {
    const uint32_t npix = P.width * P.height;
    const uint32_t K = uint32_t(std::max<size_t>(1, std::min<size_t>(size_t(std::max(P.spp, 1)), ctx->pool_slots / npix)));
    PathBuffers pb;
    for (int i = 0; i < 2; i++) { pb.q_o[i] = ctx->q_o[i].as<float4>(); pb.q_d[i] = ctx->q_d[i].as<float4>(); }
    pb.hq_hit = ctx->hq_hit.as<uint4>();
    pb.hq_dir = ctx->hq_dir.as<float4>();
    pb.hq_org = ctx->hq_org.as<float4>();
    pb.thr = ctx->thr.as<float4>();
    pb.rad = ctx->rad.as<float4>();
    pb.med = ctx->med.as<float4>();
    pb.sq_o = ctx->sq_o.as<float4>();
    pb.sq_d = ctx->sq_d.as<float4>();
    pb.sq_c = ctx->sq_c.as<float4>();
    Counters* ctr = ctx->counters.as<Counters>();
    const bool count = ctx->profile_flags & 1u;
    const int L = std::min(P.max_path_length, kMaxBounces);
    const uint32_t persistent_blocks = uint32_t(ctx->sm_count) * ctx->extend_blocks_per_sm;
    // grid-stride shade kernels: one resident wave of THEIR occupancy (8 blocks/SM for core glTF), not of the extend kernel's
    int shade_occ = 0;
    const uint32_t shade_blocks = (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&shade_occ, k_shade_bounce<FEAT, NEE>, kThreads, 0) == cudaSuccess && shade_occ > 0)
                                      ? uint32_t(ctx->sm_count) * std::min(uint32_t(shade_occ), ctx->shade_blocks_cap)
                                      : persistent_blocks;
    const uint32_t pix_blocks = (npix + kThreads - 1) / kThreads;

    RenderConst rc{};
    rc.width = P.width; rc.height = P.height; rc.npix = npix; rc.frame = P.frame; rc.max_path_length = L;
    rc.instance_mask = P.instance_mask; rc.flags = P.flags; rc.sample_offset = P.sample_offset; rc.sample_stride = P.sample_stride;
    rc.spp = P.spp;

    // persisting-L2 window over the BVH nodes + triangle records for this frame's kernels (the queues stream past it: stq / ldq)
    const bool l2_window = ctx->l2_window_bytes > 0 && ctx->accel_bytes > 0;
    if (l2_window) {
        cudaStreamAttrValue av{};
        av.accessPolicyWindow.base_ptr = ctx->accel_mem.p;
        av.accessPolicyWindow.num_bytes = std::min(ctx->accel_bytes, ctx->l2_window_bytes);
        av.accessPolicyWindow.hitRatio = float(std::min(1.0, double(ctx->l2_window_bytes) / double(ctx->accel_bytes)));
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        CU(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
    }
    const uint32_t total = uint32_t(std::max(P.spp, 0));
    uint32_t done = 0;
    bool first = true;
    do {
        const uint32_t k = std::min(K, total - done);
        rc.pass_first = done;
        rc.pass_k = k;
        rc.first_pass = first;
        rc.last_pass = (done + k >= total);
        CU(cudaMemsetAsync(ctr, 0, offsetof(Counters, rays_continuation), st));
        time_launch(ctx, false, st, [&] { k_shade_primary<FEAT, NEE><<<pix_blocks, kThreads, 0, st>>>(ctx->view, cam, rc, tex1, pb, ctr); });
        ctx->stats.kernel_launches++;
        for (int b = 1; b < L; b++) {
            const uint32_t in = uint32_t(b - 1) & 1u;
            if (NEE) { // shadow rays of the vertex shaded last: before the extend so that both add to rad[] in stream order
                const bool tl_s = ctx->view.accel.two_level != 0;
                if (ctx->query_pooled && (tl_s || (ctx->single_instance_mask & P.instance_mask & 0xFFu) != 0u)) { // (masked-out scene: see the extend launch)
                    const uint32_t qb = uint32_t(ctx->sm_count) * query_blocks_per_sm(ctx, !tl_s);
                    if (tl_s) k_query_pooled<false, true, kQueryShadow><<<qb, kThreads, 0, st>>>(ctx->view, pb.sq_o, pb.sq_d, pb.sq_c, CameraParams{}, 0u, 0u, 0u, &ctr->scount[b], P.instance_mask, pb.rad, ctr, &ctr->sfetch[b]);
                    else k_query_pooled<false, false, kQueryShadow><<<qb, kThreads, 0, st>>>(ctx->view, pb.sq_o, pb.sq_d, pb.sq_c, CameraParams{}, 0u, 0u, 0u, &ctr->scount[b], P.instance_mask, pb.rad, ctr, &ctr->sfetch[b]);
                } else if (tl_s) k_shadow<true><<<persistent_blocks, kThreads, 0, st>>>(ctx->view, pb, uint32_t(b), P.instance_mask, ctx->tune, ctr);
                else k_shadow<false><<<persistent_blocks, kThreads, 0, st>>>(ctx->view, pb, uint32_t(b), P.instance_mask, ctx->tune, ctr);
                ctx->stats.kernel_launches++;
            }
            time_launch(ctx, true, st, [&] {
                constexpr bool MED = (FEAT & F_TRANSMISSION) != 0;
                const bool tl = ctx->view.accel.two_level != 0;
                const bool plain = !count && !MED && !tl && !NEE;
                const uint32_t grid = uint32_t(ctx->sm_count) * ((ctx->extend_pooled && (tl || (ctx->single_instance_mask & P.instance_mask & 0xFFu) != 0u)) ? (plain ? ctx->pooled_plain_blocks_per_sm : ctx->pooled_blocks_per_sm)
                                                                                     : (plain ? ctx->extend_plain_blocks_per_sm : ctx->extend_blocks_per_sm));
                // (a single-level scene whose only instance the ray mask excludes: the pooled kernel's records would drop out of circulation
                // after the first rays of every warp — every ray is a miss that needs no record — so the per-lane-fetch kernel shades them)
                const bool use_pooled = ctx->extend_pooled && (tl || (ctx->single_instance_mask & P.instance_mask & 0xFFu) != 0u);
#define MPT_LAUNCH_EXTEND(C, TL)                                                                                                                          \
    do {                                                                                                                                                \
        if (use_pooled) k_extend_pooled<C, MED, TL, NEE><<<grid, kThreads, 0, st>>>(ctx->view, pb.q_o[in], pb.q_d[in], pb, uint32_t(b), P.instance_mask, ctx->tune_pooled, ctr); \
        else k_extend<C, MED, TL, NEE><<<grid, kThreads, 0, st>>>(ctx->view, pb.q_o[in], pb.q_d[in], pb, uint32_t(b), P.instance_mask, ctx->tune, ctr);              \
    } while (0)
                if (count) { if (tl) MPT_LAUNCH_EXTEND(true, true); else MPT_LAUNCH_EXTEND(true, false); }
                else { if (tl) MPT_LAUNCH_EXTEND(false, true); else MPT_LAUNCH_EXTEND(false, false); }
#undef MPT_LAUNCH_EXTEND
            });
            time_launch(ctx, false, st, [&] { k_shade_bounce<FEAT, NEE><<<shade_blocks, kThreads, 0, st>>>(ctx->view, rc, pb, uint32_t(b), ctr); });
            ctx->stats.kernel_launches += 2;
        }
        k_resolve<<<pix_blocks, kThreads, 0, st>>>(rc, pb, albedo, out, acc, ctr);
        ctx->stats.kernel_launches++;
        done += k;
        first = false;
    } while (done < total);
    CU(cudaGetLastError());
    if (l2_window) {
        cudaStreamAttrValue av{};
        av.accessPolicyWindow.num_bytes = 0;
        CU(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
    }
    ctx->stats.samples += uint64_t(npix) * total;
    return MPT_OK;
}

} // namespace

int mpt_internal_device(const mpt_context* ctx) { return ctx->device; }
mpt_status mpt_internal_fail(mpt_context* ctx, mpt_status status, const char* message) { return fail(ctx, status, message ? message : ""); }

// ======================================================================================= C-ABI
extern "C" {

MPT_API uint32_t mpt_abi_version(void) { return MPT_ABI_VERSION; }

MPT_API const char* mpt_last_error(const mpt_context* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

MPT_API mpt_status mpt_create(int device, mpt_context** out_ctx) try {
    mpt_context* ctx = nullptr;
    if (!out_ctx) return fail(nullptr, MPT_ERR_INVALID_ARGUMENT, "out_ctx is NULL");
    *out_ctx = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(nullptr, MPT_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(nullptr, MPT_ERR_INVALID_ARGUMENT, "device index out of range");
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) // the library carries sm_100a code only; there is no fallback path
        return fail(nullptr, MPT_ERR_UNSUPPORTED, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) + ", libmpt needs sm_100 (B200)");
    CU(cudaSetDevice(device));
    ctx = new mpt_context();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->view.env_type = 0;
    ctx->view.alpha_threshold = kAlphaThreshold;
    ctx->view.min_roughness = kMinRoughness;
    ctx->view.clamp_normals = 1u;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_extend<false, false, true, false>, kThreads, 0) == cudaSuccess && occ > 0) ctx->extend_blocks_per_sm = uint32_t(occ);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_extend<false, false, false, false>, kThreads, 0) == cudaSuccess && occ > 0) ctx->extend_plain_blocks_per_sm = uint32_t(occ);
    // tuning knobs for profiling sessions (defaults are the measured optimum on B200)
    auto env_u32 = [](const char* name, uint32_t def) { const char* v = getenv(name); return v ? uint32_t(strtoul(v, nullptr, 10)) : def; };
    ctx->tune.refill = env_u32("MPT_REFILL_THRESHOLD", ctx->tune.refill);
    ctx->tune_pooled.refill = env_u32("MPT_POOLED_REFILL_THRESHOLD", ctx->tune_pooled.refill);
    ctx->query_pooled = !getenv("MPT_QUERY_LEGACY");
    ctx->extend_pooled = env_u32("MPT_EXTEND_LEGACY", 0u) == 0u;
    if (env_u32("MPT_L2_PERSIST", 0u)) {
        // B200: 126 MB L2. Reserve a persisting carve-out and remember the largest window the device accepts.
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, device);
        const size_t want = size_t(env_u32("MPT_L2_PERSIST_MB", 64u)) << 20;
        const size_t carve = std::min(want, size_t(std::max(max_persist, 0)));
        if (carve > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) ctx->l2_window_bytes = std::min(carve, size_t(std::max(max_window, 0)));
        cudaGetLastError();
    }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_extend_pooled<false, false, true, false>, kThreads, 0) == cudaSuccess && occ > 0) ctx->pooled_blocks_per_sm = uint32_t(occ);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_extend_pooled<false, false, false, false>, kThreads, 0) == cudaSuccess && occ > 0) ctx->pooled_plain_blocks_per_sm = uint32_t(occ);
    ctx->pooled_blocks_per_sm = std::min(ctx->pooled_blocks_per_sm, env_u32("MPT_EXTEND_BLOCKS_PER_SM", ctx->pooled_blocks_per_sm));
    ctx->pooled_plain_blocks_per_sm = std::min(ctx->pooled_plain_blocks_per_sm, env_u32("MPT_EXTEND_BLOCKS_PER_SM", ctx->pooled_plain_blocks_per_sm));
    ctx->pool_mslots = std::max(1u, env_u32("MPT_POOL_MSLOTS", ctx->pool_mslots));
    ctx->extend_blocks_per_sm = std::min(ctx->extend_blocks_per_sm, env_u32("MPT_EXTEND_BLOCKS_PER_SM", ctx->extend_blocks_per_sm));
    ctx->extend_plain_blocks_per_sm = std::min(ctx->extend_plain_blocks_per_sm, env_u32("MPT_EXTEND_BLOCKS_PER_SM", ctx->extend_plain_blocks_per_sm));
    ctx->shade_blocks_cap = std::max(1u, env_u32("MPT_SHADE_BLOCKS_PER_SM", ctx->shade_blocks_cap));
    *out_ctx = ctx;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_LIMIT, "mpt_create: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_INVALID_ARGUMENT, std::string("mpt_create: ") + e.what());
} catch (...) {
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_INVALID_ARGUMENT, "mpt_create: unknown exception");
}

MPT_API void mpt_destroy(mpt_context* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto& t : ctx->textures) {
        if (t.obj) cudaDestroyTextureObject(t.obj);
        if (t.array) cudaFreeArray(t.array);
        if (t.mips) cudaFreeMipmappedArray(t.mips);
    }
    for (auto& f : ctx->in_flight) {
        if (f.copy) cudaStreamDestroy(f.copy);
        for (cudaEvent_t e : {f.h2d, f.done, f.d2h})
            if (e) cudaEventDestroy(e);
    }
    if (ctx->compute) cudaStreamDestroy(ctx->compute);
    if (ctx->upload_stream) cudaStreamDestroy(ctx->upload_stream);
    for (int k = 0; k < 2; k++) {
        if (ctx->upload_pinned[k]) cudaFreeHost(ctx->upload_pinned[k]);
        if (ctx->upload_done[k]) cudaEventDestroy(ctx->upload_done[k]);
    }
    for (auto* pool : {&ctx->ev_extend, &ctx->ev_shade})
        for (auto& p : *pool) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    delete ctx;
}

MPT_API mpt_status mpt_scene_upload(mpt_context* ctx, const mpt_scene_desc* d) try {
    if (!ctx || !d) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    CU(cudaSetDevice(ctx->device));
    ctx->scene_ready = false;
    // uploads use the legacy stream; a frame of mpt_render_host_begin may still be reading the old arrays on ctx->compute
    CU(quiesce(ctx));
    HostAccel H;
    {
        std::string err;
        const mpt_status bs = build_host_accel(d, H, uint32_t(kStackSize), err);
        if (bs != MPT_OK) return fail(ctx, bs, err);
    }
    const auto& nodes = H.nodes;
    const auto& tris = H.tris;
    const auto& attrs = H.attrs;
    const auto& prim_to_tri = H.prim_to_tri;
    const auto& inst = H.instances;
    const auto& tlas_prims = H.tlas_prims;
    const auto& geoms = H.geoms;
    const uint32_t root = H.root;
    const bool identity_single = H.two_level == 0;
    const uint32_t max_blas_depth = H.max_blas_depth, tlas_depth = H.tlas_depth;

    // nodes and triangle records share one allocation so that one L2 access-policy window can cover both (run_passes)
    const size_t node_bytes = (nodes.size() * sizeof(Bvh8Node) + 255) & ~size_t(255), tri_bytes = tris.size() * sizeof(HostVec4);
    CU(ctx->accel_mem.ensure(node_bytes + std::max<size_t>(tri_bytes, 16)));
    CU(cudaMemcpy(ctx->accel_mem.p, nodes.data(), nodes.size() * sizeof(Bvh8Node), cudaMemcpyHostToDevice));
    if (tri_bytes) CU(cudaMemcpy(ctx->accel_mem.as<char>() + node_bytes, tris.data(), tri_bytes, cudaMemcpyHostToDevice));
    ctx->accel_bytes = node_bytes + tri_bytes;
    CU(ctx->tri_attr.upload(attrs.data(), attrs.size() * sizeof(HostUVec4)));
    CU(ctx->prim_to_tri.upload(prim_to_tri.data(), prim_to_tri.size() * 4));
    CU(ctx->instances.upload(inst.data(), inst.size() * sizeof(HostInstanceRec)));
    CU(ctx->tlas_prims.upload(tlas_prims.data(), tlas_prims.size() * 4));
    CU(ctx->tlas_boxes.upload(H.tlas_boxes.data(), H.tlas_boxes.size() * sizeof(HostVec4)));
    CU(ctx->geometries.upload(geoms.data(), geoms.size() * sizeof(HostUVec4)));
    std::vector<float> ident(size_t(d->instance_count) * 16, 0.f);
    for (uint32_t i = 0; i < d->instance_count; i++) ident[i * 16] = ident[i * 16 + 5] = ident[i * 16 + 10] = ident[i * 16 + 15] = 1.f;
    CU(ctx->xform.upload(d->instance_transforms ? d->instance_transforms : ident.data(), size_t(d->instance_count) * 64));
    CU(ctx->xform_it.upload(d->inverse_transposed_instance_transforms ? d->inverse_transposed_instance_transforms : ident.data(), size_t(d->instance_count) * 64));

    SceneView& v = ctx->view;
    v.accel.nodes = ctx->accel_mem.as<float4>();
    v.accel.tris = reinterpret_cast<const float4*>(ctx->accel_mem.as<char>() + node_bytes);
    v.accel.instances = ctx->instances.as<InstanceRec>();
    v.accel.tlas_prims = ctx->tlas_prims.as<uint32_t>();
    v.accel.tlas_boxes = ctx->tlas_boxes.as<float4>();
    v.accel.root = root;
    v.accel.two_level = identity_single ? 0u : 1u;
    v.accel.prmt_bias = 0x47000000u;
    v.tri_attr = ctx->tri_attr.as<uint4>();
    v.geometries = ctx->geometries.as<uint4>();
    v.prim_to_tri = ctx->prim_to_tri.as<uint32_t>();
    v.inst_xform = ctx->xform.as<float4>();
    v.inst_xform_it = ctx->xform_it.as<float4>();
    v.inst_xform_prev = nullptr; // a new scene starts static
    ctx->single_instance_mask = (!H.two_level && !H.instances.empty()) ? (H.instances[0].mask_flags & 0xFFu) : 0xFFu;
    ctx->instance_count = d->instance_count;

    ctx->bvh_info.node_count = nodes.size();
    ctx->bvh_info.triangle_count = tris.size() / 3;
    ctx->bvh_info.node_bytes = nodes.size() * sizeof(Bvh8Node);
    ctx->bvh_info.triangle_bytes = tris.size() * sizeof(HostVec4);
    ctx->bvh_info.max_depth = max_blas_depth + tlas_depth;
    ctx->bvh_info.build_ms = H.build_ms;
    ctx->scene_ready = true;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_scene_upload: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_scene_upload: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_scene_upload: unknown exception");
}

MPT_API mpt_status mpt_materials_upload(mpt_context* ctx, const void* entries, uint32_t entry_size, uint32_t material_count,
                                        const mpt_material_type* types, uint32_t type_count) try {
    if (!ctx || (!entries && material_count) || !types) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    if (entry_size % 4 != 0 || entry_size < 20) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "material entry size must be a multiple of 4 and hold header + base colour");
    CU(cudaSetDevice(ctx->device));
    CU(quiesce(ctx));
    ctx->feature_union = 0;
    std::vector<uint2> tt(type_count);
    for (uint32_t i = 0; i < type_count; i++) {
        if (types[i].model > MPT_MATERIAL_GLTF) return fail(ctx, MPT_ERR_UNSUPPORTED, "unknown material model");
        tt[i] = make_uint2(types[i].model, types[i].features);
        ctx->feature_union |= types[i].features;
    }
    for (uint32_t i = 0; i < material_count; i++) {
        uint16_t tid;
        std::memcpy(&tid, (const uint8_t*)entries + size_t(i) * entry_size, 2);
        if (tid >= type_count) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "material_model_type_id out of range");
        if (types[tid].model == MPT_MATERIAL_GLTF && entry_size < 60) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "glTF material entry smaller than its core payload");
    }
    CU(ctx->materials.upload(entries, size_t(entry_size) * material_count));
    CU(ctx->material_types.upload(tt.data(), tt.size() * sizeof(uint2)));
    ctx->view.materials = ctx->materials.as<uint32_t>();
    ctx->view.material_stride_words = entry_size / 4;
    ctx->view.material_types = ctx->material_types.as<uint2>();
    ctx->materials_ready = true;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_materials_upload: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_materials_upload: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_materials_upload: unknown exception");
}

MPT_API mpt_status mpt_materials_set_hints(mpt_context* ctx, float alpha_test_threshold, float min_roughness, uint32_t clamp_normals) try {
    if (!ctx) return MPT_ERR_INVALID_ARGUMENT;
    if (!(alpha_test_threshold >= 0.f && alpha_test_threshold <= 1.f) || !(min_roughness >= 0.f && min_roughness <= 1.f))
        return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "hints out of range");
    ctx->view.alpha_threshold = alpha_test_threshold;
    ctx->view.min_roughness = min_roughness;
    ctx->view.clamp_normals = clamp_normals ? 1u : 0u;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_materials_set_hints: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_materials_set_hints: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_materials_set_hints: unknown exception");
}

} // extern "C"
namespace {
// One material texture: CUDA array (or mipmapped array) + texture object. `pixels` is host memory unless `on_device`.
// Rows of a pageable host image into a CUDA array through two pinned chunks: the CPU copies chunk k + 1 into pinned memory while the copy
// engine moves chunk k. cudaMemcpy2DToArray from pageable memory ran at 1.6 GB/s (C4: 10.8 s for 256 textures of 4096^2 RGBA8).
constexpr size_t kUploadChunkBytes = size_t(8) << 20;
mpt_status staged_upload_to_array(mpt_context* ctx, cudaArray_t array, const void* pixels, size_t src_pitch, size_t row_bytes, size_t height) {
    if (!ctx->upload_stream) {
        CU(cudaStreamCreateWithFlags(&ctx->upload_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; k++) {
            CU(cudaHostAlloc(&ctx->upload_pinned[k], kUploadChunkBytes, cudaHostAllocDefault));
            CU(cudaEventCreateWithFlags(&ctx->upload_done[k], cudaEventDisableTiming));
        }
    }
    const size_t rows_per_chunk = std::max<size_t>(1, kUploadChunkBytes / row_bytes);
    static const bool direct = std::getenv("MPT_UPLOAD_DIRECT") != nullptr; // A/B knob: the driver's own pageable path
    if (direct || row_bytes > kUploadChunkBytes) { // (a single row larger than a chunk is not worth a special path)
        CU(cudaMemcpy2DToArray(array, 0, 0, pixels, src_pitch, row_bytes, height, cudaMemcpyHostToDevice));
        return MPT_OK;
    }
    int k = 0;
    for (size_t row = 0; row < height; row += rows_per_chunk, k ^= 1) {
        const size_t rows = std::min(rows_per_chunk, height - row);
        CU(cudaEventSynchronize(ctx->upload_done[k])); // the copy that last read this chunk (a fresh event counts as complete)
        const char* src = static_cast<const char*>(pixels) + row * src_pitch;
        char* dst = static_cast<char*>(ctx->upload_pinned[k]);
        if (src_pitch == row_bytes) std::memcpy(dst, src, rows * row_bytes);
        else for (size_t r = 0; r < rows; r++) std::memcpy(dst + r * row_bytes, src + r * src_pitch, row_bytes);
        CU(cudaMemcpy2DToArrayAsync(array, 0, row, dst, row_bytes, row_bytes, rows, cudaMemcpyHostToDevice, ctx->upload_stream));
        CU(cudaEventRecord(ctx->upload_done[k], ctx->upload_stream));
    }
    CU(cudaStreamSynchronize(ctx->upload_stream));
    return MPT_OK;
}

mpt_status texture_upload_impl(mpt_context* ctx, uint32_t texture_id, uint32_t width, uint32_t height, uint32_t format, uint32_t wrap_u, uint32_t wrap_v,
                               uint32_t filter, uint32_t mipmap_mode, const void* pixels, bool on_device, size_t src_pitch, cudaStream_t st) {
    if (!ctx || !pixels || width == 0 || height == 0) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "bad texture");
    if (texture_id >= 0xFFFFu) return fail(ctx, MPT_ERR_LIMIT, "TextureID is 16 bit");
    if (format > MPT_TEX_RGBA32F || mipmap_mode > MPT_MIPMAP_LINEAR) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "unknown texture format or mipmap mode");
    CU(cudaSetDevice(ctx->device));
    CU(quiesce(ctx));
    if (ctx->textures.size() <= texture_id) ctx->textures.resize(texture_id + 1);
    TextureSlot& slot = ctx->textures[texture_id];
    if (slot.obj) { cudaDestroyTextureObject(slot.obj); slot.obj = 0; }
    if (slot.array) { cudaFreeArray(slot.array); slot.array = nullptr; }
    if (slot.mips) { cudaFreeMipmappedArray(slot.mips); slot.mips = nullptr; }
    const bool f32 = format == MPT_TEX_RGBA32F;
    const cudaChannelFormatDesc cd = f32 ? cudaCreateChannelDesc<float4>() : cudaCreateChannelDesc<uchar4>();
    const size_t pitch = size_t(width) * (f32 ? 16 : 4);
    if (src_pitch == 0) src_pitch = pitch;
    cudaResourceDesc rd{};
    cudaTextureDesc td{};
    if (mipmap_mode == MPT_MIPMAP_NONE) {
        CU(cudaMallocArray(&slot.array, &cd, width, height));
        if (on_device) {
            CU(cudaMemcpy2DToArrayAsync(slot.array, 0, 0, pixels, src_pitch, pitch, height, cudaMemcpyDeviceToDevice, st));
            CU(cudaStreamSynchronize(st));
        } else {
            const mpt_status us = staged_upload_to_array(ctx, slot.array, pixels, src_pitch, pitch, height);
            if (us != MPT_OK) return us;
        }
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = slot.array;
    } else {
        // the whole chain, built on the host the way cmd_generate_mipmaps blits it (texture_mips.h)
        std::vector<uint8_t> host;
        const void* level0 = pixels;
        if (on_device || src_pitch != pitch) {
            host.resize(pitch * height);
            CU(cudaMemcpy2D(host.data(), pitch, pixels, src_pitch, pitch, height, on_device ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost));
            level0 = host.data();
        }
        std::vector<MipLevel> levels;
        build_mip_chain(level0, width, height, f32, format == MPT_TEX_RGBA8_SRGB, levels);
        CU(cudaMallocMipmappedArray(&slot.mips, &cd, make_cudaExtent(width, height, 0), uint32_t(levels.size())));
        for (uint32_t l = 0; l < levels.size(); l++) {
            cudaArray_t la = nullptr;
            CU(cudaGetMipmappedArrayLevel(&la, slot.mips, l));
            const size_t lp = size_t(levels[l].width) * (f32 ? 16 : 4);
            const mpt_status us = staged_upload_to_array(ctx, la, levels[l].bytes.data(), lp, lp, levels[l].height);
            if (us != MPT_OK) return us;
        }
        rd.resType = cudaResourceTypeMipmappedArray;
        rd.res.mipmap.mipmap = slot.mips;
        td.mipmapFilterMode = mipmap_mode == MPT_MIPMAP_LINEAR ? cudaFilterModeLinear : cudaFilterModePoint;
        td.minMipmapLevelClamp = 0.f;
        td.maxMipmapLevelClamp = float(levels.size() - 1);
        td.maxAnisotropy = 1; // anisotropy off (gltf_scene.cpp:298)
    }
    auto am = [](uint32_t w) { return w == MPT_WRAP_CLAMP_TO_EDGE ? cudaAddressModeClamp : (w == MPT_WRAP_MIRRORED_REPEAT ? cudaAddressModeMirror : cudaAddressModeWrap); };
    td.addressMode[0] = am(wrap_u);
    td.addressMode[1] = am(wrap_v);
    td.filterMode = filter == MPT_FILTER_NEAREST ? cudaFilterModePoint : cudaFilterModeLinear;
    td.readMode = f32 ? cudaReadModeElementType : cudaReadModeNormalizedFloat;
    td.sRGB = format == MPT_TEX_RGBA8_SRGB ? 1 : 0;
    td.normalizedCoords = 1;
    CU(cudaCreateTextureObject(&slot.obj, &rd, &td, nullptr));
    std::vector<TextureRec> table(ctx->textures.size());
    for (size_t i = 0; i < table.size(); i++) { table[i].obj = ctx->textures[i].obj; table[i].valid = ctx->textures[i].obj != 0; table[i].pad = 0; }
    CU(ctx->texture_table.upload(table.data(), table.size() * sizeof(TextureRec)));
    ctx->view.textures = ctx->texture_table.as<TextureRec>();
    ctx->view.texture_count = uint32_t(table.size());
    return MPT_OK;
}
} // namespace
extern "C" {

MPT_API mpt_status mpt_mip_level(const void* level0, uint32_t width, uint32_t height, uint32_t format, uint32_t level, void* out, uint64_t capacity_bytes,
                                 uint32_t* level_width, uint32_t* level_height) try {
    mpt_context* ctx = nullptr;
    if (!level0 || !width || !height || format > MPT_TEX_RGBA32F) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "bad texture");
    if (level >= mip_level_count(width, height)) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "level beyond the chain");
    try {
        std::vector<MipLevel> levels;
        build_mip_chain(level0, width, height, format == MPT_TEX_RGBA32F, format == MPT_TEX_RGBA8_SRGB, levels);
        const MipLevel& L = levels[level];
        if (level_width) *level_width = L.width;
        if (level_height) *level_height = L.height;
        if (out) {
            if (capacity_bytes < L.bytes.size()) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "output buffer too small");
            std::memcpy(out, L.bytes.data(), L.bytes.size());
        }
    } catch (...) {
        return fail(ctx, MPT_ERR_LIMIT, "out of host memory");
    }
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_LIMIT, "mpt_mip_level: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_INVALID_ARGUMENT, std::string("mpt_mip_level: ") + e.what());
} catch (...) {
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_INVALID_ARGUMENT, "mpt_mip_level: unknown exception");
}

MPT_API mpt_status mpt_texture_upload(mpt_context* ctx, uint32_t texture_id, uint32_t width, uint32_t height, uint32_t format, uint32_t wrap_u,
                                      uint32_t wrap_v, uint32_t filter, const void* pixels) try {
    return texture_upload_impl(ctx, texture_id, width, height, format, wrap_u, wrap_v, filter, MPT_MIPMAP_NONE, pixels, false, 0, nullptr);
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_texture_upload: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_texture_upload: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_texture_upload: unknown exception");
}

MPT_API mpt_status mpt_texture_upload_ex(mpt_context* ctx, uint32_t texture_id, uint32_t width, uint32_t height, uint32_t format, uint32_t wrap_u,
                                         uint32_t wrap_v, uint32_t filter, uint32_t mipmap_mode, const void* pixels) try {
    return texture_upload_impl(ctx, texture_id, width, height, format, wrap_u, wrap_v, filter, mipmap_mode, pixels, false, 0, nullptr);
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_texture_upload_ex: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_texture_upload_ex: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_texture_upload_ex: unknown exception");
}

MPT_API mpt_status mpt_texture_upload_device(mpt_context* ctx, uint32_t texture_id, uint32_t width, uint32_t height, uint32_t format, uint32_t wrap_u,
                                             uint32_t wrap_v, uint32_t filter, uint32_t mipmap_mode, const void* d_pixels, uint64_t row_pitch_bytes,
                                             void* stream) try {
    return texture_upload_impl(ctx, texture_id, width, height, format, wrap_u, wrap_v, filter, mipmap_mode, d_pixels, true, size_t(row_pitch_bytes),
                               (cudaStream_t)stream);
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_texture_upload_device: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_texture_upload_device: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_texture_upload_device: unknown exception");
}

} // extern "C"
namespace {
mpt_status env_set_impl(mpt_context* ctx, uint32_t type, const float* to_local, float intensity, uint32_t width, uint32_t height, const float* rgba32f,
                        bool on_device);
}
extern "C" {
MPT_API mpt_status mpt_env_set(mpt_context* ctx, uint32_t type, const float* to_local, float intensity, uint32_t width, uint32_t height,
                               const float* rgba32f) try {
    return env_set_impl(ctx, type, to_local, intensity, width, height, rgba32f, false);
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_env_set: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_env_set: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_env_set: unknown exception");
}
MPT_API mpt_status mpt_env_set_device(mpt_context* ctx, uint32_t type, const float* to_local, float intensity, uint32_t width, uint32_t height,
                                      const void* d_rgba32f) try {
    return env_set_impl(ctx, type, to_local, intensity, width, height, static_cast<const float*>(d_rgba32f), true);
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_env_set_device: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_env_set_device: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_env_set_device: unknown exception");
}
} // extern "C"
namespace {
mpt_status env_set_impl(mpt_context* ctx, uint32_t type, const float* to_local, float intensity, uint32_t width, uint32_t height, const float* rgba32f,
                        bool on_device) {
    if (!ctx) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL context");
    CU(cudaSetDevice(ctx->device));
    CU(quiesce(ctx));
    if (type == MPT_ENV_EMPTY) { ctx->view.env_type = 0; return MPT_OK; }
    if (type != MPT_ENV_LATLONG) return fail(ctx, MPT_ERR_UNSUPPORTED, "only Empty and LatLong environment maps are supported");
    if (!to_local || !rgba32f || !width || !height) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "bad env map");
    if (on_device) {
        CU(ctx->env_pixels.ensure(size_t(width) * height * 16));
        CU(cudaMemcpy(ctx->env_pixels.p, rgba32f, size_t(width) * height * 16, cudaMemcpyDeviceToDevice));
    } else {
        CU(ctx->env_pixels.upload(rgba32f, size_t(width) * height * 16));
    }
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) ctx->view.env_m[r * 3 + c] = to_local[r * 4 + c];
    {   // inverse of the 3x3 part (local -> world) for the light samples
        const float* m = ctx->view.env_m;
        const double det = double(m[0]) * (double(m[4]) * m[8] - double(m[5]) * m[7]) - double(m[1]) * (double(m[3]) * m[8] - double(m[5]) * m[6]) +
                           double(m[2]) * (double(m[3]) * m[7] - double(m[4]) * m[6]);
        if (!(std::fabs(det) > 1e-12)) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "environment to_local is singular");
        const double id = 1.0 / det;
        float* o = ctx->view.env_inv;
        o[0] = float((double(m[4]) * m[8] - double(m[5]) * m[7]) * id); o[1] = float((double(m[2]) * m[7] - double(m[1]) * m[8]) * id); o[2] = float((double(m[1]) * m[5] - double(m[2]) * m[4]) * id);
        o[3] = float((double(m[5]) * m[6] - double(m[3]) * m[8]) * id); o[4] = float((double(m[0]) * m[8] - double(m[2]) * m[6]) * id); o[5] = float((double(m[2]) * m[3] - double(m[0]) * m[5]) * id);
        o[6] = float((double(m[3]) * m[7] - double(m[4]) * m[6]) * id); o[7] = float((double(m[1]) * m[6] - double(m[0]) * m[7]) * id); o[8] = float((double(m[0]) * m[4] - double(m[1]) * m[3]) * id);
    }
    ctx->env_sampling_ready = false;
    ctx->view.env_intensity = intensity;
    ctx->view.env_pixels = ctx->env_pixels.as<float4>();
    ctx->view.env_w = width;
    ctx->view.env_h = height;
    ctx->view.env_type = 1;
    return MPT_OK;
}
} // namespace
extern "C" {

MPT_API mpt_status mpt_primary_visibility(mpt_context* ctx, const mpt_camera* camera, uint32_t width, uint32_t height, uint32_t instance_mask,
                                          void* d_tex1, void* stream) try {
    if (!ctx || !camera || !d_tex1 || !width || !height) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    if (!ctx->scene_ready || !ctx->materials_ready) return fail(ctx, MPT_ERR_NOT_READY, "scene or materials not uploaded");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (!ctx->counters.p) { CU(ctx->counters.ensure(sizeof(Counters))); CU(cudaMemset(ctx->counters.p, 0, sizeof(Counters))); }
    const uint32_t tiles = ((width + 7) / 8) * ((height + 3) / 4);
    const uint32_t blocks = (tiles * 32 + kThreads - 1) / kThreads;
    const CameraParams cam = to_device_camera(*camera, ctx->scene_frame);
    const bool cnt = ctx->profile_flags & 1u, tl = ctx->view.accel.two_level != 0;
    const uint32_t persistent_blocks = uint32_t(ctx->sm_count) * ctx->extend_blocks_per_sm;
    if (blocks > persistent_blocks && !getenv("MPT_TRACE_ONE_SHOT")) { // enough pixels to fill the machine: per-lane dynamic fetch
        Counters* ctr = ctx->counters.as<Counters>();
        uint32_t* cursor = nullptr;
        { const mpt_status cs = next_cursor(ctx, st, &cursor); if (cs != MPT_OK) return cs; }
        const uint32_t query_blocks = uint32_t(ctx->sm_count) * query_blocks_per_sm(ctx, !cnt && !tl);
#define MPT_LAUNCH_PRIMARY_P(C, TL)                                                                                                                     \
    do {                                                                                                                                              \
        if (ctx->query_pooled) k_query_pooled<C, TL, kQueryPrimary><<<query_blocks, kThreads, 0, st>>>(ctx->view, nullptr, nullptr, nullptr, cam, width, height, tiles * 32u, nullptr, instance_mask, d_tex1, ctr, cursor); \
        else k_primary_persistent<C, TL><<<persistent_blocks, kThreads, 0, st>>>(ctx->view, cam, width, height, instance_mask, (PackedHitInfo*)d_tex1, ctx->tune, ctr, cursor); \
    } while (0)
        if (cnt) { if (tl) MPT_LAUNCH_PRIMARY_P(true, true); else MPT_LAUNCH_PRIMARY_P(true, false); }
        else { if (tl) MPT_LAUNCH_PRIMARY_P(false, true); else MPT_LAUNCH_PRIMARY_P(false, false); }
#undef MPT_LAUNCH_PRIMARY_P
        CU(cudaGetLastError());
        ctx->stats.kernel_launches++;
        ctx->stats.rays_primary += uint64_t(width) * height;
        return MPT_OK;
    }
#define MPT_LAUNCH_PRIMARY(C, TL) k_primary<C, TL><<<blocks, kThreads, 0, st>>>(ctx->view, cam, width, height, instance_mask, (PackedHitInfo*)d_tex1, ctx->counters.as<Counters>())
    if (cnt) { if (tl) MPT_LAUNCH_PRIMARY(true, true); else MPT_LAUNCH_PRIMARY(true, false); }
    else { if (tl) MPT_LAUNCH_PRIMARY(false, true); else MPT_LAUNCH_PRIMARY(false, false); }
#undef MPT_LAUNCH_PRIMARY
    CU(cudaGetLastError());
    ctx->stats.kernel_launches++;
    ctx->stats.rays_primary += uint64_t(width) * height;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_primary_visibility: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_primary_visibility: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_primary_visibility: unknown exception");
}

MPT_API mpt_status mpt_gbuffer(mpt_context* ctx, const mpt_camera* camera, const mpt_camera* prev_camera, uint32_t width, uint32_t height,
                               uint32_t instance_mask, const mpt_gbuffer_images* images, void* stream) try {
    if (!ctx || !camera || !images || !width || !height) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    if (!ctx->scene_ready || !ctx->materials_ready) return fail(ctx, MPT_ERR_NOT_READY, "scene or materials not uploaded");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t tiles = ((width + 7) / 8) * ((height + 3) / 4);
    const uint32_t blocks = (tiles * 32 + kThreads - 1) / kThreads;
    const CameraParams cam = to_device_camera(*camera, ctx->scene_frame), prev = to_device_camera(prev_camera ? *prev_camera : *camera, ctx->scene_frame);
    const GBufferOut out{(uint2*)images->d_tex0, (PackedHitInfo*)images->d_tex1, (float4*)images->d_tex2, (float4*)images->d_tex3, (float4*)images->d_emission};
    const bool tl = ctx->view.accel.two_level != 0;
#define MPT_LAUNCH_GBUFFER(F) do { if (tl) k_gbuffer<F, true><<<blocks, kThreads, 0, st>>>(ctx->view, cam, prev, width, height, instance_mask, out); \
                                   else k_gbuffer<F, false><<<blocks, kThreads, 0, st>>>(ctx->view, cam, prev, width, height, instance_mask, out); } while (0)
    if (ctx->feature_union == 0) MPT_LAUNCH_GBUFFER(0u);
    else if (!(ctx->feature_union & (MPT_GLTF_TRANSMISSION | MPT_GLTF_VOLUME))) MPT_LAUNCH_GBUFFER(0x3Cu);
    else MPT_LAUNCH_GBUFFER(0x3Fu);
#undef MPT_LAUNCH_GBUFFER
    CU(cudaGetLastError());
    ctx->stats.kernel_launches++;
    ctx->stats.rays_primary += uint64_t(width) * height;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_gbuffer: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_gbuffer: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_gbuffer: unknown exception");
}

MPT_API mpt_status mpt_trace(mpt_context* ctx, const void* d_rays, uint64_t n, uint32_t instance_mask, void* d_hits, void* stream) try {
    if (!ctx || (!d_rays && n) || (!d_hits && n)) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    if (!ctx->scene_ready || !ctx->materials_ready) return fail(ctx, MPT_ERR_NOT_READY, "scene or materials not uploaded");
    if (n == 0) return MPT_OK;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (!ctx->counters.p) { CU(ctx->counters.ensure(sizeof(Counters))); CU(cudaMemset(ctx->counters.p, 0, sizeof(Counters))); }
    const uint32_t blocks = uint32_t((n + kThreads - 1) / kThreads);
    const bool cnt = ctx->profile_flags & 1u, tl = ctx->view.accel.two_level != 0;
    const uint32_t persistent_blocks = uint32_t(ctx->sm_count) * ctx->extend_blocks_per_sm;
    if (n < 0xFFFFFFFFull && blocks > persistent_blocks && !getenv("MPT_TRACE_ONE_SHOT")) { // enough rays to fill the machine: per-lane dynamic fetch
        Counters* ctr = ctx->counters.as<Counters>();
        uint32_t* cursor = nullptr;
        { const mpt_status cs = next_cursor(ctx, st, &cursor); if (cs != MPT_OK) return cs; }
        const uint32_t query_blocks = uint32_t(ctx->sm_count) * query_blocks_per_sm(ctx, !cnt && !tl);
#define MPT_LAUNCH_TRACE_P(C, TL)                                                                                                                       \
    do {                                                                                                                                              \
        if (ctx->query_pooled) k_query_pooled<C, TL, kQueryTrace><<<query_blocks, kThreads, 0, st>>>(ctx->view, (const float4*)d_rays, nullptr, nullptr, CameraParams{}, 0u, 0u, uint32_t(n), nullptr, instance_mask, d_hits, ctr, cursor); \
        else k_trace_persistent<C, TL><<<persistent_blocks, kThreads, 0, st>>>(ctx->view, (const float4*)d_rays, uint32_t(n), instance_mask, (uint32_t*)d_hits, ctx->tune, ctr, cursor); \
    } while (0)
        if (cnt) { if (tl) MPT_LAUNCH_TRACE_P(true, true); else MPT_LAUNCH_TRACE_P(true, false); }
        else { if (tl) MPT_LAUNCH_TRACE_P(false, true); else MPT_LAUNCH_TRACE_P(false, false); }
#undef MPT_LAUNCH_TRACE_P
        CU(cudaGetLastError());
        ctx->stats.kernel_launches++;
        ctx->stats.rays_continuation += n;
        return MPT_OK;
    }
#define MPT_LAUNCH_TRACE(C, TL) k_trace<C, TL><<<blocks, kThreads, 0, st>>>(ctx->view, (const float4*)d_rays, n, instance_mask, (uint32_t*)d_hits, ctx->counters.as<Counters>())
    if (cnt) { if (tl) MPT_LAUNCH_TRACE(true, true); else MPT_LAUNCH_TRACE(true, false); }
    else { if (tl) MPT_LAUNCH_TRACE(false, true); else MPT_LAUNCH_TRACE(false, false); }
#undef MPT_LAUNCH_TRACE
    CU(cudaGetLastError());
    ctx->stats.kernel_launches++;
    ctx->stats.rays_continuation += n;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_trace: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_trace: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_trace: unknown exception");
}

} // extern "C"
namespace {
mpt_status render_impl(mpt_context* ctx, const mpt_camera* camera, const mpt_render_params* P, const void* d_tex1, const void* d_albedo,
                       void* d_irradiance, const AccumArgs& acc, void* stream) {
    if (!ctx || !camera || !P || !d_irradiance) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    if (!ctx->scene_ready || !ctx->materials_ready) return fail(ctx, MPT_ERR_NOT_READY, "scene or materials not uploaded");
    if (!P->width || !P->height || P->spp < 0 || P->max_path_length < 1 || P->sample_stride == 0) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "bad render parameters");
    if (P->max_path_length > kMaxBounces) return fail(ctx, MPT_ERR_LIMIT, "max_path_length above 64 (per-bounce queue counters)");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t npix = size_t(P->width) * P->height;
    // K samples per pixel in flight: about 8M path slots, at least one sample per pixel
    size_t budget = size_t(ctx->pool_mslots) << 20;
    if (budget > ctx->pool_slots) { // growing: never take more than a quarter of what is free right now
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) budget = std::max(ctx->pool_slots, std::min(budget, free_b / 4 / 160));
    }
    const size_t want = std::max<size_t>(npix, std::min<size_t>(npix * size_t(std::max(P->spp, 1)), budget));
    const size_t slots = (want / npix) * npix;
    const bool nee = (P->flags & MPT_RENDER_NEE) != 0 && ctx->view.env_type == 1u;
    const mpt_status ps = ensure_pool(ctx, slots, (ctx->feature_union & (MPT_GLTF_TRANSMISSION | MPT_GLTF_VOLUME)) != 0, nee);
    if (ps != MPT_OK) return ps;
    if (nee) {
        const mpt_status es = ensure_env_sampling(ctx);
        if (es != MPT_OK) return es;
    }
    const PackedHitInfo* tex1 = (const PackedHitInfo*)d_tex1;
    ctx->scene_frame = P->frame; // Scene::frame of this iteration: what the lens sample of a thin-lens camera is keyed on
    if (!tex1) {
        CU(ctx->tex1_internal.ensure(npix * sizeof(PackedHitInfo)));
        const mpt_status s = mpt_primary_visibility(ctx, camera, P->width, P->height, P->instance_mask, ctx->tex1_internal.p, stream);
        if (s != MPT_OK) return s;
        tex1 = ctx->tex1_internal.as<PackedHitInfo>();
    }
    const CameraParams cam = to_device_camera(*camera, P->frame);
    if (P->spp == 0) {
        CU(cudaMemsetAsync(d_irradiance, 0, npix * 16, st));
        if (acc.enabled) k_accumulate<<<uint32_t((npix + kThreads - 1) / kThreads), kThreads, 0, st>>>(acc, (const float4*)d_irradiance, uint32_t(npix));
        return MPT_OK;
    }
    // three specialisations of the shading kernels (SURVEY.md §7 "hard parts": Slang drops disabled extensions at compile time):
    // core glTF only / surface extensions / everything incl. transmission + volume (the only one that carries medium state)
#define MPT_RUN(F) (nee ? run_passes<F, true>(ctx, cam, *P, tex1, (const float4*)d_albedo, (float4*)d_irradiance, acc, st) \
                        : run_passes<F, false>(ctx, cam, *P, tex1, (const float4*)d_albedo, (float4*)d_irradiance, acc, st))
    if (ctx->feature_union == 0) return MPT_RUN(0u);
    if (!(ctx->feature_union & (MPT_GLTF_TRANSMISSION | MPT_GLTF_VOLUME))) return MPT_RUN(0x3Cu);
    return MPT_RUN(0x3Fu);
#undef MPT_RUN
}

mpt_status make_accum_args(mpt_context* ctx, const mpt_accumulate_params* A, const mpt_accumulate_buffers* B, AccumArgs& a) {
    if (!A || !B || !B->d_out || !B->d_history) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL accumulate argument");
    const bool blends = !(A->clear || A->alpha == 0.f);
    if (blends && (!B->d_prev_out || !B->d_prev_history)) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "accumulation without clear needs the previous out / history");
    if (!(A->alpha >= 0.f && A->alpha <= 1.f)) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "accumulate alpha must be in [0, 1]");
    if ((A->flags & MPT_ACC_GBUFFER_CHECK) && (A->flags & MPT_ACC_HISTORY_F32))
        return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "MPT_ACC_HISTORY_F32 stores no depth: it cannot be combined with MPT_ACC_GBUFFER_CHECK");
    if ((A->flags & MPT_ACC_GBUFFER_CHECK) && (!B->d_gbuf_tex0 || !B->d_gbuf_tex2)) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "MPT_ACC_GBUFFER_CHECK needs GBuffer tex0 and tex2");
    a.prev_out = (const float4*)B->d_prev_out;
    a.prev_history = (const uint2*)B->d_prev_history;
    a.out = (float4*)B->d_out;
    a.history = (uint2*)B->d_history;
    a.gbuf_tex0 = (const uint2*)B->d_gbuf_tex0;
    a.gbuf_tex2 = (const float4*)B->d_gbuf_tex2;
    a.alpha = A->alpha; a.max_history = A->max_history; a.hard_clamp = A->hard_clamp;
    a.normal_reject_cos = A->normal_reject_cos; a.z_reject_percent = A->z_reject_percent;
    a.clear = A->clear; a.flags = A->flags; a.enabled = 1u;
    return MPT_OK;
}
} // namespace
extern "C" {

MPT_API mpt_status mpt_render(mpt_context* ctx, const mpt_camera* camera, const mpt_render_params* P, const void* d_tex1, const void* d_albedo,
                              void* d_irradiance, void* stream) try {
    if (!d_irradiance) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    return render_impl(ctx, camera, P, d_tex1, d_albedo, d_irradiance, AccumArgs{}, stream);
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_render: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_render: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_render: unknown exception");
}

MPT_API mpt_status mpt_accumulate(mpt_context* ctx, const mpt_accumulate_params* A, const void* d_src, const mpt_accumulate_buffers* B, void* stream) try {
    if (!ctx || !d_src) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    AccumArgs a{};
    const mpt_status s = make_accum_args(ctx, A, B, a);
    if (s != MPT_OK) return s;
    if (!A->width || !A->height) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "empty image");
    CU(cudaSetDevice(ctx->device));
    const uint32_t npix = A->width * A->height;
    k_accumulate<<<(npix + kThreads - 1) / kThreads, kThreads, 0, (cudaStream_t)stream>>>(a, (const float4*)d_src, npix);
    CU(cudaGetLastError());
    ctx->stats.kernel_launches++;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_accumulate: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_accumulate: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_accumulate: unknown exception");
}

MPT_API mpt_status mpt_render_accumulate(mpt_context* ctx, const mpt_camera* camera, const mpt_render_params* P, const void* d_tex1,
                                         const void* d_albedo, const mpt_accumulate_params* A, const mpt_accumulate_buffers* B, void* stream) try {
    if (!ctx || !P) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    AccumArgs a{};
    const mpt_status s = make_accum_args(ctx, A, B, a);
    if (s != MPT_OK) return s;
    if (A->width != P->width || A->height != P->height) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "accumulate and render extents differ");
    if (P->flags & MPT_RENDER_OUTPUT_SUM) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "MPT_RENDER_OUTPUT_SUM (partial sums of one shard) cannot be accumulated: reduce first, then mpt_accumulate");
    CU(cudaSetDevice(ctx->device));
    // the running sums of a multi-pass call (and the spp = 0 image) live in a scratch image; with one pass it is never touched
    CU(ctx->stage_out.ensure(size_t(P->width) * P->height * 16));
    return render_impl(ctx, camera, P, d_tex1, d_albedo, ctx->stage_out.p, a, stream);
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_render_accumulate: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_render_accumulate: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_render_accumulate: unknown exception");
}

MPT_API mpt_status mpt_render_host(mpt_context* ctx, const mpt_camera* camera, const mpt_render_params* P, const void* h_tex1, const void* h_albedo,
                                   void* h_irradiance, void* stream) try {
    if (!ctx || !P || !h_irradiance) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t npix = size_t(P->width) * P->height;
    CU(ctx->stage_out.ensure(npix * 16));
    void* d_tex1 = nullptr;
    void* d_albedo = nullptr;
    if (h_tex1) {
        CU(ctx->stage_tex1.ensure(npix * 16));
        CU(cudaMemcpyAsync(ctx->stage_tex1.p, h_tex1, npix * 16, cudaMemcpyHostToDevice, st));
        d_tex1 = ctx->stage_tex1.p;
    }
    if (h_albedo) {
        CU(ctx->stage_albedo.ensure(npix * 16));
        CU(cudaMemcpyAsync(ctx->stage_albedo.p, h_albedo, npix * 16, cudaMemcpyHostToDevice, st));
        d_albedo = ctx->stage_albedo.p;
    }
    const mpt_status s = mpt_render(ctx, camera, P, d_tex1, d_albedo, ctx->stage_out.p, stream);
    if (s != MPT_OK) return s;
    CU(cudaMemcpyAsync(h_irradiance, ctx->stage_out.p, npix * 16, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_render_host: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_render_host: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_render_host: unknown exception");
}

// ---- wire formats at either end of the path
MPT_API mpt_status mpt_pfm_info(const char* path, uint32_t* width, uint32_t* height, uint32_t* channels) try {
    mpt_context* ctx = nullptr;
    if (!path || !width || !height || !channels) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    PfmInfo info;
    const std::string err = pfm_read_info(path, info);
    if (!err.empty()) return fail(ctx, MPT_ERR_IO, err);
    *width = uint32_t(info.width); *height = uint32_t(info.height); *channels = uint32_t(info.channels);
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_LIMIT, "mpt_pfm_info: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_INVALID_ARGUMENT, std::string("mpt_pfm_info: ") + e.what());
} catch (...) {
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_INVALID_ARGUMENT, "mpt_pfm_info: unknown exception");
}

MPT_API mpt_status mpt_pfm_load(const char* path, uint32_t desired_channels, float* out, uint64_t capacity_floats) try {
    mpt_context* ctx = nullptr;
    if (!path || !out) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    PfmInfo info;
    std::vector<float> px;
    const std::string err = pfm_load(path, int(desired_channels), info, px);
    if (!err.empty()) return fail(ctx, MPT_ERR_IO, err);
    if (px.size() > capacity_floats) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "output buffer too small for the image");
    std::memcpy(out, px.data(), px.size() * sizeof(float));
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_LIMIT, "mpt_pfm_load: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_INVALID_ARGUMENT, std::string("mpt_pfm_load: ") + e.what());
} catch (...) {
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_INVALID_ARGUMENT, "mpt_pfm_load: unknown exception");
}

MPT_API mpt_status mpt_pfm_save(const char* path, const float* data, uint32_t width, uint32_t height, uint32_t channels) try {
    mpt_context* ctx = nullptr;
    if (!path || !data) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    const std::string err = pfm_save(path, data, int(width), int(height), int(channels));
    if (!err.empty()) return fail(ctx, err.find("supports only") != std::string::npos || err.find("empty") != std::string::npos ? MPT_ERR_INVALID_ARGUMENT : MPT_ERR_IO, err);
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_LIMIT, "mpt_pfm_save: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_INVALID_ARGUMENT, std::string("mpt_pfm_save: ") + e.what());
} catch (...) {
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_INVALID_ARGUMENT, "mpt_pfm_save: unknown exception");
}

MPT_API mpt_status mpt_hdr_load(const char* path, uint32_t desired_channels, uint32_t* width, uint32_t* height, float* out, uint64_t capacity_floats) try {
    mpt_context* ctx = nullptr;
    if (!path || !width || !height) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    int w = 0, h = 0;
    std::vector<float> px;
    const std::string err = hdr_load(path, int(desired_channels), w, h, px);
    if (!err.empty()) return fail(ctx, MPT_ERR_IO, err);
    *width = uint32_t(w);
    *height = uint32_t(h);
    if (!out) return MPT_OK; // size query
    if (px.size() > capacity_floats) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "output buffer too small for the image");
    std::memcpy(out, px.data(), px.size() * sizeof(float));
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_LIMIT, "mpt_hdr_load: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_INVALID_ARGUMENT, std::string("mpt_hdr_load: ") + e.what());
} catch (...) {
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_INVALID_ARGUMENT, "mpt_hdr_load: unknown exception");
}

MPT_API mpt_status mpt_env_load(mpt_context* ctx, const char* path, const float* to_local, float intensity) try {
    if (!ctx || !path || !to_local) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    const size_t n = std::strlen(path);
    auto ends = [&](const char* e) { std::string t(n >= 4 ? path + n - 4 : ""); for (char& c : t) c = char(std::tolower((unsigned char)c)); return t == e; };
    std::vector<float> px;
    int w = 0, h = 0;
    std::string err;
    if (ends(".pfm")) { // image_load_f32 -> create_texture_from_rgba32f, scene.cpp:783-789
        PfmInfo info;
        err = pfm_load(path, 4, info, px);
        w = info.width;
        h = info.height;
    } else if (ends(".hdr")) {
        err = hdr_load(path, 4, w, h, px);
    } else {
        return fail(ctx, MPT_ERR_UNSUPPORTED, "mpt_env_load reads .pfm and .hdr lat-long maps (decode 8-bit formats on the host and call mpt_env_set)");
    }
    if (!err.empty()) return fail(ctx, MPT_ERR_IO, err);
    return mpt_env_set(ctx, MPT_ENV_LATLONG, to_local, intensity, uint32_t(w), uint32_t(h), px.data());
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_env_load: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_env_load: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_env_load: unknown exception");
}

MPT_API mpt_status mpt_image_write(mpt_context* ctx, const char* path, const void* d_rgba32f, uint32_t width, uint32_t height, void* stream) try {
    if (!ctx || !path || !d_rgba32f || !width || !height) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    CU(cudaSetDevice(ctx->device));
    std::vector<float> host(size_t(width) * height * 4);
    CU(cudaMemcpyAsync(host.data(), d_rgba32f, host.size() * sizeof(float), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    const std::string err = pfm_save(path, host.data(), int(width), int(height), 4);
    if (!err.empty()) return fail(ctx, MPT_ERR_IO, err);
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_image_write: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_image_write: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_image_write: unknown exception");
}

// Several iterations in flight (merian's graph keeps 2 by default, examples/gltf.json:39): the host->device copy of iteration i+1 and
// the device->host copy of iteration i-1 run on per-slot copy streams while iteration i renders on the context's compute stream (the
// path pool is shared, so renders are serialised among themselves).
MPT_API mpt_status mpt_render_host_begin(mpt_context* ctx, const mpt_camera* camera, const mpt_render_params* P, const void* h_tex1, const void* h_albedo,
                                         void* h_irradiance, uint32_t in_flight_index) try {
    if (!ctx || !camera || !P || !h_irradiance) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    if (in_flight_index >= mpt_context::kMaxInFlight) return fail(ctx, MPT_ERR_LIMIT, "in_flight_index must be < 4");
    CU(cudaSetDevice(ctx->device));
    mpt_context::InFlight& f = ctx->in_flight[in_flight_index];
    if (f.pending) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "this in-flight slot still holds a frame: call mpt_render_host_end first");
    if (!ctx->compute) CU(cudaStreamCreateWithFlags(&ctx->compute, cudaStreamNonBlocking));
    if (!f.copy) {
        CU(cudaStreamCreateWithFlags(&f.copy, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&f.h2d, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&f.done, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&f.d2h, cudaEventDisableTiming));
    }
    const size_t npix = size_t(P->width) * P->height;
    CU(f.out.ensure(npix * 16));
    void* d_tex1 = nullptr;
    void* d_albedo = nullptr;
    if (h_tex1) {
        CU(f.tex1.ensure(npix * 16));
        CU(cudaMemcpyAsync(f.tex1.p, h_tex1, npix * 16, cudaMemcpyHostToDevice, f.copy));
        d_tex1 = f.tex1.p;
    }
    if (h_albedo) {
        CU(f.albedo.ensure(npix * 16));
        CU(cudaMemcpyAsync(f.albedo.p, h_albedo, npix * 16, cudaMemcpyHostToDevice, f.copy));
        d_albedo = f.albedo.p;
    }
    CU(cudaEventRecord(f.h2d, f.copy));
    CU(cudaStreamWaitEvent(ctx->compute, f.h2d, 0));
    const mpt_status s = mpt_render(ctx, camera, P, d_tex1, d_albedo, f.out.p, ctx->compute);
    if (s != MPT_OK) return s;
    CU(cudaEventRecord(f.done, ctx->compute));
    CU(cudaStreamWaitEvent(f.copy, f.done, 0));
    CU(cudaMemcpyAsync(h_irradiance, f.out.p, npix * 16, cudaMemcpyDeviceToHost, f.copy));
    CU(cudaEventRecord(f.d2h, f.copy));
    f.pending = true;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_render_host_begin: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_render_host_begin: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_render_host_begin: unknown exception");
}

MPT_API mpt_status mpt_render_host_end(mpt_context* ctx, uint32_t in_flight_index) try {
    if (!ctx) return MPT_ERR_INVALID_ARGUMENT;
    if (in_flight_index >= mpt_context::kMaxInFlight) return fail(ctx, MPT_ERR_LIMIT, "in_flight_index must be < 4");
    mpt_context::InFlight& f = ctx->in_flight[in_flight_index];
    if (!f.pending) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "no frame is in flight in this slot");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventSynchronize(f.d2h));
    f.pending = false; // only once the copies are known to be complete
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_render_host_end: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_render_host_end: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_render_host_end: unknown exception");
}

MPT_API mpt_status mpt_set_scene_frame(mpt_context* ctx, uint32_t frame) try {
    if (!ctx) return MPT_ERR_INVALID_ARGUMENT;
    ctx->scene_frame = frame;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_set_scene_frame: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_set_scene_frame: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_set_scene_frame: unknown exception");
}

MPT_API mpt_status mpt_scene_set_prev_transforms(mpt_context* ctx, const float* prev_instance_transforms, uint32_t instance_count) try {
    if (!ctx) return MPT_ERR_INVALID_ARGUMENT;
    if (!ctx->scene_ready) return fail(ctx, MPT_ERR_NOT_READY, "mpt_scene_set_prev_transforms: no scene uploaded");
    CU(quiesce(ctx)); // a frame in flight may be reading the old matrices
    if (!prev_instance_transforms) {
        ctx->view.inst_xform_prev = nullptr;
        return MPT_OK;
    }
    if (instance_count != ctx->instance_count) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_scene_set_prev_transforms: instance_count differs from the uploaded scene's");
    CU(ctx->xform_prev.upload(prev_instance_transforms, size_t(instance_count) * 64));
    ctx->view.inst_xform_prev = ctx->xform_prev.as<float4>();
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_scene_set_prev_transforms: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_scene_set_prev_transforms: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_scene_set_prev_transforms: unknown exception");
}

MPT_API mpt_status mpt_set_profiling(mpt_context* ctx, uint32_t profile_flags) try {
    if (!ctx) return MPT_ERR_INVALID_ARGUMENT;
    ctx->profile_flags = profile_flags;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_set_profiling: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_set_profiling: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_set_profiling: unknown exception");
}

MPT_API mpt_status mpt_get_stats(mpt_context* ctx, mpt_stats* out) try {
    if (!ctx || !out) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());
    harvest_events(ctx);
    *out = ctx->stats;
    if (ctx->counters.p) {
        Counters c;
        CU(cudaMemcpy(&c, ctx->counters.p, sizeof(Counters), cudaMemcpyDeviceToHost));
        out->rays_continuation += c.rays_continuation;
        out->rays_shadow += c.rays_shadow;
        out->shaded_hits += c.shaded_hits;
        out->nodes_visited += c.nodes_visited;
        out->tris_tested += c.tris_tested;
    }
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_get_stats: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_get_stats: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_get_stats: unknown exception");
}

MPT_API mpt_status mpt_reset_stats(mpt_context* ctx) try {
    if (!ctx) return MPT_ERR_INVALID_ARGUMENT;
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());
    harvest_events(ctx);
    ctx->stats = mpt_stats{};
    if (ctx->counters.p) CU(cudaMemset(ctx->counters.p, 0, sizeof(Counters)));
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_reset_stats: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_reset_stats: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_reset_stats: unknown exception");
}

MPT_API mpt_status mpt_debug_ray_queue(mpt_context* ctx, uint32_t bounce, mpt_ray* h_rays, uint32_t* h_leaves, uint64_t capacity, uint64_t* count) try {
    if (!ctx || !count) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    if (bounce < 1 || bounce > uint32_t(kMaxBounces) || !ctx->counters.p || !ctx->pool_slots) return fail(ctx, MPT_ERR_NOT_READY, "no frame has been rendered");
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());
    uint32_t n = 0;
    CU(cudaMemcpy(&n, &ctx->counters.as<Counters>()->qcount[bounce], sizeof(n), cudaMemcpyDeviceToHost));
    *count = n;
    if (!h_rays || capacity == 0) return MPT_OK;
    const size_t m = size_t(std::min<uint64_t>(n, capacity));
    try {
        std::vector<float4> o(m), d(m);
        const uint32_t parity = (bounce - 1u) & 1u;
        CU(cudaMemcpy(o.data(), ctx->q_o[parity].p, m * sizeof(float4), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(d.data(), ctx->q_d[parity].p, m * sizeof(float4), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < m; i++) {
            h_rays[i].origin[0] = o[i].x; h_rays[i].origin[1] = o[i].y; h_rays[i].origin[2] = o[i].z;
            h_rays[i].tmin = 0.f; // SCENE_RAY_TMIN / SCENE_RAY_TMAX, hit.slang:7-8
            h_rays[i].direction[0] = d[i].x; h_rays[i].direction[1] = d[i].y; h_rays[i].direction[2] = d[i].z;
            h_rays[i].tmax = kRayTMax;
            if (h_leaves) h_leaves[i] = fbits(d[i].w);
        }
    } catch (...) {
        return fail(ctx, MPT_ERR_LIMIT, "out of host memory");
    }
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_debug_ray_queue: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_debug_ray_queue: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_debug_ray_queue: unknown exception");
}

MPT_API mpt_status mpt_debug_ray_setup(mpt_context* ctx, const float* h_dirs, uint64_t n, float* h_shear, uint32_t* h_k_perm, float* h_idir) try {
    if (!ctx || !h_dirs || !h_shear || !h_k_perm || !h_idir || n == 0) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    CU(cudaSetDevice(ctx->device));
    DevBuf d, s, k, r;
    CU(d.ensure(n * 12));
    CU(s.ensure(n * 12));
    CU(k.ensure(n * 8));
    CU(r.ensure(n * 24));
    CU(cudaMemcpy(d.p, h_dirs, n * 12, cudaMemcpyHostToDevice));
    k_debug_ray_setup<<<uint32_t((n + 255) / 256), 256>>>(d.as<float>(), n, s.as<float>(), k.as<uint32_t>(), r.as<float>());
    CU(cudaGetLastError());
    CU(cudaMemcpy(h_shear, s.p, n * 12, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h_k_perm, k.p, n * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h_idir, r.p, n * 24, cudaMemcpyDeviceToHost));
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_debug_ray_setup: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_debug_ray_setup: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_debug_ray_setup: unknown exception");
}

MPT_API mpt_status mpt_bvh_get_info(mpt_context* ctx, mpt_bvh_info* out) try {
    if (!ctx || !out) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    if (!ctx->scene_ready) return fail(ctx, MPT_ERR_NOT_READY, "scene not uploaded");
    *out = ctx->bvh_info;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_bvh_get_info: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_bvh_get_info: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_bvh_get_info: unknown exception");
}

struct mpt_host_accel {
    HostAccel a;
};

MPT_API mpt_status mpt_host_accel_build(const mpt_scene_desc* scene, mpt_host_accel** out) try {
    if (!scene || !out) return fail(nullptr, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    *out = nullptr;
    mpt_host_accel* h = nullptr;
    try {
        h = new mpt_host_accel();
        std::string err;
        const mpt_status s = build_host_accel(scene, h->a, uint32_t(kStackSize), err);
        if (s != MPT_OK) { delete h; return fail(nullptr, s, err); }
    } catch (...) {
        delete h;
        return fail(nullptr, MPT_ERR_LIMIT, "out of host memory while building the acceleration structure");
    }
    *out = h;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_LIMIT, "mpt_host_accel_build: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_INVALID_ARGUMENT, std::string("mpt_host_accel_build: ") + e.what());
} catch (...) {
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_INVALID_ARGUMENT, "mpt_host_accel_build: unknown exception");
}

MPT_API mpt_status mpt_host_accel_get(const mpt_host_accel* h, mpt_host_accel_view* v) try {
    if (!h || !v) return fail(nullptr, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    const HostAccel& a = h->a;
    v->nodes = a.nodes.data(); v->node_count = a.nodes.size();
    v->tris = a.tris.data(); v->triangle_count = a.tris.size() / 3;
    v->instances = a.instances.data(); v->instance_count = a.instances.size();
    v->tlas_prims = a.tlas_prims.data(); v->tlas_prim_count = a.tlas_prims.size();
    v->prim_to_tri = a.prim_to_tri.data(); v->prim_count = a.prim_to_tri.size();
    v->root = a.root; v->two_level = a.two_level; v->max_blas_depth = a.max_blas_depth; v->tlas_depth = a.tlas_depth;
    v->build_ms = a.build_ms;
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_LIMIT, "mpt_host_accel_get: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_INVALID_ARGUMENT, std::string("mpt_host_accel_get: ") + e.what());
} catch (...) {
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_INVALID_ARGUMENT, "mpt_host_accel_get: unknown exception");
}

MPT_API void mpt_host_accel_free(mpt_host_accel* h) { delete h; }

MPT_API mpt_status mpt_bvh8_build_info(const float* boxes, uint32_t n, mpt_bvh_info* out, uint32_t* prim_order) try {
    mpt_context* ctx = nullptr;
    if ((!boxes && n) || !out) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    const auto t0 = std::chrono::steady_clock::now();
    Bvh8 bvh;
    build_bvh8(reinterpret_cast<const Box*>(boxes), n, bvh);
    *out = mpt_bvh_info{};
    out->node_count = bvh.nodes.size();
    out->triangle_count = bvh.prim_order.size();
    out->node_bytes = bvh.nodes.size() * sizeof(Bvh8Node);
    out->triangle_bytes = bvh.prim_order.size() * 48;
    out->max_depth = bvh.depth;
    out->build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (prim_order) std::memcpy(prim_order, bvh.prim_order.data(), bvh.prim_order.size() * 4);
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_LIMIT, "mpt_bvh8_build_info: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_INVALID_ARGUMENT, std::string("mpt_bvh8_build_info: ") + e.what());
} catch (...) {
    return fail(static_cast<mpt_context*>(nullptr), MPT_ERR_INVALID_ARGUMENT, "mpt_bvh8_build_info: unknown exception");
}

MPT_API mpt_status mpt_bsdf_check(mpt_context* ctx, const mpt_bsdf_check_params* p, float* out16) try {
    if (!ctx || !p || !out16) return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    if (p->bsdf == 6 || p->bsdf > 9)
        return fail(ctx, MPT_ERR_UNSUPPORTED, "not a sub-tree of GltfBSDF that can be isolated (SheenBRDF alone: use 9, SheenLayer over the dielectric)");
    CU(cudaSetDevice(ctx->device));
    DevBuf o;
    CU(o.ensure(64));
    if (p->bsdf == 3 && std::fabs(p->ior_f0 - 0.04f) > 1e-6f)
        return fail(ctx, MPT_ERR_UNSUPPORTED, "FresnelMix is reachable as the clearcoat layer only, whose f0 is 0.04 (GLTF_DIELECTRIC_IOR_F0, gltf-material.slang:29)");
    k_bsdf_check<0x3Fu><<<1, 1>>>(*p, o.as<float>());
    CU(cudaGetLastError());
    CU(cudaMemcpy(out16, o.p, 64, cudaMemcpyDeviceToHost));
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return fail(ctx, MPT_ERR_LIMIT, "mpt_bsdf_check: out of host memory");
} catch (const std::exception& e) { // no C++ exception crosses the C ABI
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, std::string("mpt_bsdf_check: ") + e.what());
} catch (...) {
    return fail(ctx, MPT_ERR_INVALID_ARGUMENT, "mpt_bsdf_check: unknown exception");
}

} // extern "C"
