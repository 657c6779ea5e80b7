// bvh8_traverse_pooled.cuh — the traversal of bvh8_traverse.cuh reorganised around a per-warp ray pool in shared memory.
//
// Why: in the first organisation (k_extend of round 1) a lane whose ray finished waited, idle, until fewer than 14 lanes of its warp
// were still traversing; only then the warp re-converged, retired the finished rays and fetched new ones. ncu (profiles/r01l_extend_*):
// the node test, 47 % of all issued instructions, ran with 22.8 of 32 lanes, retire + refill (15 %) with 15.6 lanes, and the ray setup
// (3 reciprocals, 3 IEEE divisions) and the environment lookup of a miss executed inside that half-empty region.
//
// Here every warp owns a pool of 64 ray records in shared memory. In a CONVERGED phase (all 32 lanes) the warp retires the finished
// records (hit-queue compaction, miss shading), claims as many new rays from the global queue with one atomicAdd, runs the ray setup
// at full width and files the records as READY. In the DIVERGENT traversal loop a lane that finishes writes its hit into its record,
// appends the record to the DONE list (one shared-memory atomic), pops a READY record (another one) and goes on traversing without
// any warp-level synchronisation. The warp only re-converges when the READY list is empty and fewer than `refill` lanes are busy.
//
// The traversal stack's first SS entries per lane live in shared memory (lane-interleaved: conflict-free 8-byte accesses); deeper
// entries spill to local memory. The shear constants of the watertight test are read from the ray record when a triangle is tested
// (single-level scenes) instead of occupying six registers through the node tests.
#pragma once

#include "bvh8_traverse.cuh"

namespace mpt {

// One pooled ray record: 64 bytes of shared memory (shared memory is carved out of the SM's L1, which the traversal lives on: the pool
// is kept as small as the measurements allow, see k_extend_pooled).
struct alignas(16) PoolEntry {
    float4 a; // origin.xyz | path slot (bits); slot 0xFFFFFFFF = the record holds no ray
    float4 b; // 1/direction.xyz | the triangle record the ray leaves (bits); after the ray finished: hit record tri | instance | u | v
    float4 c; // shear Sx, Sy, Sz | kx | ky << 2 | kz << 4 | octant << 8 (bits)
    float4 d; // direction.xyz | the queue record's w lane again (with NEE: the BSDF pdf)
};

// One node test: visits the child at `bit` of a node group (child_base, parent_imask), intersects its 8 child boxes and returns the new
// node group G and the primitives it uncovered (fresh; fresh.y == 0: none). Same arithmetic as Traversal::step section 1 (comments there).
// TMIN_ZERO: the ray's tmin is 0 (continuation rays), the lower clamp is the 0 already inside tn.
template <bool TMIN_ZERO>
MPT_DEV void bvh8_node_test(const AccelView& A, uint32_t child_base, uint32_t parent_imask, uint32_t bit, f3 o, f3 idir, uint32_t octinv4, float tlo, float thi,
                            uint2& G, uint2& fresh) {
    const uint32_t slot = (bit - 24u) ^ (octinv4 & 0xFFu);
    const uint32_t rel = __popc(parent_imask & ~(0xFFFFFFFFu << slot));
    const uint32_t ni = child_base + rel;
    const float4* np = A.nodes + size_t(ni) * 5;
    const float4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
    const uint32_t eimask = __float_as_uint(n0.w);
    const float aix = __uint_as_float((eimask & 0xFFu) << 23) * idir.x;
    const float aiy = __uint_as_float(((eimask >> 8) & 0xFFu) << 23) * idir.y;
    const float aiz = __uint_as_float(((eimask >> 16) & 0xFFu) << 23) * idir.z;
    const float ox = fmaf(-32768.0f, aix, (n0.x - o.x) * idir.x);
    const float oy = fmaf(-32768.0f, aiy, (n0.y - o.y) * idir.y);
    const float oz = fmaf(-32768.0f, aiz, (n0.z - o.z) * idir.z);
    const float ex = 0.00390625f * fabsf(aix), ey = 0.00390625f * fabsf(aiy), ez = 0.00390625f * fabsf(aiz);
    const float oxn = ox - ex, oxf = ox + ex, oyn = oy - ey, oyf = oy + ey, ozn = oz - ez, ozf = oz + ez;
    const uint32_t bias = A.prmt_bias;
    const uint32_t imask = eimask >> 24;
    uint32_t hitmask = 0;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const uint32_t meta4 = __float_as_uint(half ? n1.w : n1.z);
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
        const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1F1F1F1Fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
        const uint32_t qlox = __float_as_uint(half ? n2.y : n2.x), qloy = __float_as_uint(half ? n2.w : n2.z);
        const uint32_t qloz = __float_as_uint(half ? n3.y : n3.x), qhix = __float_as_uint(half ? n3.w : n3.z);
        const uint32_t qhiy = __float_as_uint(half ? n4.y : n4.x), qhiz = __float_as_uint(half ? n4.w : n4.z);
        const uint32_t nearx = idir.x < 0.f ? qhix : qlox, farx = idir.x < 0.f ? qlox : qhix;
        const uint32_t neary = idir.y < 0.f ? qhiy : qloy, fary = idir.y < 0.f ? qloy : qhiy;
        const uint32_t nearz = idir.z < 0.f ? qhiz : qloz, farz = idir.z < 0.f ? qloz : qhiz;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float t0x = fmaf(byte_biased(nearx, bias, j), aix, oxn), t1x = fmaf(byte_biased(farx, bias, j), aix, oxf);
            const float t0y = fmaf(byte_biased(neary, bias, j), aiy, oyn), t1y = fmaf(byte_biased(fary, bias, j), aiy, oyf);
            const float t0z = fmaf(byte_biased(nearz, bias, j), aiz, ozn), t1z = fmaf(byte_biased(farz, bias, j), aiz, ozf);
            const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.f)) * 0.9999995f;
            const float tf = fminf(fminf(t1x, t1y), t1z) * 1.0000005f;
            if ((TMIN_ZERO ? tn : fmaxf(tn, tlo)) <= fminf(tf, thi)) {
                const uint32_t cb = (child_bits4 >> (8 * j)) & 0xFFu;
                const uint32_t bi = (bit_index4 >> (8 * j)) & 0xFFu;
                hitmask |= cb << bi;
            }
        }
    }
    G.x = __float_as_uint(n1.x);
    G.y = (hitmask & 0xFF000000u) | imask;
    fresh = make_uint2(__float_as_uint(n1.y), hitmask & 0x00FFFFFFu);
}

// Resumable closest-hit traversal fed from a WarpPool. SS = stack entries kept in shared memory (per lane, stride STRIDE uint2).
template <bool COUNT, bool TWO_LEVEL, class AlphaTest, int SS, int STRIDE> struct TraversalP {
    const AccelView& A;
    const AlphaTest& alpha_test;
    uint2* sst;            // shared-memory part of the stack: entry i at sst[i * STRIDE]
    uint2* lst;            // local-memory overflow: entry i >= SS at lst[i - SS]
    const PoolEntry* ent;  // the record this lane is traversing
    f3 o, idir;            // current-space ray (object space inside an instance)
    uint32_t octinv4;
    Shear shear;           // TWO_LEVEL only (registers); single-level scenes read the record's shear when a triangle is tested
    float best_t;
    uint32_t ray_mask;
    int sp;
    uint2 G, Gt;
    bool in_blas, force_opaque;
    int cull;
    uint32_t cur_instance;
    HitRec hit;
    uint32_t n_nodes, n_tris;

    MPT_DEV TraversalP(const AccelView& a, const AlphaTest& at, uint2* s, uint2* l) : A(a), alpha_test(at), sst(s), lst(l), ent(nullptr), n_nodes(0), n_tris(0) {}
    MPT_DEV bool found() const { return hit.instance != kMiss; }

    MPT_DEV void push(uint2 v) {
        if (SS > 0 && sp < SS) sst[sp * STRIDE] = v;
        else lst[sp - SS] = v;
        sp++;
    }
    MPT_DEV uint2 top() const { return (SS > 0 && sp - 1 < SS) ? sst[(sp - 1) * STRIDE] : lst[sp - 1 - SS]; }

    // returns the record's b.w lane (the triangle the ray leaves)
    MPT_DEV uint32_t load_world_setup() {
        const float4 a = ent->a, b = ent->b, c = ent->c;
        o = mk3(a.x, a.y, a.z);
        idir = mk3(b.x, b.y, b.z);
        const uint32_t kk = __float_as_uint(c.w);
        octinv4 = (7u - ((kk >> 8) & 7u)) * 0x01010101u;
        if (TWO_LEVEL) {
            shear.Sx = c.x; shear.Sy = c.y; shear.Sz = c.z;
            shear.kx = int(kk & 3u); shear.ky = int((kk >> 2) & 3u); shear.kz = int((kk >> 4) & 3u);
        }
        return __float_as_uint(b.w);
    }

    // scene_flags: mask_flags of instance 0 (single-level scenes). use_skip: leave out the triangle the ray starts on (record lane b.w)
    MPT_DEV void start(const PoolEntry* e, uint32_t mask, uint32_t scene_flags, bool use_skip) {
        ent = e;
        const uint32_t leaves = load_world_setup();
        const uint32_t skip_tri = use_skip ? leaves : kMiss;
        best_t = kRayTMaxPooled();
        ray_mask = mask;
        hit.tri = TWO_LEVEL ? kMiss : skip_tri;
        hit.instance = kMiss;
        hit.u = hit.v = 0.f;
        sp = 0;
        in_blas = !TWO_LEVEL;
        cur_instance = 0;
        cull = 0;
        force_opaque = true;
        if (!TWO_LEVEL) {
            const uint32_t fl = scene_flags >> 8;
            cull = (fl & 1u) ? 2 : ((fl & 2u) ? 1 : 0);
            force_opaque = (fl & 4u) != 0;
        }
        G = make_uint2(A.root, 0x80000000u);
        Gt = make_uint2(0u, 0u);
    }
    static MPT_DEV float kRayTMaxPooled() { return 1e7f; } // SCENE_RAY_TMAX, hit.slang:8 (continuation rays: tmin 0)

    MPT_DEV bool step() {
        // ---------------------------------------------------------------- 1. one node test
        if (G.y & 0xFF000000u) {
            const uint32_t bit = 31u - uint32_t(__clz(G.y));
            uint2 rest = G;
            rest.y &= ~(1u << bit);
            if (rest.y & 0xFF000000u) push(rest);
            if (COUNT) n_nodes++;
            uint2 fresh;
            bvh8_node_test<true>(A, G.x, G.y & 0xFFu, bit, o, idir, octinv4, 0.f, best_t, G, fresh);
            if (fresh.y) {
                if (Gt.y) push(fresh); // one group is already pending: park the new one
                else Gt = fresh;
            }
        }
        // the stack top this lane will need in section 3, loaded before the primitive test (single-level scenes: section 2 never pushes)
        uint2 pre_top = make_uint2(0u, 0u);
        const bool pre_valid = !TWO_LEVEL && (G.y & 0xFF000000u) == 0u && sp > 0;
        if (pre_valid) pre_top = top();
        // ---------------------------------------------------------------- 2. one primitive of the pending group
        if (Gt.y) {
            do {
                const uint32_t i = 31u - uint32_t(__clz(Gt.y));
                Gt.y &= ~(1u << i);
                if (!TWO_LEVEL || in_blas) {
                    const uint32_t ti = Gt.x + i;
                    if (TWO_LEVEL || ti != hit.tri) {
                        if (COUNT) n_tris++;
                        const float4* tp = A.tris + size_t(ti) * 3;
                        const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                        Shear s;
                        if (TWO_LEVEL) {
                            s = shear;
                        } else {
                            const float4 sc = ent->c;
                            s.Sx = sc.x; s.Sy = sc.y; s.Sz = sc.z;
                            const uint32_t kk = __float_as_uint(sc.w);
                            s.kx = int(kk & 3u); s.ky = int((kk >> 2) & 3u); s.kz = int((kk >> 4) & 3u);
                        }
                        float t, u, v;
                        if (intersect_triangle(s, o, xyz(a), xyz(b), xyz(c), cull, 0.f, best_t, t, u, v)) {
                            if (force_opaque || !(__float_as_uint(c.w) & kTriNonOpaque) || !alpha_test(ti, cur_instance, u, v)) {
                                best_t = t;
                                hit.tri = ti;
                                hit.instance = cur_instance;
                                hit.u = u;
                                hit.v = v;
                            }
                        }
                    }
                } else {
                    // TLAS leaf: enter the instance
                    const uint32_t inst = __ldg(A.tlas_prims + Gt.x + i);
                    const InstanceRec* ip = A.instances + inst;
                    const uint4 meta = __ldg(reinterpret_cast<const uint4*>(ip) + 3);
                    if (((meta.z & 0xFFu) & ray_mask) == 0) continue;
                    const float4 r0 = __ldg(&ip->inv0), r1 = __ldg(&ip->inv1), r2 = __ldg(&ip->inv2);
                    if (G.y & 0xFF000000u) push(G);
                    if (Gt.y) push(Gt);
                    push(make_uint2(0u, 0u)); // sentinel: leave the instance
                    const float4 wa = ent->a, wd = ent->d;
                    const f3 org_w = mk3(wa.x, wa.y, wa.z), dir_w = mk3(wd.x, wd.y, wd.z);
                    const f3 oo = mk3(affine_rn(r0, org_w, r0.w), affine_rn(r1, org_w, r1.w), affine_rn(r2, org_w, r2.w));
                    const f3 od = mk3(linear_rn(r0, dir_w), linear_rn(r1, dir_w), linear_rn(r2, dir_w));
                    RaySetup R;
                    setup_ray(R, oo, od);
                    o = R.o; idir = R.idir; octinv4 = R.octinv4; shear = R.shear;
                    cur_instance = inst;
                    const uint32_t fl = meta.z >> 8;
                    cull = (fl & 1u) ? 2 : ((fl & 2u) ? 1 : 0);
                    force_opaque = (fl & 4u) != 0;
                    in_blas = true;
                    G = make_uint2(meta.x, 0x80000000u);
                    Gt = make_uint2(0u, 0u);
                }
            } while (Gt.y && sp > kStackSize - 6);
        }
        // ---------------------------------------------------------------- 3. refill the groups from the stack
        bool first_pop = true;
        while ((G.y & 0xFF000000u) == 0) {
            if (sp == 0) return Gt.y == 0;
            const uint2 t = (pre_valid && first_pop) ? pre_top : top();
            first_pop = false;
            if (t.y & 0xFF000000u) {
                G = t;
                --sp;
                break;
            }
            if (Gt.y) break;
            --sp;
            if (t.y) {
                Gt = t;
            } else if (TWO_LEVEL) { // sentinel: back to the TLAS with the world-space ray of the record
                load_world_setup();
                in_blas = false;
            }
        }
        return false;
    }
};

} // namespace mpt
