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

// Resumable closest-hit traversal fed from a per-warp ray pool. One step() = at most one node test and one primitive test, as in
// Traversal (bvh8_traverse.cuh), with two differences: node groups and parked primitive groups live on SEPARATE stacks — refilling the
// current groups is then a pair of predicated pops instead of a loop that has to tell the entry kinds apart (10 % of the round-1
// kernel's instructions, at 8.6 of 32 lanes) — and the shear constants of the watertight test are read from the ray record when a
// triangle is tested (single-level scenes) instead of occupying registers through the node tests.
// SS = node-stack entries kept in shared memory (per lane, stride STRIDE uint2); 0 = all in local memory (measured best, see k_extend_pooled).
#ifndef MPT_TLAS_BOX_RECHECK
#define MPT_TLAS_BOX_RECHECK 1 // instance_box_missed() before an instance is entered (bvh8_traverse.cuh)
#endif
#ifndef MPT_TLAS_GATE
#define MPT_TLAS_GATE 0 // see step(): measured, no gain with the trigger form of the pool
#endif
constexpr int kPrimStackSize = 24; // parked primitive groups: at most one per node test while another group is pending

// TMIN_ZERO = true: continuation rays (tmin 0, tmax SCENE_RAY_TMAX). false: rays with their own interval (mpt_trace, primary visibility):
// the record's b.w lane holds tmin and its d.w lane tmax instead of the origin triangle / the BSDF pdf.
template <bool COUNT, bool TWO_LEVEL, class AlphaTest, int SS, int STRIDE, bool TMIN_ZERO = true> struct TraversalP {
    const AccelView& A;
    const AlphaTest& alpha_test;
    uint2* sst;            // shared-memory part of the node stack: entry i at sst[i * STRIDE]
    uint2* lst;            // local-memory part: entry i >= SS at lst[i - SS]
    uint4* pst;            // parked primitive groups (local memory): first record | hits (slot space) + stack height << 24 | leaf_bits | -
    const PoolEntry* ent;  // the record this lane is traversing
    f3 o, idir;            // current-space ray (object space inside an instance)
    uint32_t perm;         // make_perm(octant) of the current-space ray
    Shear shear;           // TWO_LEVEL only (registers); single-level scenes read the record's shear when a triangle is tested
    float best_t;
    float tmin;            // !TMIN_ZERO only
    uint32_t ray_mask;
    int sp, psp, pbase;    // node stack pointer; primitive stack pointer and its floor inside the current instance
    uint32_t ptop_h;       // the node-stack height the TOP parked group was uncovered at (each parked entry's w lane keeps the value for
                           // the entry below it), so that section 3 needs no local-memory read to decide
    uint2 G;               // current node group: x = first inner child, y = hits << 24 | imask
    uint2 Gt;              // pending primitive group: x = first primitive record of its node, y = hit primitives (slot-space bits)
    uint32_t Gv;           //   ... and the node's leaf_bits: the primitive at bit b is record Gt.x + popc(Gv & ((1 << b) - 1)). Kept in a
                           //   register: re-reading it from the node when a primitive is tested put a second dependent load in front of
                           //   the triangle fetch (long-scoreboard stalls per issue 2.3 -> 3.9, extend +3.5 %, profiles/r02d_*)
    bool in_blas, force_opaque;
    int cull;
    uint32_t cur_instance;
    HitRec hit;
    uint32_t n_nodes, n_tris;

    MPT_DEV TraversalP(const AccelView& a, const AlphaTest& at, uint2* s, uint2* l, uint4* p)
        : A(a), alpha_test(at), sst(s), lst(l), pst(p), ent(nullptr), n_nodes(0), n_tris(0) {}
    MPT_DEV bool found() const { return hit.instance != kMiss; }

    MPT_DEV void push(uint2 v) {
        if (SS > 0 && sp < SS) sst[sp * STRIDE] = v;
        else lst[sp - SS] = v;
        sp++;
    }
    MPT_DEV void unpark() {
        const uint4 p = pst[--psp];
        Gt = make_uint2(p.x, p.y & 0x00FFFFFFu);
        Gv = p.z;
        ptop_h = p.w;
    }
    MPT_DEV uint2 top() const { return (SS > 0 && sp - 1 < SS) ? sst[(sp - 1) * STRIDE] : lst[sp - 1 - SS]; }

    // returns the record's b.w lane (the triangle the ray leaves)
    MPT_DEV uint32_t load_world_setup() {
        const float4 a = ent->a, b = ent->b, c = ent->c;
        o = mk3(a.x, a.y, a.z);
        idir = mk3(b.x, b.y, b.z);
        const uint32_t kk = __float_as_uint(c.w);
        perm = make_perm((kk >> 8) & 7u);
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
        if (TMIN_ZERO) {
            best_t = 1e7f; // SCENE_RAY_TMAX, hit.slang:8 (continuation rays: tmin 0)
        } else {
            tmin = __uint_as_float(leaves);
            best_t = e->d.w;
        }
        ray_mask = mask;
        hit.tri = (TWO_LEVEL || !use_skip || !TMIN_ZERO) ? kMiss : leaves;
        hit.instance = kMiss;
        hit.u = hit.v = 0.f;
        sp = 0;
        psp = 0;
        pbase = 0;
        ptop_h = 0u;
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
        Gv = 0u;
    }

    // tick: the caller's loop counter (the same in every lane of a warp), see MPT_TLAS_GATE
    MPT_DEV bool step(uint32_t tick = 0u) {
        // ---------------------------------------------------------------- 1. one node test
        if (G.y & 0xFF000000u) {
            const uint32_t bit = 31u - uint32_t(__clz(G.y));
            G.y &= ~(1u << bit);
            if (G.y & 0xFF000000u) push(G);
            const uint32_t slot = (bit - 24u) ^ (perm >> 24);
            const uint32_t ni = G.x + uint32_t(__popc(G.y & ~(0xFFFFFFFFu << slot)));
            if (COUNT) n_nodes++;
            uint32_t leaf;
            uint2 pl; // first primitive record, leaf_bits
            bvh8_node_test<TMIN_ZERO>(A, ni, o, idir, perm, TMIN_ZERO ? 0.f : tmin, best_t, G, leaf, &pl);
            if (leaf) {
                if (Gt.y) { // one group is already pending: park the new one
                    pst[psp++] = make_uint4(pl.x, leaf | (uint32_t(sp) << 24), pl.y, ptop_h);
                    ptop_h = uint32_t(sp);
                }
                else { Gt = make_uint2(pl.x, leaf); Gv = pl.y; }
            }
        }
        // the node-stack top this lane will need in section 3, loaded before the primitive test so that the local-memory latency hides
        // behind it (single-level scenes only: entering an instance in section 2 pushes)
        uint2 pre_top = make_uint2(0u, 0u);
        const bool pre_valid = !TWO_LEVEL && (G.y & 0xFF000000u) == 0u && sp > 0;
        if (pre_valid) pre_top = top();
        // ---------------------------------------------------------------- 2. one primitive of the pending group
        // (all of it when the primitive stack is nearly full, so that parking can never overflow)
        // MPT_TLAS_GATE = m > 0 (two-level scenes): a lane at the TLAS level enters an instance only in the loop iterations with
        // (tick & m) == 0 and goes on with the node tests of the TLAS in between: entering an instance is ~190 instructions that ran at
        // 5 of 32 lanes when every lane entered the moment it uncovered an instance (profiles/r02p_c5_extend_summary.txt); the counter
        // is the same in all lanes, so the entries of a warp fall into the same iterations.
        // Measured on C5 (profiles/r02v_ab.md): with the ramp-down pool m = 1: +0.9 %, m = 3: -0.9 % (a lane that has nothing else to do idles
        // until its turn); with the trigger pool m = 1: +-0 (1929 vs 1930 Msamples/s). Off by default.
        const bool gated = TWO_LEVEL && MPT_TLAS_GATE > 0 && !in_blas && (tick & uint32_t(MPT_TLAS_GATE)) != 0u && psp <= kPrimStackSize - 4;
        if (Gt.y && !gated) {
            do {
                const uint32_t i = 31u - uint32_t(__clz(Gt.y));
                Gt.y &= ~(1u << i);
                const uint32_t rec = Gt.x + uint32_t(__popc(Gv & ((1u << i) - 1u)));
                if (!TWO_LEVEL || in_blas) {
                    // single-level scenes: not the triangle the ray leaves (nor, trivially, the current closest hit). Two-level scenes test
                    // everything: the same record index occurs in every instance of a BLAS, so equality with hit.tri says nothing there
                    if (TWO_LEVEL || rec != hit.tri) {
                        if (COUNT) n_tris++;
                        const float4* tp = A.tris + size_t(rec) * 3;
                        const float4 a = ld_tri(tp), b = ld_tri(tp + 1), c = ld_tri(tp + 2);
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
                        if (intersect_triangle(s, o, xyz(a), xyz(b), xyz(c), cull, TMIN_ZERO ? 0.f : tmin, best_t, t, u, v)) {
                            if (force_opaque || !(__float_as_uint(c.w) & kTriNonOpaque) || !alpha_test(rec, cur_instance, u, v)) {
                                best_t = t;
                                hit.tri = rec;
                                hit.instance = cur_instance;
                                hit.u = u;
                                hit.v = v;
                            }
                        }
                    }
                } else {
                    // TLAS leaf: enter the instance
#if MPT_TLAS_BOX_RECHECK
                    if (instance_box_missed(__ldg(A.tlas_boxes + 2 * rec), __ldg(A.tlas_boxes + 2 * rec + 1), o, idir, best_t)) continue; // (o, idir): the world-space ray at the TLAS level
#endif
                    const uint32_t inst = __ldg(A.tlas_prims + rec);
                    const InstanceRec* ip = A.instances + inst;
                    const uint4 meta = __ldg(reinterpret_cast<const uint4*>(ip) + 3);
                    if (((meta.z & 0xFFu) & ray_mask) == 0) continue;
                    const float4 r0 = __ldg(&ip->inv0), r1 = __ldg(&ip->inv1), r2 = __ldg(&ip->inv2);
                    if (G.y & 0xFF000000u) push(G);
                    if (Gt.y) {
                        pst[psp++] = make_uint4(Gt.x, Gt.y | (uint32_t(sp) << 24), Gv, ptop_h);
                        ptop_h = uint32_t(sp);
                    }
                    push(make_uint2(uint32_t(pbase), 0u)); // sentinel: leave the instance (x = the TLAS level's primitive-stack floor)
                    pbase = psp;
                    const float4 wa = ent->a, wd = ent->d;
                    const f3 org_w = mk3(wa.x, wa.y, wa.z), dir_w = mk3(wd.x, wd.y, wd.z);
                    // world -> object, every product and sum rounded separately (no FMA contraction) so that the object-space
                    // ray, and with it the barycentrics, are bit-identical to a plain IEEE evaluation on the host
                    const f3 oo = mk3(affine_rn(r0, org_w, r0.w), affine_rn(r1, org_w, r1.w), affine_rn(r2, org_w, r2.w));
                    const f3 od = mk3(linear_rn(r0, dir_w), linear_rn(r1, dir_w), linear_rn(r2, dir_w));
                    RaySetup R;
                    setup_ray<true>(R, oo, od);
                    o = R.o; idir = R.idir; perm = R.perm; shear = R.shear;
                    cur_instance = inst;
                    const uint32_t fl = meta.z >> 8;
                    cull = (fl & 1u) ? 2 : ((fl & 2u) ? 1 : 0);
                    force_opaque = (fl & 4u) != 0;
                    in_blas = true;
                    G = make_uint2(meta.x, 0x80000000u);
                    Gt = make_uint2(0u, 0u);
                }
            } while (Gt.y && psp > kPrimStackSize - 4);
        }
        // ---------------------------------------------------------------- 3. refill the groups from the stacks
        // A parked primitive group remembers the node-stack height it was uncovered at (top byte of y): node groups below that height
        // wait until it has been tested — the order of the single-stack traversal (a group uncovered deeper in the tree is tested before
        // the walk moves on to the siblings of its ancestors), which keeps the hit distance fresh: taking sibling groups first cost
        // +1.4 % node tests and +4.6 % triangle tests on C3 (profiles/r02d_ab.jsonl). The stricter rule "never go back to a stacked node
        // group while ANY primitive is pending" brings the counts down to the oracle's immediate-test traversal (11.03 nodes, 3.57
        // triangles per ray) but costs 2-4 % time on every scene tried (C1, C3, C5: profiles/r02h_drain_first.jsonl): lanes idle through
        // node sections while they drain.
#ifndef MPT_POP_LOOP
        // Written out flat (predicated moves instead of a loop with a nest of divergent branches, which was ~45 instructions at 5-9 of 32
        // lanes on C3 and 18 % of the instructions at 2-15 lanes on C5: profiles/r02d_extend_summary.txt, r02p_c5_extend_summary.txt).
        if (!TWO_LEVEL) {
            // Single-level scenes: the node stack holds node groups only and section 2 never pushes: the stack top is pre_top.
            if (Gt.y == 0u && psp > 0) unpark();
            if ((G.y & 0xFF000000u) == 0u) {
                if (sp == 0) return Gt.y == 0u; // nothing left to visit: finished once the pending primitives are done
                if (!(psp > 0 && ptop_h >= uint32_t(sp))) {
                    G = pre_top;
                    --sp;
                }
            }
            return false;
        } else {
            // Two-level scenes: the top is a node group, or the sentinel that ends the instance (then the next entry, if any, is a node
            // group of the TLAS: instances do not nest). The same decisions as the loop below, at most one instance exit per step.
            if (Gt.y == 0u && psp > pbase) unpark();
            if ((G.y & 0xFF000000u) != 0u) return false;
            if (sp == 0) return Gt.y == 0u;
            if (psp > pbase && ptop_h >= uint32_t(sp)) return false;
            uint2 t = top();
            if ((t.y & 0xFF000000u) == 0u) { // the sentinel
                if (Gt.y) return false;      // the instance's pending primitives go first
                --sp;                        // back to the TLAS with the world-space ray of the record
                pbase = int(t.x);
                load_world_setup();
                in_blas = false;
                if (psp > pbase) unpark();
                if (sp == 0) return Gt.y == 0u;
                if (psp > pbase && ptop_h >= uint32_t(sp)) return false;
                t = top();
            }
            G = t;
            --sp;
            return false;
        }
#endif
        if (Gt.y == 0u && psp > pbase) unpark();
        bool first_pop = true;
        while ((G.y & 0xFF000000u) == 0u) {
            if (sp == 0) return Gt.y == 0u; // nothing left to visit: finished once the pending primitives are done
            if (psp > pbase && (pst[psp - 1].y >> 24) >= uint32_t(sp)) break;
            const uint2 t = (pre_valid && first_pop) ? pre_top : top();
            first_pop = false;
            if (t.y & 0xFF000000u) { // a node group
                G = t;
                --sp;
                break;
            }
            if (!TWO_LEVEL || Gt.y) break; // the sentinel: the instance's pending primitives go first
            --sp;                          // back to the TLAS with the world-space ray of the record
            pbase = int(t.x);
            load_world_setup();
            in_blas = false;
            if (psp > pbase) unpark();
        }
        return false;
    }
};

} // namespace mpt
