// oracle_bvh8.hpp — TEST INFRASTRUCTURE ONLY (included by oracle.cpp after `struct Scene`).
//
// CPU walker over the PRODUCT's acceleration structure: the compressed 8-wide BVH, triangle records and instance records
// exactly as libmpt.so lays them out in HBM (exported host-side by mpt_host_accel_build, include/mpt.h). SURVEY.md §8(d)
// defines the roofline numerator through it: "n̄_node, n̄_tri = mean nodes popped / triangles tested per ray, counted by the
// instrumented CPU oracle traversing the same BVH8 with the same front-to-back order on the same rays".
//
// It is deliberately the plain textbook traversal, not a transcription of the device code: one node at a time, children in the
// octant order the node layout encodes (slot ^ octant, nearest first), every triangle of a hit leaf child tested IMMEDIATELY
// (the device defers them, so its hit distance lags and it may visit more), exact double-precision slab tests on the decoded
// quantised boxes (the device widens them by a guard band). Hence device counts >= these counts; tests bound the excess.
// Hits use the same watertight triangle test as the BVH2 oracle (intersect_triangle above), so ids and barycentrics of the two
// oracle traversals and of the device must agree bit for bit.
//
// Node layout restated from merian_b200/csrc/bvh8.h (Ylitie, Karras, Laine 2017); semantics of the ray query from
// include/merian-shaders/scene/scene.slang:62-76, 500-519 (back-face culling, instance mask, alpha test of non-opaque candidates).
#pragma once

namespace orc {

struct Bvh8NodeRaw { // 80 bytes
    float p[3];
    uint8_t e[3];
    uint8_t imask;
    uint32_t child_base, prim_base;
    uint32_t leaf_bits; // 3 bits per slot: unary primitive count of a leaf child (1, 3, 7); 0 for inner / empty slots
    uint32_t reserved;
    uint8_t qlo[3][8]; // qlox, qloy, qloz
    uint8_t qhi[3][8]; // qhix, qhiy, qhiz
};
static_assert(sizeof(Bvh8NodeRaw) == 80, "BVH8 node is 80 bytes");

struct Bvh8InstanceRaw { // 64 bytes
    float inv[12];
    uint32_t blas_root, custom_index, mask_flags, pad;
};
static_assert(sizeof(Bvh8InstanceRaw) == 64, "instance record is 64 bytes");

struct Bvh8Counts {
    uint64_t rays = 0, nodes = 0, tris = 0, instance_entries = 0, max_stack = 0;
};

struct Bvh8Walker {
    const mpt_host_accel_view& A;
    const Scene* scene; // for the alpha test of non-opaque candidates; may be null when every geometry is opaque
    const Bvh8NodeRaw* nodes;
    const float* tris; // 12 floats per record
    const Bvh8InstanceRaw* inst;

    Bvh8Walker(const mpt_host_accel_view& a, const Scene* s)
        : A(a), scene(s), nodes((const Bvh8NodeRaw*)a.nodes), tris((const float*)a.tris), inst((const Bvh8InstanceRaw*)a.instances) {}

    static uint32_t fbits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

    struct State {
        Hit best;
        float best_t;
        uint32_t best_record = 0xFFFFFFFFu;
        Bvh8Counts cnt;
    };

    // children of `ni` hit by the ray within [max(tmin, 0), best_t], as an 8-bit slot mask
    uint32_t test_node(const Bvh8NodeRaw& n, const double o[3], const double idir[3], double tlo, double thi) const {
        uint32_t hits = 0;
        for (int s = 0; s < 8; s++) {
            if (!((n.imask >> s) & 1u) && !((n.leaf_bits >> (3 * s)) & 7u)) continue; // empty slot
            double tn = tlo, tf = thi;
            for (int a = 0; a < 3; a++) {
                const double sc = std::ldexp(1.0, int(n.e[a]) - 127);
                const double lo = double(n.p[a]) + double(n.qlo[a][s]) * sc, hi = double(n.p[a]) + double(n.qhi[a][s]) * sc;
                double t0 = (lo - o[a]) * idir[a], t1 = (hi - o[a]) * idir[a];
                if (t0 > t1) std::swap(t0, t1);
                tn = std::max(tn, t0);
                tf = std::min(tf, t1);
            }
            if (tn <= tf) hits |= 1u << s;
        }
        return hits;
    }

    // one BLAS (or the TLAS when `tlas` is set) with the given object-space ray
    void walk(uint32_t root, bool tlas, float3 o, float3 d, float tmin, uint32_t mask, uint32_t instance_index, uint32_t skip_record,
              State& st, int depth_base) const {
        auto safe = [](float v) { return std::fabs(v) > 1e-25f ? double(v) : double(std::copysign(1e-25f, v)); };
        const double od[3] = {o.x, o.y, o.z}, idir[3] = {1.0 / safe(d.x), 1.0 / safe(d.y), 1.0 / safe(d.z)};
        const uint32_t oct = (d.x < 0.f ? 1u : 0u) | (d.y < 0.f ? 2u : 0u) | (d.z < 0.f ? 4u : 0u);
        const uint32_t octinv = 7u - oct;
        const RayShear rs(d);
        int cull = CULL_BACK;
        bool force_opaque = true;
        uint32_t custom_index = 0;
        if (!tlas) {
            const Bvh8InstanceRaw& in = inst[instance_index];
            const uint32_t fl = in.mask_flags >> 8;
            cull = (fl & 1u) ? CULL_NONE : ((fl & 2u) ? CULL_FRONT : CULL_BACK);
            force_opaque = (fl & 4u) != 0;
            custom_index = in.custom_index;
        }
        uint32_t stack[128];
        int sp = 0;
        stack[sp++] = root;
        while (sp) {
            st.cnt.max_stack = std::max<uint64_t>(st.cnt.max_stack, uint64_t(sp + depth_base));
            const Bvh8NodeRaw& n = nodes[stack[--sp]];
            st.cnt.nodes++;
            const uint32_t hits = test_node(n, od, idir, std::max(double(tmin), 0.0), double(st.best_t));
            // leaf children first (their triangles can only shorten the ray), every triangle immediately
            for (int r = 7; r >= 0; r--) {
                const uint32_t s = uint32_t(r) ^ octinv;
                if (!(hits & (1u << s)) || (n.imask & (1u << s))) continue;
                const uint32_t unary = (n.leaf_bits >> (3 * s)) & 7u;
                const uint32_t off = uint32_t(__builtin_popcount(n.leaf_bits & ((1u << (3 * s)) - 1u))); // records of the leaf slots below s
                const uint32_t count = unary == 7u ? 3u : (unary == 3u ? 2u : 1u);
                for (uint32_t k = 0; k < count; k++) {
                    const uint32_t rec = n.prim_base + off + k;
                    if (tlas) {
                        const uint32_t ii = A.tlas_prims[rec];
                        const Bvh8InstanceRaw& in = inst[ii];
                        if (((in.mask_flags & 0xFFu) & mask) == 0) continue;
                        st.cnt.instance_entries++;
                        const float* m = in.inv; // world -> object; t is preserved (the direction is not renormalised)
                        const float3 oo(m[0] * o.x + m[1] * o.y + m[2] * o.z + m[3], m[4] * o.x + m[5] * o.y + m[6] * o.z + m[7],
                                        m[8] * o.x + m[9] * o.y + m[10] * o.z + m[11]);
                        const float3 dd(m[0] * d.x + m[1] * d.y + m[2] * d.z, m[4] * d.x + m[5] * d.y + m[6] * d.z, m[8] * d.x + m[9] * d.y + m[10] * d.z);
                        walk(in.blas_root, false, oo, dd, tmin, mask, ii, 0xFFFFFFFFu, st, depth_base + sp + 1);
                        continue;
                    }
                    if (rec == skip_record) continue;
                    st.cnt.tris++;
                    const float* tp = tris + size_t(rec) * 12;
                    float t, u, v;
                    if (!intersect_triangle(rs, o, float3(tp[0], tp[1], tp[2]), float3(tp[4], tp[5], tp[6]), float3(tp[8], tp[9], tp[10]), cull, tmin,
                                            st.best_t, t, u, v))
                        continue;
                    const uint32_t prim = fbits(tp[3]), geom = fbits(tp[7]), flags = fbits(tp[11]);
                    if (!force_opaque && (flags & 1u)) { // CANDIDATE_NON_OPAQUE_TRIANGLE: alpha test, scene.slang:503-516
                        if (scene) {
                            const mpt_geometry_desc& gd = scene->geometries[custom_index + geom];
                            const Mesh& m = scene->meshes[gd.mesh];
                            uint32_t ia, ib, ic;
                            m.tri(prim, ia, ib, ic);
                            const float3 b3(1.0f - u - v, u, v);
                            const float2 uv = b3.x * scene->vertex_uv(m.vertices[ia]) + b3.y * scene->vertex_uv(m.vertices[ib]) + b3.z * scene->vertex_uv(m.vertices[ic]);
                            if (scene->alpha_test(gd.material_id, uv)) continue;
                        }
                    }
                    st.best_t = t;
                    st.best.instance_id = custom_index;
                    st.best.instance_index = instance_index;
                    st.best.primitive_id = prim;
                    st.best.geometry_index = geom;
                    st.best.bary = float2(u, v);
                    st.best_record = rec;
                }
            }
            // inner children: push farthest first so that the nearest (highest slot ^ octinv) is popped next
            for (int r = 0; r < 8; r++) {
                const uint32_t s = uint32_t(r) ^ octinv;
                if (!(hits & (1u << s)) || !(n.imask & (1u << s))) continue;
                const uint32_t rel = uint32_t(__builtin_popcount(n.imask & ((1u << s) - 1u)));
                if (sp < 128) stack[sp++] = n.child_base + rel;
            }
        }
    }

    Hit trace(const Ray& ray, uint32_t mask, uint32_t skip_record, float* t_out, Bvh8Counts& cnt) const {
        State st;
        st.best_t = ray.tmax;
        if (A.two_level) {
            walk(A.root, true, ray.o, ray.d, ray.tmin, mask, 0, 0xFFFFFFFFu, st, 0);
        } else if ((inst[0].mask_flags & 0xFFu) & mask) {
            walk(A.root, false, ray.o, ray.d, ray.tmin, mask, 0, skip_record, st, 0);
        }
        st.cnt.rays = 1;
        cnt = st.cnt;
        if (t_out) *t_out = st.best.valid() ? st.best_t : -1.f;
        return st.best;
    }
};

} // namespace orc
