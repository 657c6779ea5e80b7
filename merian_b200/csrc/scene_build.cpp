// scene_build.cpp — see scene_build.h. Host-only; OpenMP over primitives / instances where the work is independent.
#include "scene_build.h"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>

namespace mpt {
namespace {

inline float bitsf(uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

void invert_affine(const float* m, float* o) { // row-major 3x4
    const double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    const double id = 1.0 / det;
    const double r[9] = {(e * i - f * h) * id, (c * h - b * i) * id, (b * f - c * e) * id, (f * g - d * i) * id, (a * i - c * g) * id,
                         (c * d - a * f) * id, (d * h - e * g) * id, (b * g - a * h) * id, (a * e - b * d) * id};
    for (int row = 0; row < 3; row++) {
        for (int col = 0; col < 3; col++) o[row * 4 + col] = float(r[row * 3 + col]);
        o[row * 4 + 3] = float(-(r[row * 3] * m[3] + r[row * 3 + 1] * m[7] + r[row * 3 + 2] * m[11]));
    }
}

inline void fetch_indices(const mpt_mesh_desc& m, uint32_t p, uint32_t ix[3]) {
    if (m.index_flag == MPT_GEOM_INDEX_U32) { const uint32_t* q = (const uint32_t*)m.indices + size_t(p) * 3; ix[0] = q[0]; ix[1] = q[1]; ix[2] = q[2]; }
    else if (m.index_flag == MPT_GEOM_INDEX_U16) { const uint16_t* q = (const uint16_t*)m.indices + size_t(p) * 3; ix[0] = q[0]; ix[1] = q[1]; ix[2] = q[2]; }
    else if (m.index_flag == MPT_GEOM_INDEX_U8) { const uint8_t* q = (const uint8_t*)m.indices + size_t(p) * 3; ix[0] = q[0]; ix[1] = q[1]; ix[2] = q[2]; }
    else { ix[0] = p * 3; ix[1] = ix[0] + 1; ix[2] = ix[0] + 2; }
}

} // namespace

// Cost of one instance in a TLAS leaf slot relative to a node test, for the collapse's dynamic program. With the triangles' 0.5 a quarter of
// C5's instances shared a slot with a neighbour, and a ray that hits the slot's box enters both: 1 / 2 / 4 -> +2.9 % on C5 (1929 -> 1984
// Msamples/s, profiles/r02z_c5_tlas.jsonl), every instance in a slot of its own. MPT_BVH_CPRIM_TLAS overrides (tuning sessions).
constexpr float kTlasPrimCost = 2.0f;

mpt_status build_host_accel(const mpt_scene_desc* d, HostAccel& out, uint32_t stack_entries, std::string& err) {
    if (d->instance_count == 0 || d->blas_count == 0) { err = "scene has no instances"; return MPT_ERR_INVALID_ARGUMENT; }
    if (d->instance_count > 0xFFFFu) { err = "more than 65535 TLAS instances (PackedHitInfo.instance_index is 16 bit, gbuffer.slang:44)"; return MPT_ERR_LIMIT; }
    const auto t_start = std::chrono::steady_clock::now();
    out = HostAccel();

    std::vector<uint32_t> blas_root(d->blas_count);
    std::vector<Box> blas_bounds(d->blas_count);
    std::vector<std::vector<uint32_t>> blas_geom_offset(d->blas_count);

    for (uint32_t bi = 0; bi < d->blas_count; bi++) {
        const mpt_blas_desc& bd = d->blas[bi];
        if (bd.mesh_count > 0xFFFFu) { err = "more than 65535 geometries in one BLAS (PackedHitInfo.geometry_index is 16 bit)"; return MPT_ERR_LIMIT; }
        struct Ref { uint32_t geom, prim; };
        blas_geom_offset[bi].resize(bd.mesh_count);
        size_t total = 0;
        std::vector<size_t> first(bd.mesh_count);
        for (uint32_t g = 0; g < bd.mesh_count; g++) {
            if (bd.meshes[g] >= d->mesh_count) { err = "BLAS references a mesh out of range"; return MPT_ERR_INVALID_ARGUMENT; }
            first[g] = total;
            blas_geom_offset[bi][g] = uint32_t(out.prim_to_tri.size() + total);
            total += d->meshes[bd.meshes[g]].primitive_count;
        }
        if (total > 0xFFFFFFF0ull || out.tris.size() / 3 + total > 0xFFFFFFF0ull) { err = "more than 2^32 triangle records"; return MPT_ERR_LIMIT; }
        RawVector<Ref> refs(total);
        RawVector<Box> boxes(total);
        bool bad_index = false;
        for (uint32_t g = 0; g < bd.mesh_count; g++) {
            const mpt_mesh_desc& m = d->meshes[bd.meshes[g]];
            const size_t f = first[g];
#pragma omp parallel for schedule(static) reduction(|| : bad_index) if (m.primitive_count > 20000)
            for (int64_t p = 0; p < int64_t(m.primitive_count); p++) {
                uint32_t ix[3];
                fetch_indices(m, uint32_t(p), ix);
                if (ix[0] >= m.vertex_count || ix[1] >= m.vertex_count || ix[2] >= m.vertex_count) { bad_index = true; continue; }
                Box bx;
                for (int a = 0; a < 3; a++) {
                    const float x0 = m.vertices[ix[0]].position[a], x1 = m.vertices[ix[1]].position[a], x2 = m.vertices[ix[2]].position[a];
                    bx.lo[a] = std::min(x0, std::min(x1, x2));
                    bx.hi[a] = std::max(x0, std::max(x1, x2));
                }
                boxes[f + size_t(p)] = bx;
                refs[f + size_t(p)] = {g, uint32_t(p)};
            }
        }
        if (bad_index) { err = "vertex index out of range"; return MPT_ERR_INVALID_ARGUMENT; }
        Bvh8 bvh;
        build_bvh8(boxes.data(), uint32_t(boxes.size()), bvh);
        const uint32_t node_off = uint32_t(out.nodes.size()), tri_off = uint32_t(out.tris.size() / 3);
        out.max_blas_depth = std::max(out.max_blas_depth, bvh.depth);
        blas_root[bi] = node_off;
        blas_bounds[bi] = bvh.bounds;
        out.nodes.reserve(out.nodes.size() + bvh.nodes.size());
        for (Bvh8Node nd : bvh.nodes) {
            nd.child_base += node_off;
            nd.prim_base += tri_off;
            out.nodes.push_back(nd);
        }
        const size_t p2t_base = out.prim_to_tri.size();
        out.prim_to_tri.resize(p2t_base + refs.size());
        const size_t nrec = bvh.prim_order.size();
        out.tris.resize((size_t(tri_off) + nrec) * 3);
        out.attrs.resize((size_t(tri_off) + nrec) * 3);
#pragma omp parallel for schedule(static) if (nrec > 20000)
        for (int64_t k = 0; k < int64_t(nrec); k++) {
            const uint32_t ri = bvh.prim_order[size_t(k)];
            const Ref r = refs[ri];
            const mpt_mesh_desc& m = d->meshes[bd.meshes[r.geom]];
            uint32_t ix[3];
            fetch_indices(m, r.prim, ix);
            const uint32_t w[3] = {r.prim, r.geom, m.opaque ? 0u : kTriNonOpaqueHost};
            uint32_t aw[12] = {0};
            HostVec4* tp = &out.tris[(size_t(tri_off) + size_t(k)) * 3];
            for (int c = 0; c < 3; c++) {
                const mpt_packed_vertex& v = m.vertices[ix[c]];
                tp[c] = HostVec4{v.position[0], v.position[1], v.position[2], bitsf(w[c])};
                aw[c * 3 + 0] = v.encoded_normal;
                aw[c * 3 + 1] = uint32_t(v.uv[0]) | (uint32_t(v.uv[1]) << 16);
                aw[c * 3 + 2] = v.encoded_tangent;
            }
            HostUVec4* ap = &out.attrs[(size_t(tri_off) + size_t(k)) * 3];
            ap[0] = HostUVec4{aw[0], aw[1], aw[2], aw[3]};
            ap[1] = HostUVec4{aw[4], aw[5], aw[6], aw[7]};
            ap[2] = HostUVec4{aw[8], 0, 0, 0};
            out.prim_to_tri[p2t_base + ri] = tri_off + uint32_t(k);
        }
    }

    // ---- instances + TLAS
    // vertex transforms needed for tight instance boxes; beyond the budget (about a second of host time) fall back to box corners
    uint64_t instance_vertex_work = 0;
    for (uint32_t i = 0; i < d->instance_count; i++) {
        if (d->instances[i].blas >= d->blas_count) { err = "instance references a BLAS out of range"; return MPT_ERR_INVALID_ARGUMENT; }
        const mpt_blas_desc& bd = d->blas[d->instances[i].blas];
        for (uint32_t g = 0; g < bd.mesh_count; g++) instance_vertex_work += d->meshes[bd.meshes[g]].vertex_count;
    }
    const bool tight_instance_boxes = instance_vertex_work <= 300000000ull && !std::getenv("MPT_LOOSE_INSTANCE_BOXES");
    out.instances.resize(d->instance_count);
    std::vector<Box> ibox(d->instance_count);
    bool identity_single = d->instance_count == 1;
    static const float ident12[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    for (uint32_t i = 0; i < d->instance_count; i++)
        if (std::memcmp(d->instances[i].transform, ident12, sizeof(ident12)) != 0) identity_single = false;
#pragma omp parallel for schedule(dynamic, 4) if (d->instance_count > 8)
    for (int64_t ii = 0; ii < int64_t(d->instance_count); ii++) {
        const uint32_t i = uint32_t(ii);
        const mpt_instance_desc& in = d->instances[i];
        HostInstanceRec& rec = out.instances[i];
        invert_affine(in.transform, rec.inv);
        rec.blas_root = blas_root[in.blas];
        rec.custom_index = in.instance_custom_index;
        rec.mask_flags = (in.mask & 0xFFu) | ((in.flags & 0xFFu) << 8);
        rec.pad = 0;
        const Box& ob = blas_bounds[in.blas];
        Box wb;
        for (int a = 0; a < 3; a++) { wb.lo[a] = 3e38f; wb.hi[a] = -3e38f; }
        auto grow = [&](const float* p) {
            for (int a = 0; a < 3; a++) {
                const float* m = in.transform + a * 4;
                const float v = m[0] * p[0] + m[1] * p[1] + m[2] * p[2] + m[3];
                wb.lo[a] = std::min(wb.lo[a], v);
                wb.hi[a] = std::max(wb.hi[a], v);
            }
        };
        if (tight_instance_boxes) {
            // the world box of the transformed VERTICES: for a rotated instance it is about half the surface area of the box around the
            // transformed object-space box, and every false TLAS hit costs an instance entry (ray transform + setup) in k_extend
            const mpt_blas_desc& bd = d->blas[in.blas];
            for (uint32_t g = 0; g < bd.mesh_count; g++) {
                const mpt_mesh_desc& m = d->meshes[bd.meshes[g]];
                for (uint32_t v = 0; v < m.vertex_count; v++) grow(m.vertices[v].position);
            }
        } else {
            for (int k = 0; k < 8; k++) {
                const float p[3] = {(k & 1) ? ob.hi[0] : ob.lo[0], (k & 2) ? ob.hi[1] : ob.lo[1], (k & 4) ? ob.hi[2] : ob.lo[2]};
                grow(p);
            }
        }
        for (int a = 0; a < 3; a++) { // the object-space test is exact; the world box only prunes
            const float pad = (wb.hi[a] - wb.lo[a]) * 1e-5f + 1e-6f;
            wb.lo[a] -= pad;
            wb.hi[a] += pad;
        }
        ibox[i] = wb;
    }
    out.root = blas_root[d->instances[0].blas];
    if (!identity_single) {
        Bvh8 tl;
        // an instance in a leaf slot is entered whenever the slot's box is hit (ray transform + ray setup + at least one BLAS node test,
        // ~450 instructions in k_extend_pooled against ~150 for a triangle): the collapse prices it accordingly, which keeps instances in
        // slots of their own instead of three to a slot
        const char* ev = std::getenv("MPT_BVH_CPRIM_TLAS");
        build_bvh8(ibox.data(), d->instance_count, tl, 3, ev ? float(std::atof(ev)) : kTlasPrimCost);
        const uint32_t node_off = uint32_t(out.nodes.size());
        out.root = node_off;
        for (Bvh8Node nd : tl.nodes) {
            nd.child_base += node_off;
            out.nodes.push_back(nd); // prim_base indexes tlas_prims directly
        }
        out.tlas_prims = tl.prim_order;
        // the exact world box of every TLAS leaf record's instance, in record order: a leaf slot's quantised box and the hit distance at
        // the time of the node test only say "maybe" (instance_box_missed, bvh8_traverse.cuh)
        out.tlas_boxes.resize(tl.prim_order.size() * 2);
        for (size_t k = 0; k < tl.prim_order.size(); k++) {
            const Box& b = ibox[tl.prim_order[k]];
            out.tlas_boxes[2 * k] = HostVec4{b.lo[0], b.lo[1], b.lo[2], 0.f};
            out.tlas_boxes[2 * k + 1] = HostVec4{b.hi[0], b.hi[1], b.hi[2], 0.f};
        }
        out.tlas_depth = tl.depth;
    }
    out.two_level = identity_single ? 0u : 1u;
    if (2u * (out.max_blas_depth + out.tlas_depth) + 2u > stack_entries - 6u) {
        err = "acceleration structure too deep for the traversal stack (BLAS " + std::to_string(out.max_blas_depth) + " + TLAS " +
              std::to_string(out.tlas_depth) + " wide levels)";
        return MPT_ERR_LIMIT;
    }

    // ---- geometry table
    out.geoms.resize(d->geometry_count);
    for (uint32_t g = 0; g < d->geometry_count; g++)
        if (d->geometries[g].flags & MPT_GEOM_HAS_PREV_VERTICES) {
            err = "MPT_GEOM_HAS_PREV_VERTICES (previous-frame vertices of morphed / skinned meshes, transform_prev_vertex.slang) is not supported";
            return MPT_ERR_UNSUPPORTED;
        }
    for (uint32_t g = 0; g < d->geometry_count; g++)
        out.geoms[g] = HostUVec4{uint32_t(d->geometries[g].material_id) | (uint32_t(d->geometries[g].flags) << 16), d->geometries[g].primitive_count, 0u, 0u};
    for (uint32_t i = 0; i < d->instance_count; i++) {
        const mpt_instance_desc& in = d->instances[i];
        const mpt_blas_desc& bd = d->blas[in.blas];
        for (uint32_t g = 0; g < bd.mesh_count; g++) {
            const uint32_t gid = in.instance_custom_index + g;
            if (gid >= d->geometry_count) { err = "instance_custom_index + geometry_index exceeds the geometry table"; return MPT_ERR_INVALID_ARGUMENT; }
            if (d->geometries[gid].mesh != bd.meshes[g]) { err = "geometry table does not match the BLAS mesh list"; return MPT_ERR_INVALID_ARGUMENT; }
            out.geoms[gid].z = blas_geom_offset[in.blas][g];
        }
    }
    out.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
    return MPT_OK;
}

} // namespace mpt
