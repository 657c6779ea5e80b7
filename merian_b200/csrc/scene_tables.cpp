// scene_tables.cpp — mptt_* (include/mpt_tables.h): the host logic of merian::Scene that defines the ids and vertex bits `pt` sees,
// for hosts without a merian::Scene. Host-only; compiled with -ffp-contract=off so that every product and sum below is rounded on its
// own, like the numpy mirror (merian_b200/scene.py) the tests compare it with.
//   grouping            src/merian-shaders/scene/scene.cpp:1023-1105
//   GeometryData order  src/merian-shaders/scene/scene.cpp:1236-1291
//   TLAS instances      src/merian-shaders/scene/scene.cpp:1808-1857
//   vertex packing      src/merian-shaders/scene/gltf_scene.cpp:107-118, src/merian/utils/normal_encoding.cpp:9-25,
//                       src/merian-shaders/scene/transform_vertex.slang:12-40, include/merian-shaders/utils/encoding.slang:6-30
#include "../../include/mpt_tables.h"
#include "mpt_internal.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <new>
#include <string>
#include <vector>

namespace {

constexpr uint32_t kGroupSplitMask = MPT_MESH_FLIP_FACING | MPT_MESH_TWO_SIDED; // scene.hpp:350-351

struct V2 { float x, y; };
struct V3 { float x, y, z; };

// octahedral_project, encoding.slang:17-23
V2 oct_project(V3 n) {
    const float inv = 1.0f / ((std::fabs(n.x) + std::fabs(n.y)) + std::fabs(n.z));
    const float px = n.x * inv, py = n.y * inv;
    if (n.z < 0.f) return V2{std::copysign(1.0f - std::fabs(py), n.x), std::copysign(1.0f - std::fabs(px), n.y)};
    return V2{px, py};
}
// merian::encode_normal on the host, normal_encoding.cpp:9-25: x 32768, roundf (half away from zero), clamp to int16
uint32_t host_encode_normal(V3 n) {
    const V2 p = oct_project(n);
    auto enc = [](float v) {
        float e = v * 32768.0f;
        e = std::min(std::max(e, -32768.0f), 32767.0f);
        const float r = e >= 0.f ? std::floor(e + 0.5f) : std::ceil(e - 0.5f);
        return uint32_t(int32_t(r)) & 0xFFFFu;
    };
    return enc(p.x) | (enc(p.y) << 16);
}
V3 normalize(V3 v) {
    const float l = std::sqrt((v.x * v.x + v.y * v.y) + v.z * v.z);
    return V3{v.x / l, v.y / l, v.z / l};
}
// decode_normal, encoding.slang:6-13 (unpackSnorm2x16: int16 / 32767, clamped)
V3 device_decode_normal(uint32_t enc) {
    const float a = float(int16_t(enc & 0xFFFFu)), b = float(int16_t(enc >> 16));
    const float x = std::min(std::max(a / 32767.0f, -1.0f), 1.0f), y = std::min(std::max(b / 32767.0f, -1.0f), 1.0f);
    const float z = (1.0f - std::fabs(x)) - std::fabs(y);
    V3 v{x, y, z};
    if (z < 0.f) v = V3{std::copysign(1.0f - std::fabs(y), x), std::copysign(1.0f - std::fabs(x), y), z};
    return normalize(v);
}
// encode_normal, encoding.slang:28-30 (packSnorm2x16: round(clamp(v, -1, 1) * 32767), ties to even)
uint32_t device_encode_normal(V3 n) {
    const V2 p = oct_project(n);
    auto enc = [](float v) {
        v = std::min(std::max(v, -1.0f), 1.0f);
        return uint32_t(int32_t(std::nearbyintf(v * 32767.0f))) & 0xFFFFu;
    };
    return enc(p.x) | (enc(p.y) << 16);
}
// float -> IEEE binary16, round to nearest even (what `half2(uv)` stores, gltf_scene.cpp:113)
uint16_t float_to_half(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7FFFFFFFu;
    if (x >= 0x7F800000u) return uint16_t(sign | 0x7C00u | (x > 0x7F800000u ? 0x200u : 0u)); // inf / nan
    if (x >= 0x477FF000u) return uint16_t(sign | 0x7C00u);                                      // rounds to inf
    if (x < 0x33000001u) return uint16_t(sign);                                                 // rounds to zero
    const int e = int(x >> 23) - 127;
    uint32_t m = (x & 0x7FFFFFu) | 0x800000u;
    int shift;
    uint32_t base;
    if (e < -14) { shift = 13 + (-14 - e); base = 0; } // subnormal half
    else { shift = 13; base = uint32_t(e + 15) << 10; m &= 0x7FFFFFu; }
    const uint32_t q = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
    uint32_t h = base + q;
    if (rem > half || (rem == half && (h & 1u))) h++;
    return uint16_t(sign | h);
}

struct M4 { float m[16]; };
const M4 kIdentity{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}};

// inverse transpose of a 4x4 in double (Gauss-Jordan with partial pivoting), rounded to float. False if singular.
bool inverse_transpose(const float* M, float* out) {
    double a[4][8];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) { a[r][c] = double(M[r * 4 + c]); a[r][4 + c] = r == c ? 1.0 : 0.0; }
    for (int c = 0; c < 4; c++) {
        int piv = c;
        for (int r = c + 1; r < 4; r++)
            if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
        if (a[piv][c] == 0.0) return false;
        if (piv != c)
            for (int k = 0; k < 8; k++) std::swap(a[piv][k], a[c][k]);
        const double d = a[c][c];
        for (int k = 0; k < 8; k++) a[c][k] /= d;
        for (int r = 0; r < 4; r++) {
            if (r == c) continue;
            const double f = a[r][c];
            if (f == 0.0) continue;
            for (int k = 0; k < 8; k++) a[r][k] -= f * a[c][k];
        }
    }
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) out[c * 4 + r] = float(a[r][4 + c]); // transposed
    return true;
}

bool pack_vertices(const float* positions, const float* normals, const float* uvs, const float* tangents, uint32_t n, const float* transform,
                   mpt_packed_vertex* out) {
    const float* M = transform ? transform : kIdentity.m;
    float Mit[16];
    if (!inverse_transpose(M, Mit)) return false;
#pragma omp parallel for schedule(static) if (n > 20000)
    for (int64_t ii = 0; ii < int64_t(n); ii++) {
        const size_t i = size_t(ii);
        const V3 p{positions[i * 3], positions[i * 3 + 1], positions[i * 3 + 2]};
        // ---- host packing, gltf_scene.cpp:107-118
        const uint32_t host_n = host_encode_normal(V3{normals[i * 3], normals[i * 3 + 1], normals[i * 3 + 2]});
        uint32_t host_t = 0u; // no tangents: zero word
        if (tangents) {
            const V3 t{tangents[i * 4], tangents[i * 4 + 1], tangents[i * 4 + 2]};
            const float l = std::sqrt((t.x * t.x + t.y * t.y) + t.z * t.z);
            const V3 tn = l > 0.f ? V3{t.x / l, t.y / l, t.z / l} : V3{0.f, 0.f, 1.f};
            host_t = (host_encode_normal(tn) & 0xFFFFFFFEu) | (tangents[i * 4 + 3] < 0.f ? 1u : 0u);
        }
        // ---- device pass, transform_vertex.slang:12-40
        mpt_packed_vertex& v = out[i];
        for (int r = 0; r < 3; r++) v.position[r] = ((M[r * 4] * p.x + M[r * 4 + 1] * p.y) + M[r * 4 + 2] * p.z) + M[r * 4 + 3];
        const V3 nd = device_decode_normal(host_n);
        v.encoded_normal = device_encode_normal(normalize(V3{(Mit[0] * nd.x + Mit[1] * nd.y) + Mit[2] * nd.z, (Mit[4] * nd.x + Mit[5] * nd.y) + Mit[6] * nd.z,
                                                             (Mit[8] * nd.x + Mit[9] * nd.y) + Mit[10] * nd.z}));
        v.uv[0] = uvs ? float_to_half(uvs[i * 2]) : uint16_t(0);
        v.uv[1] = uvs ? float_to_half(uvs[i * 2 + 1]) : uint16_t(0);
        const V3 td = device_decode_normal(host_t & 0xFFFFFFFEu);
        const V3 tw = normalize(V3{(M[0] * td.x + M[1] * td.y) + M[2] * td.z, (M[4] * td.x + M[5] * td.y) + M[6] * td.z, (M[8] * td.x + M[9] * td.y) + M[10] * td.z});
        v.encoded_tangent = (device_encode_normal(tw) & 0xFFFFFFFEu) | (host_t & 1u);
    }
    return true;
}

struct Group {
    std::vector<uint32_t> meshes;    // SmallSet<MeshID>: ascending
    std::vector<uint32_t> instances; // nodes, = the instances of the FIRST mesh (MeshGroup::get_instances, scene.hpp:380-383)
    uint32_t flags = 0, mask = 0;
    bool all_opaque = true, instanced = false;
};

} // namespace

struct mpt_tables {
    std::vector<std::vector<mpt_packed_vertex>> vertices;
    std::vector<std::vector<uint8_t>> indices;
    std::vector<mpt_mesh_desc> meshes;
    std::vector<std::vector<uint32_t>> blas_meshes;
    std::vector<mpt_blas_desc> blas;
    std::vector<mpt_geometry_desc> geometries;
    std::vector<mpt_instance_desc> instances;
    std::vector<float> xf, xf_it;
    mpt_scene_desc desc{};
};

static mpt_status tables_fail(mpt_status s, const std::string& msg) { return mpt_internal_fail(nullptr, s, msg.c_str()); }

MPT_API mpt_status mptt_pack_vertices(const float* positions, const float* normals, const float* uvs, const float* tangents, uint32_t vertex_count,
                                      const float* transform, mpt_packed_vertex* out) try {
    if ((!positions || !normals || !out) && vertex_count) return tables_fail(MPT_ERR_INVALID_ARGUMENT, "mptt_pack_vertices: NULL argument");
    if (!pack_vertices(positions, normals, uvs, tangents, vertex_count, transform, out)) return tables_fail(MPT_ERR_INVALID_ARGUMENT, "mptt_pack_vertices: singular transform");
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return tables_fail(MPT_ERR_LIMIT, "mptt_pack_vertices: out of host memory");
} catch (...) { // no C++ exception crosses the C ABI
    return tables_fail(MPT_ERR_INVALID_ARGUMENT, "mptt_pack_vertices: unexpected exception");
}

MPT_API mpt_status mptt_build(const mpt_src_mesh* meshes, uint32_t mesh_count, const mpt_src_node* nodes, uint32_t node_count, mpt_tables** out) try {
    if (!out || (!meshes && mesh_count) || (!nodes && node_count)) return tables_fail(MPT_ERR_INVALID_ARGUMENT, "mptt_build: NULL argument");
    *out = nullptr;
    // ---- instances of every mesh: SmallSet<NodeID> = ascending, unique (small_set.hpp:13)
    std::vector<std::vector<uint32_t>> instances_of(mesh_count);
    for (uint32_t nid = 0; nid < node_count; nid++) {
        if (nodes[nid].mesh_count && !nodes[nid].meshes) return tables_fail(MPT_ERR_INVALID_ARGUMENT, "mptt_build: node with a mesh count but no mesh list");
        for (uint32_t k = 0; k < nodes[nid].mesh_count; k++) {
            const uint32_t m = nodes[nid].meshes[k];
            if (m >= mesh_count) return tables_fail(MPT_ERR_INVALID_ARGUMENT, "mptt_build: node references a mesh out of range");
            if (instances_of[m].empty() || instances_of[m].back() != nid) instances_of[m].push_back(nid); // nid ascends
        }
    }
    // ---- grouping, scene.cpp:1023-1105
    std::vector<Group> groups;
    std::map<uint32_t, uint32_t> groups_static;
    std::map<std::vector<uint32_t>, std::map<uint32_t, uint32_t>> groups_instanced;
    for (uint32_t mid = 0; mid < mesh_count; mid++) { // ascending MeshID
        const std::vector<uint32_t>& inst = instances_of[mid];
        if (inst.empty()) continue;
        const mpt_src_mesh& mesh = meshes[mid];
        if (mesh.flags & (MPT_MESH_IS_MORPHED | MPT_MESH_HAS_VARIABLE_TOPOLOGY))
            return tables_fail(MPT_ERR_UNSUPPORTED, "mptt_build: morphed / variable-topology meshes are not supported");
        if (!mesh.positions || !mesh.normals) return tables_fail(MPT_ERR_INVALID_ARGUMENT, "mptt_build: mesh without positions or normals");
        if (mesh.indices && mesh.index_size != 1 && mesh.index_size != 2 && mesh.index_size != 4)
            return tables_fail(MPT_ERR_INVALID_ARGUMENT, "mptt_build: index_size must be 1, 2 or 4");
        if (mesh.material_id > 0xFFFFu) return tables_fail(MPT_ERR_LIMIT, "mptt_build: material_id exceeds 16 bits (GeometryData, scene-data.slangh:119-129)");
        const uint32_t key = (mesh.flags & kGroupSplitMask) | ((mesh.instance_mask & 0xFFu) << 16);
        std::map<uint32_t, uint32_t>& table = inst.size() > 1 ? groups_instanced[inst] : groups_static;
        auto it = table.find(key);
        if (it == table.end()) {
            it = table.emplace(key, uint32_t(groups.size())).first;
            groups.emplace_back();
        }
        Group& g = groups[it->second];
        if (inst.size() > 1) g.instanced = true;
        g.flags = key & kGroupSplitMask;
        g.mask = mesh.instance_mask & 0xFFu;
        g.all_opaque = g.all_opaque && (mesh.flags & MPT_MESH_IS_OPAQUE);
        g.meshes.push_back(mid);
        if (g.meshes.size() == 1) g.instances = inst;
    }

    std::unique_ptr<mpt_tables> T(new mpt_tables);
    T->vertices.resize(mesh_count);
    T->indices.resize(mesh_count);
    T->meshes.assign(mesh_count, mpt_mesh_desc{});
    // ---- device vertex / index data per mesh (A13); a pre-transformed (static) mesh bakes the transform of ITS OWN node (scene.cpp:1498-1501)
    for (const Group& g : groups)
        for (uint32_t mid : g.meshes) {
            const mpt_src_mesh& mesh = meshes[mid];
            const float* M = g.instanced ? nullptr : nodes[instances_of[mid][0]].transform;
            T->vertices[mid].resize(mesh.vertex_count);
            if (!pack_vertices(mesh.positions, mesh.normals, mesh.uvs, mesh.tangents, mesh.vertex_count, M, T->vertices[mid].data()))
                return tables_fail(MPT_ERR_INVALID_ARGUMENT, "mptt_build: singular node transform");
            mpt_mesh_desc& md = T->meshes[mid];
            md.vertices = T->vertices[mid].data();
            md.vertex_count = mesh.vertex_count;
            md.opaque = (mesh.flags & MPT_MESH_IS_OPAQUE) ? 1u : 0u;
            if (mesh.indices) {
                const size_t bytes = size_t(mesh.primitive_count) * 3 * mesh.index_size;
                T->indices[mid].assign(static_cast<const uint8_t*>(mesh.indices), static_cast<const uint8_t*>(mesh.indices) + bytes);
                md.indices = T->indices[mid].data();
                md.primitive_count = mesh.primitive_count;
                md.index_flag = mesh.index_size == 1 ? MPT_GEOM_INDEX_U8 : (mesh.index_size == 2 ? MPT_GEOM_INDEX_U16 : MPT_GEOM_INDEX_U32);
            } else {
                md.indices = nullptr;
                md.primitive_count = mesh.vertex_count / 3;
                md.index_flag = 0;
            }
        }
    // ---- GeometryData[] + TLAS instances, the same iteration order for both (scene.cpp:1236-1291, 1808-1857)
    uint32_t instance_id = 0;
    T->blas_meshes.resize(groups.size());
    T->blas.resize(groups.size());
    for (size_t gid = 0; gid < groups.size(); gid++) {
        const Group& g = groups[gid];
        T->blas_meshes[gid] = g.meshes;
        T->blas[gid] = mpt_blas_desc{T->blas_meshes[gid].data(), uint32_t(g.meshes.size())};
        const bool pre = !g.instanced;
        for (uint32_t nid : g.instances) {
            const float* Tn = nodes[nid].transform;
            const bool is_ident = std::memcmp(Tn, kIdentity.m, sizeof(kIdentity.m)) == 0;
            for (uint32_t mid : g.meshes) {
                const mpt_src_mesh& mesh = meshes[mid];
                uint32_t f = T->meshes[mid].index_flag;
                if (pre || is_ident) f |= MPT_GEOM_PRETRANSFORMED;
                if (mesh.flags & MPT_MESH_FLIP_FACING) f |= MPT_GEOM_FLIP_FACING;
                if (mesh.flags & MPT_MESH_HAS_TANGENTS) f |= MPT_GEOM_HAS_TANGENTS;
                if (mesh.flags & MPT_MESH_FLAT_SHADING) f |= MPT_GEOM_FLAT_SHADING;
                if (mesh.flags & MPT_MESH_USE_ENV_MAP) f |= MPT_GEOM_USE_ENV_MAP;
                if (mesh.flags & MPT_MESH_TWO_SIDED) f |= MPT_GEOM_TWO_SIDED;
                T->geometries.push_back(mpt_geometry_desc{mid, uint16_t(mesh.material_id), uint16_t(f), T->meshes[mid].primitive_count});
            }
            uint32_t iflags = 0;
            if (g.flags & MPT_MESH_FLIP_FACING) iflags |= MPT_INSTANCE_FLIP_FACING;
            if (g.flags & MPT_MESH_TWO_SIDED) iflags |= MPT_INSTANCE_FACING_CULL_DISABLE;
            if (g.all_opaque) iflags |= MPT_INSTANCE_FORCE_OPAQUE;
            const float* Ti = pre ? kIdentity.m : Tn;
            mpt_instance_desc in{};
            std::memcpy(in.transform, Ti, 12 * sizeof(float));
            in.instance_custom_index = instance_id;
            in.mask = g.mask;
            in.flags = iflags;
            in.blas = uint32_t(gid);
            T->instances.push_back(in);
            T->xf.insert(T->xf.end(), Ti, Ti + 16);
            float it16[16];
            if (!inverse_transpose(Ti, it16)) return tables_fail(MPT_ERR_INVALID_ARGUMENT, "mptt_build: singular node transform");
            T->xf_it.insert(T->xf_it.end(), it16, it16 + 16);
            instance_id += uint32_t(g.meshes.size());
        }
    }
    if (instance_id > 0xFFFFFFu) return tables_fail(MPT_ERR_LIMIT, "mptt_build: instanceCustomIndex exceeds 24 bits");
    mpt_scene_desc& d = T->desc;
    d.meshes = T->meshes.data(); d.mesh_count = mesh_count;
    d.blas = T->blas.data(); d.blas_count = uint32_t(T->blas.size());
    d.geometries = T->geometries.data(); d.geometry_count = uint32_t(T->geometries.size());
    d.instances = T->instances.data(); d.instance_count = uint32_t(T->instances.size());
    d.instance_transforms = T->xf.data();
    d.inverse_transposed_instance_transforms = T->xf_it.data();
    *out = T.release();
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return tables_fail(MPT_ERR_LIMIT, "mptt_build: out of host memory");
} catch (...) { // no C++ exception crosses the C ABI
    return tables_fail(MPT_ERR_INVALID_ARGUMENT, "mptt_build: unexpected exception");
}

MPT_API const mpt_scene_desc* mptt_scene_desc(const mpt_tables* t) { return t ? &t->desc : nullptr; }

MPT_API void mptt_free(mpt_tables* t) { delete t; }

// ---------------------------------------------------------------------------------------------------------------- material buffer
// MaterialSystem::add_material / repack_host_buffer (material_system.cpp:124-195), payloads of diffuse-material.slang:8-12 and
// gltf_material.hpp:93-146.
struct mpt_material_table {
    std::vector<uint8_t> entries;
    std::vector<mpt_material_type> types;
    uint32_t entry_size = 0, count = 0;
};

namespace {
struct Payload {
    std::vector<uint8_t> b;
    void f32(float v) { const size_t n = b.size(); b.resize(n + 4); std::memcpy(&b[n], &v, 4); }
    void u16(uint32_t v) { const uint16_t h = uint16_t(v); const size_t n = b.size(); b.resize(n + 2); std::memcpy(&b[n], &h, 2); }
};
} // namespace

MPT_API mpt_status mptt_materials_build(const mpt_src_material* materials, uint32_t material_count, mpt_material_table** out) try {
    if (!out || (!materials && material_count)) return tables_fail(MPT_ERR_INVALID_ARGUMENT, "mptt_materials_build: NULL argument");
    *out = nullptr;
    std::unique_ptr<mpt_material_table> T(new mpt_material_table);
    std::vector<Payload> payloads(material_count);
    std::vector<uint32_t> type_of(material_count), alpha_tex(material_count);
    size_t max_payload = 0;
    for (uint32_t i = 0; i < material_count; i++) {
        const mpt_src_material& m = materials[i];
        if (m.model != MPT_MATERIAL_DIFFUSE && m.model != MPT_MATERIAL_GLTF) return tables_fail(MPT_ERR_INVALID_ARGUMENT, "mptt_materials_build: unknown material model");
        const uint32_t features = m.model == MPT_MATERIAL_GLTF ? (m.features & 63u) : 0u;
        uint32_t tid = 0;
        for (; tid < T->types.size(); tid++) // type ids in first-registration order (material_system.cpp:124-139)
            if (T->types[tid].model == m.model && T->types[tid].features == features) break;
        if (tid == T->types.size()) T->types.push_back(mpt_material_type{m.model, features});
        if (tid > 0xFFFFu || m.base_color_texture > 0xFFFFu) return tables_fail(MPT_ERR_LIMIT, "mptt_materials_build: type / texture ids are 16 bit");
        type_of[i] = tid;
        alpha_tex[i] = m.base_color_texture;
        Payload& p = payloads[i];
        for (int k = 0; k < 4; k++) p.f32(m.base_color_factor[k]);
        if (m.model == MPT_MATERIAL_GLTF) {
            for (int k = 0; k < 3; k++) p.f32(m.emissive_factor[k]);
            p.f32(m.metallic_factor); p.f32(m.roughness_factor); p.f32(m.normal_scale); p.f32(m.occlusion_strength); p.f32(m.ior);
            p.u16(m.metallic_roughness_texture); p.u16(m.normal_texture); p.u16(m.occlusion_texture); p.u16(m.emissive_texture);
            if (features & MPT_GLTF_TRANSMISSION) p.f32(m.transmission_weight);
            if (features & MPT_GLTF_VOLUME) for (int k = 0; k < 3; k++) p.f32(m.volume_absorption[k]);
            if (features & MPT_GLTF_CLEARCOAT) { p.f32(m.clearcoat_weight); p.f32(m.clearcoat_roughness); p.u16(m.clearcoat_texture); p.u16(m.clearcoat_roughness_texture); }
            if (features & MPT_GLTF_SHEEN) {
                for (int k = 0; k < 3; k++) p.f32(m.sheen_color[k]);
                p.f32(m.sheen_roughness); p.u16(m.sheen_color_texture); p.u16(m.sheen_roughness_texture);
            }
            if (features & MPT_GLTF_IRIDESCENCE) {
                p.f32(m.iridescence_factor); p.f32(m.iridescence_ior); p.f32(m.iridescence_thickness_min); p.f32(m.iridescence_thickness_max);
                p.u16(m.iridescence_texture); p.u16(m.iridescence_thickness_texture);
            }
            if (features & MPT_GLTF_ANISOTROPY) { p.f32(m.anisotropy_strength); p.f32(m.anisotropy_rotation); p.u16(m.anisotropy_texture); }
        }
        max_payload = std::max(max_payload, p.b.size());
    }
    T->entry_size = uint32_t(4 + ((max_payload + 3) / 4) * 4); // header + payload_size_in_uints * 4
    T->count = material_count;
    T->entries.assign(size_t(T->entry_size) * material_count, uint8_t(0));
    for (uint32_t i = 0; i < material_count; i++) {
        uint8_t* e = &T->entries[size_t(i) * T->entry_size];
        const uint16_t hdr[2] = {uint16_t(type_of[i]), uint16_t(alpha_tex[i])};
        std::memcpy(e, hdr, 4);
        std::memcpy(e + 4, payloads[i].b.data(), payloads[i].b.size());
    }
    *out = T.release();
    return MPT_OK;
} catch (const std::bad_alloc&) {
    return tables_fail(MPT_ERR_LIMIT, "mptt_materials_build: out of host memory");
} catch (...) { // no C++ exception crosses the C ABI
    return tables_fail(MPT_ERR_INVALID_ARGUMENT, "mptt_materials_build: unexpected exception");
}

MPT_API mpt_status mptt_materials_get(const mpt_material_table* t, mpt_material_table_view* out) {
    if (!t || !out) return tables_fail(MPT_ERR_INVALID_ARGUMENT, "mptt_materials_get: NULL argument");
    out->entries = t->entries.data();
    out->entry_size = t->entry_size;
    out->material_count = t->count;
    out->types = t->types.data();
    out->type_count = uint32_t(t->types.size());
    return MPT_OK;
}

MPT_API void mptt_materials_free(mpt_material_table* t) { delete t; }

// ---------------------------------------------------------------------------------------------------------------- camera
MPT_API mpt_status mptt_camera_look_at(const float position[3], const float target[3], const float up[3], float yfov_degrees, float aspect,
                                       float near_plane, float far_plane, float aperture_radius, const float jitter[2], mpt_camera* out) {
    if (!position || !target || !up || !out) return tables_fail(MPT_ERR_INVALID_ARGUMENT, "mptt_camera_look_at: NULL argument");
    float fwd[3] = {target[0] - position[0], target[1] - position[1], target[2] - position[2]};
    const float focus = std::sqrt((fwd[0] * fwd[0] + fwd[1] * fwd[1]) + fwd[2] * fwd[2]); // focus_distance_on_target, camera.cpp:255-260
    if (!(focus > 0.f)) return tables_fail(MPT_ERR_INVALID_ARGUMENT, "mptt_camera_look_at: position and target coincide");
    for (float& f : fwd) f /= focus;
    float rgt[3] = {fwd[1] * up[2] - fwd[2] * up[1], fwd[2] * up[0] - fwd[0] * up[2], fwd[0] * up[1] - fwd[1] * up[0]};
    const float rl = std::sqrt((rgt[0] * rgt[0] + rgt[1] * rgt[1]) + rgt[2] * rgt[2]);
    if (!(rl > 0.f)) return tables_fail(MPT_ERR_INVALID_ARGUMENT, "mptt_camera_look_at: up is parallel to the viewing direction");
    for (float& r : rgt) r /= rl;
    const float upc[3] = {rgt[1] * fwd[2] - rgt[2] * fwd[1], rgt[2] * fwd[0] - rgt[0] * fwd[2], rgt[0] * fwd[1] - rgt[1] * fwd[0]};
    const float half_tan = std::tan(yfov_degrees * 0.017453292519943295f * 0.5f);
    const float su = (half_tan * aspect) * focus, sv = half_tan * focus;
    std::memset(out, 0, sizeof(*out));
    for (int k = 0; k < 3; k++) {
        out->position[k] = position[k];
        out->target[k] = target[k];
        out->up[k] = up[k];
        out->U[k] = rgt[k] * su;
        out->V[k] = upc[k] * sv;
        out->W[k] = fwd[k] * focus;
    }
    out->near_plane = near_plane;
    out->far_plane = far_plane;
    out->aspect_ratio = aspect;
    out->aperture_radius = aperture_radius;
    out->jitter[0] = jitter ? jitter[0] : 0.f;
    out->jitter[1] = jitter ? jitter[1] : 0.f;
    return MPT_OK;
}
