// scene_build.h — host side of mpt_scene_upload: flattens an mpt_scene_desc into the arrays the device traversal reads
// (BVH8 nodes of every BLAS + the TLAS, triangle and attribute records in leaf order, instance records, geometry table).
// Host-only (no CUDA): the same function feeds mpt_scene_upload and the host-only export mpt_host_accel_* that tests and the
// CPU oracle's BVH8 walker use ("the instrumented CPU oracle traversing the same BVH8").
// Replaces Scene::build_blas / build_tlas of the reference (src/merian-shaders/scene/scene.cpp:1672-1892).
#pragma once

#include "../../include/mpt.h"
#include "bvh8.h"

#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace mpt {

struct alignas(16) HostVec4 {
    float x, y, z, w;
};
struct alignas(16) HostUVec4 {
    uint32_t x, y, z, w;
};
struct alignas(16) HostInstanceRec { // same 64 bytes as InstanceRec in bvh8_traverse.cuh
    float inv[12];                   // rows of the 3x4 world -> object matrix
    uint32_t blas_root, custom_index, mask_flags, pad;
};
static_assert(sizeof(HostInstanceRec) == 64, "instance record is 64 bytes");

// resize() without the zero fill: the large per-triangle arrays are written exactly once, by the parallel loops of build_host_accel
// (zeroing 100 MB on one thread, page faults included, was a third of the build time of a 1 M-triangle scene)
template <class T> struct NoInitAllocator : std::allocator<T> {
    template <class U> struct rebind { using other = NoInitAllocator<U>; };
    NoInitAllocator() = default;
    template <class U> NoInitAllocator(const NoInitAllocator<U>&) {}
    template <class U, class... A> void construct(U* p, A&&... a) {
        if constexpr (sizeof...(A) == 0) ::new (static_cast<void*>(p)) U;
        else ::new (static_cast<void*>(p)) U(std::forward<A>(a)...);
    }
};
template <class T> using RawVector = std::vector<T, NoInitAllocator<T>>;

struct HostAccel {
    std::vector<Bvh8Node> nodes;
    RawVector<HostVec4> tris;   // 3 per record: v0|primitive_id, v1|geometry_index, v2|flags
    RawVector<HostUVec4> attrs; // 3 per record
    RawVector<uint32_t> prim_to_tri;
    std::vector<HostInstanceRec> instances;
    std::vector<uint32_t> tlas_prims;
    std::vector<HostVec4> tlas_boxes; // 2 per TLAS primitive record: lo, hi of the instance's world box (unquantised), two-level scenes only
    std::vector<HostUVec4> geoms;
    uint32_t root = 0;
    uint32_t two_level = 0;
    uint32_t max_blas_depth = 0, tlas_depth = 0;
    double build_ms = 0.0;
};

constexpr uint32_t kTriNonOpaqueHost = 1u; // v2.w flag bit (kTriNonOpaque on the device)

// Returns MPT_OK or an error code with `err` set. stack_entries = capacity of the traversal stack the result must fit.
mpt_status build_host_accel(const mpt_scene_desc* d, HostAccel& out, uint32_t stack_entries, std::string& err);

} // namespace mpt
