// bvh8.h — compressed 8-wide BVH: host-side builder interface and the node / record layouts shared with
// the device traversal (bvh8_traverse.cuh).
//
// This replaces the driver-side acceleration structure the reference builds with
// vkCmdBuildAccelerationStructuresKHR (src/merian-shaders/scene/scene.cpp:1672-1892): B200 has no RT cores,
// so traversal is software over a layout chosen for 128-bit loads (Ylitie, Karras, Laine: "Efficient
// Incoherent Ray Traversal on GPUs Through Compressed Wide BVHs", HPG 2017 — restated from the paper).
#pragma once

#include <cstdint>
#include <vector>

namespace mpt {

// 80 bytes = 5 x 16-byte loads.
struct alignas(16) Bvh8Node {
    float px, py, pz;         // quantisation origin (node box lo)
    uint8_t ex, ey, ez;       // per-axis exponents: scale = 2^(e - 127), stored biased so float bits = e << 23
    uint8_t imask;            // bit i set: child slot i is an internal node
    uint32_t child_base;      // index of the first internal child (absolute into the node array)
    uint32_t prim_base;       // index of the first primitive record referenced by leaf children (absolute)
    uint32_t leaf_bits;       // 3 bits per child slot s at 3s: unary count (1, 3, 7) of a leaf child's primitives, 0 for inner / empty slots.
                              // The records of the leaf children lie back to back in slot order from prim_base: the primitive at bit b
                              // is record prim_base + popcount(leaf_bits & ((1 << b) - 1)). Fixed bit positions let the traversal turn
                              // its 8-bit "child hit" mask into primitive bits with three shift-or steps instead of a shift per child.
    uint32_t reserved;
    uint8_t qlox[8], qloy[8], qloz[8];
    uint8_t qhix[8], qhiy[8], qhiz[8];
};
static_assert(sizeof(Bvh8Node) == 80, "BVH8 node must be 80 bytes");

struct Box {
    float lo[3], hi[3];
};

// Result of building one BVH8 over `n` primitive boxes: nodes (node 0 = root; child_base relative to this
// vector) and the primitive order: prim_order[k] = input primitive stored at record k (prim_base relative).
struct Bvh8 {
    std::vector<Bvh8Node> nodes;
    std::vector<uint32_t> prim_order;
    Box bounds;
    uint32_t depth = 1; // levels of wide nodes on the longest root-to-leaf path (bounds the traversal stack)
};

// Depth bound build_bvh8 guarantees (by falling back to median splits below the SAH levels for degenerate inputs).
constexpr uint32_t kMaxBvh8Depth = 14;

// max_leaf: primitives per leaf child (<= 3). Deterministic (no threads inside).
// c_prim: cost of testing one primitive relative to one node test in the dynamic program of the wide collapse; < 0 = the default for
// triangles (0.5, MPT_BVH_CPRIM). The TLAS passes a larger value: "testing" an instance means transforming the ray, a ray setup and at
// least one node test of its BLAS.
void build_bvh8(const Box* boxes, uint32_t n, Bvh8& out, int max_leaf = 3, float c_prim = -1.f);

} // namespace mpt
