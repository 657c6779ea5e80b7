// bvh8_build.cpp — host builder: binned-SAH BVH2 (one primitive per leaf) -> SAH-optimal 8-wide collapse (dynamic program of Ylitie,
// Karras, Laine 2017, section 4.1, restated from the paper) -> quantised 80-byte nodes.
// Replaces the driver BLAS/TLAS build of the reference (src/merian-shaders/scene/scene.cpp:1672-1892).
#include "bvh8.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>

namespace mpt {
namespace {

struct B3 {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    void grow(const Box& b) {
        for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); }
    }
    void grow(const B3& b) {
        for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); }
    }
    void grow_pt(const float* p) {
        for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); }
    }
    float half_area() const {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0) return 0.f;
        return dx * dy + dy * dz + dz * dx;
    }
};

struct Node2 {
    B3 box;
    uint32_t left = 0, right = 0; // internal children
    uint32_t first = 0, count = 0; // leaf range into prim index array (count > 0 => leaf)
    uint32_t rfirst = 0, rcount = 0; // range of every node (leaf or internal) in the prim index array
};

constexpr int NBINS = 16;


// Builder knobs for tuning sessions (host code only; defaults are the measured optimum).
float env_float(const char* name, float def) {
    const char* v = std::getenv(name);
    return v ? float(std::atof(v)) : def;
}

struct Builder2 {
    const Box* boxes;
    std::vector<float> cent; // 3 per prim
    std::vector<uint32_t> idx;
    std::vector<Node2> nodes;
    int max_leaf;
    uint32_t sah_depth_limit = 0xFFFFFFFFu; // BVH2 levels built with SAH before the builder switches to median splits
    float trav_cost = 1.0f; // SAH: cost of one node test in units of one primitive test
    float balance = 0.f;    // SAH candidates must leave at least this fraction of the primitives on the smaller side (0 = plain SAH)

    void build(uint32_t n) {
        idx.resize(n);
        std::iota(idx.begin(), idx.end(), 0u);
        cent.resize(size_t(n) * 3);
        for (uint32_t i = 0; i < n; i++)
            for (int a = 0; a < 3; a++) cent[size_t(i) * 3 + a] = 0.5f * (boxes[i].lo[a] + boxes[i].hi[a]);
        nodes.reserve(size_t(n) * 2);
        nodes.emplace_back();
        struct Job { uint32_t node, first, count, depth; };
        std::vector<Job> stack;
        stack.push_back({0, 0, n, 0});
        while (!stack.empty()) {
            const Job j = stack.back();
            stack.pop_back();
            B3 box, cbox;
            for (uint32_t i = j.first; i < j.first + j.count; i++) {
                box.grow(boxes[idx[i]]);
                cbox.grow_pt(&cent[size_t(idx[i]) * 3]);
            }
            nodes[j.node].box = box;
            nodes[j.node].rfirst = j.first;
            nodes[j.node].rcount = j.count;
            if (j.count <= 1) { make_leaf(j.node, j.first, j.count); continue; }

            // binned SAH over the 3 axes
            float best_cost = FLT_MAX;
            int best_axis = -1, best_bin = -1;
            for (int a = 0; a < 3; a++) {
                const float cmin = cbox.lo[a], cmax = cbox.hi[a];
                if (!(cmax > cmin)) continue;
                B3 bb[NBINS];
                uint32_t bc[NBINS] = {0};
                const float k = float(NBINS) * (1.0f - 1e-6f) / (cmax - cmin);
                for (uint32_t i = j.first; i < j.first + j.count; i++) {
                    const uint32_t p = idx[i];
                    int b = int((cent[size_t(p) * 3 + a] - cmin) * k);
                    b = std::min(std::max(b, 0), NBINS - 1);
                    bb[b].grow(boxes[p]);
                    bc[b]++;
                }
                float ra[NBINS];
                uint32_t rc[NBINS];
                B3 acc;
                uint32_t cnt = 0;
                for (int b = NBINS - 1; b > 0; b--) {
                    acc.grow(bb[b]);
                    cnt += bc[b];
                    ra[b] = acc.half_area();
                    rc[b] = cnt;
                }
                acc = B3();
                cnt = 0;
                for (int b = 0; b < NBINS - 1; b++) {
                    acc.grow(bb[b]);
                    cnt += bc[b];
                    if (cnt == 0 || rc[b + 1] == 0) continue;
                    if (float(std::min(cnt, rc[b + 1])) < balance * float(j.count)) continue;
                    const float cost = acc.half_area() * float(cnt) + ra[b + 1] * float(rc[b + 1]);
                    if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = b; }
                }
            }
            const float parent_area = box.half_area();
            const float leaf_cost = float(j.count);
            const float split_cost = (parent_area > 0.f && best_axis >= 0) ? trav_cost + best_cost / parent_area : FLT_MAX;
            if (int(j.count) <= max_leaf && split_cost >= leaf_cost) { make_leaf(j.node, j.first, j.count); continue; }

            uint32_t mid;
            if (j.depth >= sah_depth_limit) {
                // a degenerate SAH chain (one primitive peeled off per level) would outgrow the traversal stack: below this depth
                // split at the object median of the widest centroid axis, which bounds the remaining depth by log2(count)
                int a = 0;
                for (int k = 1; k < 3; k++)
                    if (cbox.hi[k] - cbox.lo[k] > cbox.hi[a] - cbox.lo[a]) a = k;
                mid = j.count / 2;
                std::nth_element(idx.begin() + j.first, idx.begin() + j.first + mid, idx.begin() + j.first + j.count,
                                 [&](uint32_t p, uint32_t q) { return cent[size_t(p) * 3 + a] < cent[size_t(q) * 3 + a]; });
            } else if (best_axis >= 0) {
                const int a = best_axis;
                const float cmin = cbox.lo[a], cmax = cbox.hi[a];
                const float k = float(NBINS) * (1.0f - 1e-6f) / (cmax - cmin);
                auto it = std::partition(idx.begin() + j.first, idx.begin() + j.first + j.count, [&](uint32_t p) {
                    int b = int((cent[size_t(p) * 3 + a] - cmin) * k);
                    b = std::min(std::max(b, 0), NBINS - 1);
                    return b <= best_bin;
                });
                mid = uint32_t(it - idx.begin()) - j.first;
            } else {
                mid = 0;
            }
            if (mid == 0 || mid == j.count) { // all centroids coincide: split the index range in half
                if (int(j.count) <= max_leaf) { make_leaf(j.node, j.first, j.count); continue; }
                mid = j.count / 2;
            }
            const uint32_t l = uint32_t(nodes.size());
            nodes.emplace_back();
            nodes.emplace_back();
            nodes[j.node].left = l;
            nodes[j.node].right = l + 1;
            nodes[j.node].count = 0;
            stack.push_back({l + 1, j.first + mid, j.count - mid, j.depth + 1});
            stack.push_back({l, j.first, mid, j.depth + 1});
        }
    }
    void make_leaf(uint32_t node, uint32_t first, uint32_t count) {
        nodes[node].first = first;
        nodes[node].count = count;
    }
};

inline uint8_t exponent_for(float extent) {
    // smallest e with extent / 2^e <= 255
    if (!(extent > 0.f)) return 1;
    int e = int(std::ceil(std::log2(double(extent) / 255.0)));
    while (double(extent) / std::ldexp(1.0, e) > 255.0) e++;
    e = std::min(std::max(e + 127, 1), 254);
    return uint8_t(e);
}

} // namespace

static void build_bvh8_once(const Box* boxes, uint32_t n, Bvh8& out, int max_leaf, uint32_t sah_depth_limit) {
    out.nodes.clear();
    out.prim_order.clear();
    B3 all;
    for (uint32_t i = 0; i < n; i++) all.grow(boxes[i]);
    for (int a = 0; a < 3; a++) { out.bounds.lo[a] = n ? all.lo[a] : 0.f; out.bounds.hi[a] = n ? all.hi[a] : 0.f; }
    if (n == 0) {
        Bvh8Node root;
        std::memset(&root, 0, sizeof(root));
        root.ex = root.ey = root.ez = 1;
        out.nodes.push_back(root);
        return;
    }
    Builder2 b2;
    b2.boxes = boxes;
    const int collapse_mode = int(env_float("MPT_BVH_COLLAPSE", 2.f)); // 2 = SAH-optimal dynamic program (default), 0 = greedy by area
    b2.max_leaf = collapse_mode == 2 ? 1 : std::min(std::max(int(env_float("MPT_BVH_MAX_LEAF", float(max_leaf))), 1), 3);
    b2.trav_cost = env_float("MPT_BVH_TRAV_COST", 1.0f);
    b2.sah_depth_limit = sah_depth_limit;
    b2.balance = env_float("MPT_BVH_BALANCE", 0.f);
    b2.build(n);
    const std::vector<Node2>& n2 = b2.nodes;

    // ---- dynamic program of the SAH-optimal collapse: cost[n][i-1] = C(n, i), the cheapest way to represent BVH2 subtree n as at most i
    // BVH8 children (leaves of <= 3 primitives or internal nodes); children have larger indices than their parent, so a reverse sweep
    // is bottom-up.
    std::vector<float> cost;
    std::vector<uint8_t> dec, dist8;
    if (collapse_mode == 2) {
        const float c_prim = env_float("MPT_BVH_CPRIM", 0.5f), c_node = 1.0f; // measured optimum on C3 (0.15 ... 2.0 within 0.6 %)
        const size_t nn = n2.size();
        cost.assign(nn * 7, 0.f);
        dec.assign(nn * 7, 0);
        dist8.assign(nn, 1);
        for (size_t k = nn; k-- > 0;) {
            const Node2& nd = n2[k];
            const float area = nd.box.half_area();
            float* C = &cost[k * 7];
            uint8_t* D = &dec[k * 7];
            if (nd.count > 0) { // BVH2 leaf (one primitive)
                for (int i = 0; i < 7; i++) { C[i] = area * float(nd.count) * c_prim; D[i] = 0; }
                continue;
            }
            const float* L = &cost[size_t(nd.left) * 7];
            const float* R = &cost[size_t(nd.right) * 7];
            auto distribute = [&](int j, uint8_t& best_k) { // min over 0 < k < j of C(left, k) + C(right, j - k), slots per side capped at 7
                float best = FLT_MAX;
                for (int kk = std::max(1, j - 7); kk <= std::min(j - 1, 7); kk++) {
                    const float v = L[kk - 1] + R[j - kk - 1];
                    if (v < best) { best = v; best_k = uint8_t(kk); }
                }
                return best;
            };
            uint8_t k8 = 1;
            const float c_internal = distribute(8, k8) + area * c_node;
            dist8[k] = k8;
            const float c_leaf = nd.rcount <= 3u ? area * float(nd.rcount) * c_prim : FLT_MAX;
            if (c_leaf <= c_internal) { C[0] = c_leaf; D[0] = 0; } else { C[0] = c_internal; D[0] = 1; }
            for (int i = 2; i <= 7; i++) {
                uint8_t kk = 1;
                const float cd = distribute(i, kk);
                if (cd < C[i - 2]) { C[i - 1] = cd; D[i - 1] = kk; } else { C[i - 1] = C[i - 2]; D[i - 1] = 0; }
            }
        }
    }

    out.prim_order.reserve(n);
    out.nodes.reserve(n / 3 + 16);
    struct Task { uint32_t wide, root2, depth; };
    std::vector<Task> queue;
    out.nodes.emplace_back();
    queue.push_back({0, 0, 1});
    out.depth = 1;
    for (size_t qi = 0; qi < queue.size(); qi++) {
        const Task t = queue[qi];
        out.depth = std::max(out.depth, t.depth);
        // ---- gather up to 8 children
        uint32_t ch[8];
        bool ch_leaf[8];
        int nch = 0;
        if (n2[t.root2].count > 0) {
            ch_leaf[nch] = true;
            ch[nch++] = t.root2; // the BVH2 root itself is a leaf
        } else if (collapse_mode == 2) {
            // SAH-optimal collapse (Ylitie, Karras, Laine 2017, section 4.1): follow the decisions of the dynamic program below
            struct Item { uint32_t node; int slots; };
            Item work[16];
            int nw = 0;
            const int k8 = dist8[t.root2];
            work[nw++] = {n2[t.root2].right, 8 - k8};
            work[nw++] = {n2[t.root2].left, k8};
            while (nw) {
                Item it = work[--nw];
                if (n2[it.node].count > 0) { ch_leaf[nch] = true; ch[nch++] = it.node; continue; }
                while (it.slots > 1 && dec[size_t(it.node) * 7 + size_t(it.slots - 1)] == 0) it.slots--;
                const uint8_t d = dec[size_t(it.node) * 7 + size_t(it.slots - 1)];
                if (it.slots == 1) { ch_leaf[nch] = (d == 0); ch[nch++] = it.node; continue; }
                work[nw++] = {n2[it.node].right, it.slots - int(d)};
                work[nw++] = {n2[it.node].left, int(d)};
            }
        } else {
            // greedy: repeatedly open the internal child with the largest area
            ch[nch++] = n2[t.root2].left;
            ch[nch++] = n2[t.root2].right;
            while (nch < 8) {
                int best = -1;
                float best_area = -1.f;
                for (int i = 0; i < nch; i++) {
                    if (n2[ch[i]].count > 0) continue;
                    float a = n2[ch[i]].box.half_area();
                    if (collapse_mode == 1) { // prefer opening nodes whose children are much smaller than they are
                        const Node2& c = n2[ch[i]];
                        a = a - 0.5f * (n2[c.left].box.half_area() + n2[c.right].box.half_area());
                    }
                    if (a > best_area) { best_area = a; best = i; }
                }
                if (best < 0) break;
                const uint32_t open = ch[best];
                ch[best] = n2[open].left;
                ch[nch++] = n2[open].right;
            }
            for (int i = 0; i < nch; i++) ch_leaf[i] = n2[ch[i]].count > 0;
        }
        // ---- slot assignment: slot s is visited first by rays with octant s (direction sign bit set = negative),
        // so it should hold the child that is nearest along d_s: minimise sum of dot(centroid_c - centroid_p, d_s).
        const B3& pb = n2[t.root2].box;
        float pc[3];
        for (int a = 0; a < 3; a++) pc[a] = 0.5f * (pb.lo[a] + pb.hi[a]);
        float cost[8][8];
        for (int c = 0; c < nch; c++) {
            const B3& cb = n2[ch[c]].box;
            float d[3];
            for (int a = 0; a < 3; a++) d[a] = 0.5f * (cb.lo[a] + cb.hi[a]) - pc[a];
            for (int s = 0; s < 8; s++)
                cost[c][s] = ((s & 1) ? -d[0] : d[0]) + ((s & 2) ? -d[1] : d[1]) + ((s & 4) ? -d[2] : d[2]);
        }
        int slot_of[8], child_in_slot[8];
        for (int i = 0; i < 8; i++) { slot_of[i] = -1; child_in_slot[i] = -1; }
        for (int k = 0; k < nch; k++) {
            float bc = FLT_MAX;
            int bi = -1, bs = -1;
            for (int c = 0; c < nch; c++) {
                if (slot_of[c] >= 0) continue;
                for (int s = 0; s < 8; s++) {
                    if (child_in_slot[s] >= 0) continue;
                    if (cost[c][s] < bc) { bc = cost[c][s]; bi = c; bs = s; }
                }
            }
            slot_of[bi] = bs;
            child_in_slot[bs] = bi;
        }
        // ---- encode
        Bvh8Node node;
        std::memset(&node, 0, sizeof(node));
        node.px = pb.lo[0]; node.py = pb.lo[1]; node.pz = pb.lo[2];
        node.ex = exponent_for(pb.hi[0] - pb.lo[0]);
        node.ey = exponent_for(pb.hi[1] - pb.lo[1]);
        node.ez = exponent_for(pb.hi[2] - pb.lo[2]);
        const double sc[3] = {std::ldexp(1.0, int(node.ex) - 127), std::ldexp(1.0, int(node.ey) - 127), std::ldexp(1.0, int(node.ez) - 127)};
        node.child_base = uint32_t(out.nodes.size());
        node.prim_base = uint32_t(out.prim_order.size());
        uint32_t n_internal = 0;
        uint8_t* qlo[3] = {node.qlox, node.qloy, node.qloz};
        uint8_t* qhi[3] = {node.qhix, node.qhiy, node.qhiz};
        for (int s = 0; s < 8; s++) {
            const int c = child_in_slot[s];
            if (c < 0) continue;
            const Node2& cn = n2[ch[c]];
            for (int a = 0; a < 3; a++) {
                const double lo = std::floor((double(cn.box.lo[a]) - double(pb.lo[a])) / sc[a]);
                const double hi = std::ceil((double(cn.box.hi[a]) - double(pb.lo[a])) / sc[a]);
                qlo[a][s] = uint8_t(std::min(std::max(lo, 0.0), 255.0));
                qhi[a][s] = uint8_t(std::min(std::max(hi, 0.0), 255.0));
            }
            if (!ch_leaf[c]) {
                node.imask |= uint8_t(1u << s);
                n_internal++;
            } else {
                const uint32_t unary = (1u << cn.rcount) - 1u; // 1, 3, 7
                node.leaf_bits |= unary << (3 * s);
                for (uint32_t k = 0; k < cn.rcount; k++) out.prim_order.push_back(b2.idx[cn.rfirst + k]);
            }
        }
        out.nodes[t.wide] = node;
        out.nodes.resize(out.nodes.size() + n_internal);
        uint32_t j = 0;
        for (int s = 0; s < 8; s++) {
            const int c = child_in_slot[s];
            if (c < 0 || ch_leaf[c]) continue;
            queue.push_back({node.child_base + j, ch[c], t.depth + 1});
            j++;
        }
    }
}

// The traversal keeps at most two stack entries per wide level (bvh8_traverse.cuh), so the depth must be bounded. SAH trees of
// real scenes are 8-10 wide levels deep (1 M-triangle terrain: 8); a degenerate input (primitives of geometrically growing size on a
// line) makes SAH peel one primitive per level. When that happens the tree is rebuilt with SAH only for the top levels and object-median
// splits below, which bounds the depth by about log8(n) + the SAH levels.
void build_bvh8(const Box* boxes, uint32_t n, Bvh8& out, int max_leaf) {
    if (std::getenv("MPT_BVH_SAH_DEPTH")) { // tuning sessions: SAH for the top levels only, object-median splits below
        build_bvh8_once(boxes, n, out, max_leaf, uint32_t(env_float("MPT_BVH_SAH_DEPTH", 64.f)));
        if (out.depth <= kMaxBvh8Depth) return;
    }
    for (uint32_t limit : {0xFFFFFFFFu, 16u, 8u, 0u}) {
        build_bvh8_once(boxes, n, out, max_leaf, limit);
        if (out.depth <= kMaxBvh8Depth) return;
    }
}

} // namespace mpt
