// bvh8_build.cpp — host builder: binned-SAH BVH2 (one primitive per leaf) -> SAH-optimal 8-wide collapse (dynamic program of Ylitie,
// Karras, Laine 2017, section 4.1, restated from the paper) -> quantised 80-byte nodes.
// Replaces the driver BLAS/TLAS build of the reference (src/merian-shaders/scene/scene.cpp:1672-1892), which runs on the device per
// geometry change; here it runs on the host cores, in parallel (OpenMP tasks over subtrees, level-synchronous emission); per-phase
// times in profiles/r02_bvh_build.md. The output is a deterministic function of the input boxes (thread timing only permutes the
// numbering of the intermediate BVH2 nodes).
#include "bvh8.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <numeric>
#include <omp.h>
#include <vector>

#if defined(__SSE2__)
#include <emmintrin.h>
#define MPT_BVH_SSE 1
#endif

namespace mpt {
namespace {

struct B3 {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    void grow(const Box& b) {
        for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); }
    }
    void grow(const B3& b) { // growing by an empty box (lo = FLT_MAX, hi = -FLT_MAX) is a no-op
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

struct Box6 { // a B3 without constructors, so that arrays of nodes need no initialisation pass
    float lo[3], hi[3];
    float half_area() const {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0) return 0.f;
        return dx * dy + dy * dz + dz * dx;
    }
};
struct Node2 { // trivially constructible: the 2n-node array is allocated uninitialised and first touched by the thread that builds the node
    Box6 box;
    uint32_t left, right;    // internal children
    uint32_t count;          // > 0: BVH2 leaf (one primitive, or up to 3 with coincident centroids)
    uint32_t rfirst, rcount; // range of every node (leaf or internal) in the prim index array
};
template <class T> struct RawArray { // uninitialised storage for trivially constructible T
    T* p = nullptr;
    size_t n = 0;
    ~RawArray() { std::free(p); }
    bool alloc(size_t count) {
        std::free(p);
        n = count;
        p = static_cast<T*>(std::malloc(std::max<size_t>(count, 1) * sizeof(T)));
        return p != nullptr;
    }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
    T* data() { return p; }
};

constexpr int NBINS = 16;
constexpr uint32_t kSmallNode = 8;              // at most this many primitives: exact sweep over sorted centroids instead of binning
constexpr uint32_t kTaskGrain = 4096;           // subtrees below this size are built by the thread that reaches them
constexpr uint32_t kParallelBinning = 1u << 17; // nodes of at least this size are split by all threads together (Builder::split_top)
constexpr uint32_t kLocalSubtree = 16384;       // subtrees of at most this many primitives are built from private contiguous copies
constexpr uint32_t kMaxRecursion = 256;         // SAH levels before the builder switches to median splits (bounds the recursion depth)

// Builder knobs for tuning sessions (host code only; defaults are the measured optimum).
float env_float(const char* name, float def) {
    const char* v = std::getenv(name);
    return v ? float(std::atof(v)) : def;
}

// BVH2 (binned SAH, 16 bins, the three axes binned in ONE pass; exact sweep for nodes of <= 8 primitives) with the dynamic program of the
// wide collapse evaluated in post-order inside the same task tree: cost[n][i-1] = C(n, i), the cheapest way to represent BVH2 subtree n
// as at most i BVH8 children (leaves of <= 3 primitives, or internal nodes).
// Parallel: every split of a large node spawns an OpenMP task for one child; nodes come from a pre-sized array through an atomic
// counter. A job carries its box and centroid box (accumulated from the parent's bins), so a node costs one binning pass and one
// partition pass.
struct Builder {
    const Box* boxes = nullptr;
    RawArray<float> cent; // 3 per prim
    RawArray<uint32_t> idx;
    RawArray<Node2> nodes;
    RawArray<float> cost;         // 7 per node
    RawArray<uint8_t> dec, dist8; // 7 per node; 1 per node
    uint32_t node_count = 0;
    std::atomic<uint32_t> next_node{1};
    uint32_t sah_depth_limit = kMaxRecursion;
    float balance = 0.f; // SAH candidates must leave at least this fraction of the primitives on the smaller side (0 = plain SAH)
    float c_prim = 0.5f, c_node = 1.0f;

    // Node ids are handed out in blocks of kIdBlock per thread: with one shared counter the nodes that neighbour each other in the arrays
    // (44 B of Node2, 28 B of costs, 7 + 1 B of decisions per node) were being written by different threads at the same time, and the
    // cache lines they share bounced between the cores (8 threads: 2.3x over one thread; profiles/r02_bvh_build.md).
    static constexpr uint32_t kIdBlock = 256;
    struct alignas(64) IdBlock { uint32_t next = 0, end = 0; };
    std::vector<IdBlock> id_blocks;
    uint32_t alloc_pair() {
        IdBlock& b = id_blocks[size_t(omp_get_thread_num())];
        if (b.next + 2 > b.end) {
            b.next = next_node.fetch_add(kIdBlock);
            b.end = b.next + kIdBlock;
        }
        const uint32_t l = b.next;
        b.next += 2;
        return l;
    }
    struct Job { uint32_t node, first, count, depth; B3 box, cbox; };
    // The arrays a subtree is built from: the global ones, or — once a subtree is small enough to live in L2 — private contiguous copies
    // of its primitives' boxes and centroids (the global index array is a random permutation by then, and every level of the tree would
    // gather 36 bytes per primitive from all over memory again). `first0`: position in the global index array of local position 0.
    struct View { const Box* boxes; const float* cent; uint32_t* idx; uint32_t first0; };
    struct Bins {
        B3 bb[3][NBINS], cb[3][NBINS];
        uint32_t bc[3][NBINS];
        Bins() { std::memset(bc, 0, sizeof(bc)); }
        void merge(const Bins& o) {
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < NBINS; b++) {
                    if (!o.bc[a][b]) continue;
                    bb[a][b].grow(o.bb[a][b]);
                    cb[a][b].grow(o.cb[a][b]);
                    bc[a][b] += o.bc[a][b];
                }
        }
    };

    void build(uint32_t n) {
        const auto t0 = std::chrono::steady_clock::now();
        id_blocks.assign(size_t(std::max(1, omp_get_max_threads())), IdBlock());
        const size_t nn = size_t(n) * 2 + 1 + id_blocks.size() * kIdBlock;
        if (!idx.alloc(n) || !cent.alloc(size_t(n) * 3 + 1) || !nodes.alloc(nn) || !cost.alloc(nn * 7) || !dec.alloc(nn * 7) || !dist8.alloc(nn))
            throw std::bad_alloc();
        B3 box, cbox;
#pragma omp parallel
        {
            B3 lb, lc;
#pragma omp for schedule(static) nowait
            for (int64_t i = 0; i < int64_t(n); i++) {
                idx[size_t(i)] = uint32_t(i);
                float c[3];
                for (int a = 0; a < 3; a++) c[a] = cent[size_t(i) * 3 + a] = 0.5f * (boxes[i].lo[a] + boxes[i].hi[a]);
                lb.grow(boxes[i]);
                lc.grow_pt(c);
            }
#pragma omp critical
            {
                box.grow(lb);
                cbox.grow(lc);
            }
        }
        if (std::getenv("MPT_BUILD_TIMING"))
            std::fprintf(stderr, "[bvh8_build] alloc + centroids %.1f ms; ", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        const View global{boxes, cent.data(), idx.data(), 0u};
        // The top of the tree (nodes of >= kParallelBinning primitives: 7 nodes for 1 M primitives) is split breadth-first with every
        // thread working on ONE node at a time (parallel binning, parallel stable partition). As tasks, these nodes were the critical
        // path: a thread waiting for its binning tasks picks up a whole subtree task in the meantime and returns to the split of the
        // large node only when that is built (the 500 k-primitive node took 227 ms to bin, profiles/r02_bvh_build.md). The subtrees
        // below become tasks; the dynamic program of the top nodes runs afterwards, children before parents.
        std::vector<Job> work{Job{0, 0, n, 0, box, cbox}}, subtrees;
        std::vector<uint32_t> top_nodes;
        for (size_t head = 0; head < work.size(); head++) {
            const Job j = work[head];
            Job lj, rj;
            if (j.count >= kParallelBinning && j.depth < sah_depth_limit && split_top(j, global, lj, rj)) {
                top_nodes.push_back(j.node);
                work.push_back(lj);
                work.push_back(rj);
            } else {
                subtrees.push_back(j);
            }
        }
        if (std::getenv("MPT_BUILD_TIMING"))
            std::fprintf(stderr, "top %zu nodes done at %.1f ms; ", top_nodes.size(), std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
#pragma omp parallel
#pragma omp single
        for (size_t i = 0; i < subtrees.size(); i++) {
            const Job sj = subtrees[i];
#pragma omp task default(shared) firstprivate(sj)
            build_node(sj, global);
        }
        for (size_t i = top_nodes.size(); i-- > 0;) collapse_cost(top_nodes[i]);
        node_count = next_node.load();
    }

    // The binned SAH split of build_subtree for one LARGE node, all threads on it. Returns false when the bins do not separate the
    // primitives (the caller then hands the node to build_subtree, which knows the fallbacks).
    bool split_top(const Job& j, const View& V, Job& lj, Job& rj) {
        float k[3];
        bool axis_ok[3];
        for (int a = 0; a < 3; a++) {
            axis_ok[a] = j.cbox.hi[a] > j.cbox.lo[a];
            k[a] = axis_ok[a] ? float(NBINS) * (1.0f - 1e-6f) / (j.cbox.hi[a] - j.cbox.lo[a]) : 0.f;
        }
        constexpr int kParts = 64; // fixed: the result does not depend on the number of threads
        std::vector<Bins> part(kParts);
        auto part_first = [&](int t) { return j.first + uint32_t(uint64_t(j.count) * uint64_t(t) / kParts); };
#pragma omp parallel for schedule(dynamic, 1)
        for (int t = 0; t < kParts; t++) bin_range(j, V, part_first(t), part_first(t + 1), k, axis_ok, part[size_t(t)]);
        Bins B;
        for (int t = 0; t < kParts; t++) B.merge(part[size_t(t)]);
        int best_axis = -1, best_bin = -1;
        best_split(j, B, axis_ok, best_axis, best_bin);
        if (best_axis < 0) return false;
        const int a = best_axis;
        const float cmin = j.cbox.lo[a], ka = k[a];
        const float* cent = V.cent;
        uint32_t* const idx = V.idx;
        auto goes_left = [&](uint32_t p) {
            int b = int((cent[size_t(p) * 3 + a] - cmin) * ka);
            b = std::min(std::max(b, 0), NBINS - 1);
            return b <= best_bin;
        };
        // stable partition in kParts chunks: count, prefix sum, scatter into a scratch array, copy back
        uint32_t nl[kParts + 1];
        nl[kParts] = 0;
#pragma omp parallel for schedule(static)
        for (int t = 0; t < kParts; t++) {
            uint32_t c = 0;
            for (uint32_t i = part_first(t); i < part_first(t + 1); i++) c += goes_left(idx[i]) ? 1u : 0u;
            nl[t] = c;
        }
        uint32_t lbase[kParts], rbase[kParts], mid = 0;
        for (int t = 0; t < kParts; t++) { lbase[t] = mid; mid += nl[t]; }
        if (mid == 0 || mid == j.count) return false;
        uint32_t rcur = mid;
        for (int t = 0; t < kParts; t++) { rbase[t] = rcur; rcur += (part_first(t + 1) - part_first(t)) - nl[t]; }
        RawArray<uint32_t> tmp;
        if (!tmp.alloc(j.count)) throw std::bad_alloc();
#pragma omp parallel for schedule(static)
        for (int t = 0; t < kParts; t++) {
            uint32_t l = lbase[t], r = rbase[t];
            for (uint32_t i = part_first(t); i < part_first(t + 1); i++) {
                const uint32_t p = idx[i];
                if (goes_left(p)) tmp[l++] = p; else tmp[r++] = p;
            }
        }
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < int64_t(j.count); i++) idx[j.first + size_t(i)] = tmp[size_t(i)];

        Node2& nd = nodes[j.node];
        set_box(nd, j.box);
        nd.rfirst = V.first0 + j.first;
        nd.rcount = j.count;
        nd.count = 0;
        const uint32_t l = alloc_pair();
        nd.left = l;
        nd.right = l + 1;
        lj = Job{l, j.first, mid, j.depth + 1, B3(), B3()};
        rj = Job{l + 1, j.first + mid, j.count - mid, j.depth + 1, B3(), B3()};
        for (int b = 0; b < NBINS; b++) { // the children's boxes are unions of the bins on either side of the split
            Job& t = b <= best_bin ? lj : rj;
            t.box.grow(B.bb[a][b]);
            t.cbox.grow(B.cb[a][b]);
        }
        return true;
    }

    // the cheapest bin boundary over the three axes (SAH: area x count on either side); best_axis stays -1 if no boundary separates
    void best_split(const Job& j, const Bins& B, const bool* axis_ok, int& best_axis, int& best_bin) const {
        float best_cost = FLT_MAX;
        for (int a = 0; a < 3; a++) {
            if (!axis_ok[a]) continue;
            float ra[NBINS];
            uint32_t rc[NBINS];
            B3 acc;
            uint32_t cnt = 0;
            for (int b = NBINS - 1; b > 0; b--) {
                acc.grow(B.bb[a][b]);
                cnt += B.bc[a][b];
                ra[b] = acc.half_area();
                rc[b] = cnt;
            }
            acc = B3();
            cnt = 0;
            for (int b = 0; b < NBINS - 1; b++) {
                acc.grow(B.bb[a][b]);
                cnt += B.bc[a][b];
                if (cnt == 0 || rc[b + 1] == 0) continue;
                if (float(std::min(cnt, rc[b + 1])) < balance * float(j.count)) continue;
                const float c = acc.half_area() * float(cnt) + ra[b + 1] * float(rc[b + 1]);
                if (c < best_cost) { best_cost = c; best_axis = a; best_bin = b; }
            }
        }
    }

    void bin_range(const Job& j, const View& V, uint32_t first, uint32_t last, const float* k, const bool* axis_ok, Bins& B) const {
        const Box* boxes = V.boxes;
        const float* cent = V.cent;
        const uint32_t* idx = V.idx;
#ifdef MPT_BVH_SSE
        // The same arithmetic four lanes at a time: a box and a centroid are one vector each (lane 3 is padding and never read back), so a
        // bin update is four min / max instead of twelve scalar ones. Bin index: the scalar form's subtract, multiply and truncation, with
        // the clamp to [0, NBINS - 1] applied before the conversion (same result: truncation is towards zero). Centroids are read 16 bytes
        // at a time from a 12-byte stride, hence the one float of padding behind every centroid array.
        struct alignas(16) VB { __m128 blo, bhi, clo, chi; };
        VB vb[3][NBINS];
        const __m128 big = _mm_set1_ps(FLT_MAX), small = _mm_set1_ps(-FLT_MAX);
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < NBINS; b++) vb[a][b] = VB{big, small, big, small};
        const __m128 cmin = _mm_setr_ps(j.cbox.lo[0], j.cbox.lo[1], j.cbox.lo[2], 0.f), kk = _mm_setr_ps(k[0], k[1], k[2], 0.f);
        const __m128 zero = _mm_setzero_ps(), top = _mm_set1_ps(float(NBINS) - 0.5f);
        for (uint32_t i = first; i < last; i++) {
            const uint32_t p = idx[i];
            const __m128 c = _mm_loadu_ps(&cent[size_t(p) * 3]);
            const __m128 lo = _mm_loadu_ps(&boxes[p].lo[0]);             // lo.x lo.y lo.z hi.x
            const __m128 t = _mm_loadu_ps(&boxes[p].lo[2]);              // lo.z hi.x hi.y hi.z
            const __m128 hi = _mm_shuffle_ps(t, t, _MM_SHUFFLE(3, 3, 2, 1)); // hi.x hi.y hi.z hi.z
            const __m128 f = _mm_min_ps(_mm_max_ps(_mm_mul_ps(_mm_sub_ps(c, cmin), kk), zero), top); // max(NaN, 0) = 0, as int(NaN) clamps
            alignas(16) int bi[4];
            _mm_store_si128(reinterpret_cast<__m128i*>(bi), _mm_cvttps_epi32(f));
            for (int a = 0; a < 3; a++) {
                if (!axis_ok[a]) continue;
                VB& v = vb[a][bi[a]];
                B.bc[a][bi[a]]++;
                v.blo = _mm_min_ps(v.blo, lo);
                v.bhi = _mm_max_ps(v.bhi, hi);
                v.clo = _mm_min_ps(v.clo, c);
                v.chi = _mm_max_ps(v.chi, c);
            }
        }
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < NBINS; b++) {
                if (!B.bc[a][b]) continue;
                alignas(16) float q[4][4];
                _mm_store_ps(q[0], vb[a][b].blo); _mm_store_ps(q[1], vb[a][b].bhi); _mm_store_ps(q[2], vb[a][b].clo); _mm_store_ps(q[3], vb[a][b].chi);
                for (int d = 0; d < 3; d++) {
                    B.bb[a][b].lo[d] = q[0][d]; B.bb[a][b].hi[d] = q[1][d];
                    B.cb[a][b].lo[d] = q[2][d]; B.cb[a][b].hi[d] = q[3][d];
                }
            }
#else
        for (uint32_t i = first; i < last; i++) {
            const uint32_t p = idx[i];
            const float* c = &cent[size_t(p) * 3];
            for (int a = 0; a < 3; a++) {
                if (!axis_ok[a]) continue;
                int b = int((c[a] - j.cbox.lo[a]) * k[a]);
                b = std::min(std::max(b, 0), NBINS - 1);
                B.bc[a][b]++;
                B.bb[a][b].grow(boxes[p]);
                B.cb[a][b].grow_pt(c);
            }
        }
#endif
    }

    // the dynamic program for one node whose children are done
    void collapse_cost(uint32_t k) {
        const Node2& nd = nodes[k];
        const float area = nd.box.half_area();
        float* C = &cost[size_t(k) * 7];
        uint8_t* D = &dec[size_t(k) * 7];
        if (nd.count > 0) { // BVH2 leaf
            for (int i = 0; i < 7; i++) { C[i] = area * float(nd.count) * c_prim; D[i] = 0; }
            return;
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

    static void set_box(Node2& nd, const B3& b) {
        for (int a = 0; a < 3; a++) { nd.box.lo[a] = b.lo[a]; nd.box.hi[a] = b.hi[a]; }
    }
    void make_leaf(uint32_t node, uint32_t count) {
        nodes[node].count = count;
        collapse_cost(node);
    }

    // SAH split of a small node: exact sweep over the primitives sorted by centroid, per axis. Returns false if no axis separates them.
    bool split_small(const Job& j, const View& V, uint32_t& mid, B3& lbox, B3& lcb, B3& rbox, B3& rcb) {
        const Box* boxes = V.boxes;
        const float* cent = V.cent;
        uint32_t* idx = V.idx;
        const uint32_t n = j.count;
        if (n == 2) { // one possible split; keep the index order when the centroids allow it
            const uint32_t p = idx[j.first], q = idx[j.first + 1];
            bool differ = false;
            for (int a = 0; a < 3; a++) differ |= cent[size_t(p) * 3 + a] != cent[size_t(q) * 3 + a];
            if (!differ) return false;
            lbox.grow(boxes[p]); lcb.grow_pt(&cent[size_t(p) * 3]);
            rbox.grow(boxes[q]); rcb.grow_pt(&cent[size_t(q) * 3]);
            mid = 1;
            return true;
        }
        uint32_t order[kSmallNode], best_order[kSmallNode];
        float best_cost = FLT_MAX;
        uint32_t best_mid = 0;
        for (int a = 0; a < 3; a++) {
            if (!(j.cbox.hi[a] > j.cbox.lo[a])) continue;
            for (uint32_t i = 0; i < n; i++) order[i] = idx[j.first + i];
            std::sort(order, order + n, [&](uint32_t p, uint32_t q) {
                const float cp = cent[size_t(p) * 3 + a], cq = cent[size_t(q) * 3 + a];
                return cp < cq || (cp == cq && p < q);
            });
            float ra[kSmallNode];
            B3 acc;
            for (uint32_t i = n - 1; i > 0; i--) { acc.grow(boxes[order[i]]); ra[i] = acc.half_area(); }
            acc = B3();
            for (uint32_t i = 0; i + 1 < n; i++) {
                acc.grow(boxes[order[i]]);
                if (float(std::min(i + 1, n - i - 1)) < balance * float(n)) continue;
                const float c = acc.half_area() * float(i + 1) + ra[i + 1] * float(n - i - 1);
                if (c < best_cost) { best_cost = c; best_mid = i + 1; std::memcpy(best_order, order, n * sizeof(uint32_t)); }
            }
        }
        if (best_mid == 0) return false;
        for (uint32_t i = 0; i < n; i++) {
            idx[j.first + i] = best_order[i];
            B3& bx = i < best_mid ? lbox : rbox;
            B3& cb = i < best_mid ? lcb : rcb;
            bx.grow(boxes[best_order[i]]);
            cb.grow_pt(&cent[size_t(best_order[i]) * 3]);
        }
        mid = best_mid;
        return true;
    }

    // j.first is a position in V.idx; the node records the position in the GLOBAL index array
    void build_node(Job j, const View& V) {
        Node2& nd = nodes[j.node];
        set_box(nd, j.box);
        nd.left = nd.right = 0;
        nd.rfirst = V.first0 + j.first;
        nd.rcount = j.count;
        if (j.count <= 1) { make_leaf(j.node, j.count); return; }
        if (V.first0 == 0 && V.idx == idx.data() && j.count <= kLocalSubtree && j.count > kSmallNode) {
            // small enough for L2: build this subtree from private contiguous copies, then write the resulting order back
            std::vector<Box> lb(j.count);
            std::vector<float> lc(size_t(j.count) * 3 + 1); // + 1: bin_range reads centroids 16 bytes at a time
            std::vector<uint32_t> li(j.count), lg(j.count);
            for (uint32_t i = 0; i < j.count; i++) {
                const uint32_t p = idx[j.first + i];
                lg[i] = p;
                li[i] = i;
                lb[i] = boxes[p];
                for (int a = 0; a < 3; a++) lc[size_t(i) * 3 + a] = cent[size_t(p) * 3 + a];
            }
            const View local{lb.data(), lc.data(), li.data(), j.first};
            Job lj = j;
            lj.first = 0;
            build_subtree(lj, local);
            for (uint32_t i = 0; i < j.count; i++) idx[j.first + i] = lg[li[i]];
            return;
        }
        build_subtree(j, V);
    }

    void build_subtree(Job j, const View& V) {
        Node2& nd = nodes[j.node];
        const Box* boxes = V.boxes;
        const float* cent = V.cent;
        uint32_t* const idx = V.idx;
        set_box(nd, j.box);
        nd.left = nd.right = 0;
        nd.rfirst = V.first0 + j.first;
        nd.rcount = j.count;
        if (j.count <= 1) { make_leaf(j.node, j.count); return; }

        uint32_t mid = 0;
        Job lj{0, 0, 0, j.depth + 1, B3(), B3()}, rj = lj;
        bool have_child_boxes = false;
        if (j.depth >= sah_depth_limit) {
            // a degenerate SAH chain (one primitive peeled off per level) would outgrow the traversal stack: below this depth
            // split at the object median of the widest centroid axis, which bounds the remaining depth by log2(count)
            int a = 0;
            for (int q = 1; q < 3; q++)
                if (j.cbox.hi[q] - j.cbox.lo[q] > j.cbox.hi[a] - j.cbox.lo[a]) a = q;
            mid = j.count / 2;
            std::nth_element(idx + j.first, idx + j.first + mid, idx + j.first + j.count, [&](uint32_t p, uint32_t q) {
                const float cp = cent[size_t(p) * 3 + a], cq = cent[size_t(q) * 3 + a];
                return cp < cq || (cp == cq && p < q);
            });
        } else if (j.count <= kSmallNode) {
            have_child_boxes = split_small(j, V, mid, lj.box, lj.cbox, rj.box, rj.cbox);
        } else {
            // binned SAH over the 3 axes, one pass (several tasks for the largest nodes)
            float k[3];
            bool axis_ok[3];
            for (int a = 0; a < 3; a++) {
                axis_ok[a] = j.cbox.hi[a] > j.cbox.lo[a];
                k[a] = axis_ok[a] ? float(NBINS) * (1.0f - 1e-6f) / (j.cbox.hi[a] - j.cbox.lo[a]) : 0.f;
            }
            Bins B;
            bin_range(j, V, j.first, j.first + j.count, k, axis_ok, B);
            int best_axis = -1, best_bin = -1;
            best_split(j, B, axis_ok, best_axis, best_bin);
            if (best_axis >= 0) {
                const int a = best_axis;
                const float cmin = j.cbox.lo[a], ka = k[a];
                auto it = std::partition(idx + j.first, idx + j.first + j.count, [&](uint32_t p) {
                    int b = int((cent[size_t(p) * 3 + a] - cmin) * ka);
                    b = std::min(std::max(b, 0), NBINS - 1);
                    return b <= best_bin;
                });
                mid = uint32_t(it - idx) - j.first;
                for (int b = 0; b < NBINS; b++) { // the children's boxes are unions of the bins on either side of the split
                    Job& t = b <= best_bin ? lj : rj;
                    t.box.grow(B.bb[a][b]);
                    t.cbox.grow(B.cb[a][b]);
                }
                have_child_boxes = true;
            }
        }
        if (mid == 0 || mid == j.count) { // all centroids coincide: one BVH2 leaf of up to 3 primitives, else split the index range in half
            if (j.count <= 3u) { make_leaf(j.node, j.count); return; }
            mid = j.count / 2;
            have_child_boxes = false;
        }
        const uint32_t l = alloc_pair();
        nd.left = l;
        nd.right = l + 1;
        nd.count = 0;
        lj.node = l; lj.first = j.first; lj.count = mid;
        rj.node = l + 1; rj.first = j.first + mid; rj.count = j.count - mid;
        if (!have_child_boxes) {
            lj.box = lj.cbox = rj.box = rj.cbox = B3();
            for (Job* t : {&lj, &rj})
                for (uint32_t i = t->first; i < t->first + t->count; i++) {
                    t->box.grow(boxes[idx[i]]);
                    t->cbox.grow_pt(&cent[size_t(idx[i]) * 3]);
                }
        }
        if (rj.count > kTaskGrain && lj.count > kTaskGrain) {
#pragma omp task default(shared) firstprivate(rj)
            build_node(rj, V);
            build_node(lj, V);
#pragma omp taskwait
        } else {
            build_node(lj, V);
            build_node(rj, V);
        }
        collapse_cost(j.node);
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

// One wide node under construction: its (up to 8) children as BVH2 subtrees, slot-assigned and quantised.
struct Wide {
    uint32_t root2 = 0; // BVH2 subtree this wide node stands for
    uint32_t child[8];  // BVH2 node per slot (valid where slot_used)
    uint8_t slot_used = 0, slot_leaf = 0;
    uint32_t n_internal = 0, n_prims = 0;
    Bvh8Node node;
};

} // namespace

static void build_bvh8_once(const Box* boxes, uint32_t n, Bvh8& out, uint32_t sah_depth_limit, float c_prim) {
    out.nodes.clear();
    out.prim_order.clear();
    out.depth = 1;
    if (n == 0) {
        for (int a = 0; a < 3; a++) out.bounds.lo[a] = out.bounds.hi[a] = 0.f;
        Bvh8Node root;
        std::memset(&root, 0, sizeof(root));
        root.ex = root.ey = root.ez = 1;
        out.nodes.push_back(root);
        return;
    }
    Builder b2;
    b2.boxes = boxes;
    b2.sah_depth_limit = std::min(sah_depth_limit, kMaxRecursion);
    b2.balance = env_float("MPT_BVH_BALANCE", 0.f);
    b2.c_prim = c_prim >= 0.f ? c_prim : env_float("MPT_BVH_CPRIM", 0.5f); // triangles: measured optimum on C3 (0.15 ... 2.0 within 0.6 %)
    const bool timing = std::getenv("MPT_BUILD_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
    auto t_phase = now();
    b2.build(n);
    if (timing) { std::fprintf(stderr, "[bvh8_build] n=%u bvh2 + collapse dp %.1f ms", n, ms_since(t_phase)); t_phase = now(); }
    const RawArray<Node2>& n2 = b2.nodes;
    const RawArray<uint8_t>& dec = b2.dec;
    for (int a = 0; a < 3; a++) { out.bounds.lo[a] = n2[0].box.lo[a]; out.bounds.hi[a] = n2[0].box.hi[a]; }

    // ---- emission, one level of wide nodes at a time: the nodes of a level are independent (children, slot assignment, quantisation in
    // parallel); a prefix sum over the level then gives every node its child_base / prim_base in breadth-first order.
    std::vector<Wide> level(1), next;
    level[0].root2 = 0;
    uint32_t depth = 0;
    while (!level.empty()) {
        depth++;
        const size_t base_index = out.nodes.size();
#pragma omp parallel for schedule(dynamic, 64) if (level.size() > 256)
        for (int64_t li = 0; li < int64_t(level.size()); li++) {
            Wide& w = level[size_t(li)];
            // ---- gather up to 8 children along the decisions of the dynamic program (Ylitie, Karras, Laine 2017, section 4.1)
            uint32_t ch[8];
            bool ch_leaf[8];
            int nch = 0;
            if (n2[w.root2].count > 0) { // the BVH2 root itself is a leaf
                ch_leaf[nch] = true;
                ch[nch++] = w.root2;
            } else {
                struct Item { uint32_t node; int slots; };
                Item work[16];
                int nw = 0;
                const int k8 = b2.dist8[w.root2];
                work[nw++] = {n2[w.root2].right, 8 - k8};
                work[nw++] = {n2[w.root2].left, k8};
                while (nw) {
                    Item it = work[--nw];
                    if (n2[it.node].count > 0) { ch_leaf[nch] = true; ch[nch++] = it.node; continue; }
                    while (it.slots > 1 && dec[size_t(it.node) * 7 + size_t(it.slots - 1)] == 0) it.slots--;
                    const uint8_t d = dec[size_t(it.node) * 7 + size_t(it.slots - 1)];
                    if (it.slots == 1) { ch_leaf[nch] = (d == 0); ch[nch++] = it.node; continue; }
                    work[nw++] = {n2[it.node].right, it.slots - int(d)};
                    work[nw++] = {n2[it.node].left, int(d)};
                }
            }
            // ---- slot assignment: slot s is visited first by rays with octant s (direction sign bit set = negative),
            // so it should hold the child that is nearest along d_s: minimise sum of dot(centroid_c - centroid_p, d_s).
            const Box6& pb = n2[w.root2].box;
            float pc[3];
            for (int a = 0; a < 3; a++) pc[a] = 0.5f * (pb.lo[a] + pb.hi[a]);
            float cst[8][8];
            for (int c = 0; c < nch; c++) {
                const Box6& cb = n2[ch[c]].box;
                float d[3];
                for (int a = 0; a < 3; a++) d[a] = 0.5f * (cb.lo[a] + cb.hi[a]) - pc[a];
                for (int s = 0; s < 8; s++) cst[c][s] = ((s & 1) ? -d[0] : d[0]) + ((s & 2) ? -d[1] : d[1]) + ((s & 4) ? -d[2] : d[2]);
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
                        if (cst[c][s] < bc) { bc = cst[c][s]; bi = c; bs = s; }
                    }
                }
                slot_of[bi] = bs;
                child_in_slot[bs] = bi;
            }
            // ---- encode (child_base / prim_base are filled in after the prefix sum)
            Bvh8Node& node = w.node;
            std::memset(&node, 0, sizeof(node));
            node.px = pb.lo[0]; node.py = pb.lo[1]; node.pz = pb.lo[2];
            node.ex = exponent_for(pb.hi[0] - pb.lo[0]);
            node.ey = exponent_for(pb.hi[1] - pb.lo[1]);
            node.ez = exponent_for(pb.hi[2] - pb.lo[2]);
            const double sc[3] = {std::ldexp(1.0, int(node.ex) - 127), std::ldexp(1.0, int(node.ey) - 127), std::ldexp(1.0, int(node.ez) - 127)};
            uint8_t* qlo[3] = {node.qlox, node.qloy, node.qloz};
            uint8_t* qhi[3] = {node.qhix, node.qhiy, node.qhiz};
            w.slot_used = w.slot_leaf = 0;
            w.n_internal = w.n_prims = 0;
            for (int s = 0; s < 8; s++) {
                const int c = child_in_slot[s];
                if (c < 0) continue;
                const Node2& cn = n2[ch[c]];
                w.child[s] = ch[c];
                w.slot_used |= uint8_t(1u << s);
                for (int a = 0; a < 3; a++) {
                    const double lo = std::floor((double(cn.box.lo[a]) - double(pb.lo[a])) / sc[a]);
                    const double hi = std::ceil((double(cn.box.hi[a]) - double(pb.lo[a])) / sc[a]);
                    qlo[a][s] = uint8_t(std::min(std::max(lo, 0.0), 255.0));
                    qhi[a][s] = uint8_t(std::min(std::max(hi, 0.0), 255.0));
                }
                if (!ch_leaf[c]) {
                    node.imask |= uint8_t(1u << s);
                    w.n_internal++;
                } else {
                    const uint32_t unary = (1u << cn.rcount) - 1u; // 1, 3, 7
                    node.leaf_bits |= unary << (3 * s);
                    w.slot_leaf |= uint8_t(1u << s);
                    w.n_prims += cn.rcount;
                }
            }
        }
        // ---- prefix sums: breadth-first numbering of the next level and of the primitive records
        const size_t next_base = base_index + level.size();
        size_t child_cursor = next_base, prim_cursor = out.prim_order.size();
        std::vector<size_t> next_first(level.size());
        for (size_t li = 0; li < level.size(); li++) {
            Wide& w = level[li];
            w.node.child_base = uint32_t(child_cursor);
            w.node.prim_base = uint32_t(prim_cursor);
            next_first[li] = child_cursor - next_base;
            child_cursor += w.n_internal;
            prim_cursor += w.n_prims;
        }
        out.nodes.resize(next_base);
        out.prim_order.resize(prim_cursor);
        next.assign(child_cursor - next_base, Wide());
#pragma omp parallel for schedule(static) if (level.size() > 256)
        for (int64_t li = 0; li < int64_t(level.size()); li++) {
            const Wide& w = level[size_t(li)];
            out.nodes[base_index + size_t(li)] = w.node;
            size_t nf = next_first[size_t(li)], pf = w.node.prim_base;
            for (int s = 0; s < 8; s++) {
                if (!(w.slot_used & (1u << s))) continue;
                if (w.slot_leaf & (1u << s)) {
                    const Node2& cn = n2[w.child[s]];
                    for (uint32_t k = 0; k < cn.rcount; k++) out.prim_order[pf++] = b2.idx[cn.rfirst + k];
                } else {
                    next[nf++].root2 = w.child[s];
                }
            }
        }
        level.swap(next);
    }
    out.depth = depth;
    if (timing) std::fprintf(stderr, ", emit %.1f ms (%zu nodes, %u levels)\n", ms_since(t_phase), out.nodes.size(), depth);
}

// The traversal keeps a bounded number of stack entries per wide level (bvh8_traverse.cuh), so the depth must be bounded. SAH trees of
// real scenes are 8-10 wide levels deep (1 M-triangle terrain: 8); a degenerate input (primitives of geometrically growing size on a
// line) makes SAH peel one primitive per level. When that happens the tree is rebuilt with SAH only for the top levels and object-median
// splits below, which bounds the depth by about log8(n) + the SAH levels.
void build_bvh8(const Box* boxes, uint32_t n, Bvh8& out, int /*max_leaf*/, float c_prim) {
    if (std::getenv("MPT_BVH_SAH_DEPTH")) { // tuning sessions: SAH for the top levels only, object-median splits below
        build_bvh8_once(boxes, n, out, uint32_t(env_float("MPT_BVH_SAH_DEPTH", 64.f)), c_prim);
        if (out.depth <= kMaxBvh8Depth) return;
    }
    for (uint32_t limit : {0xFFFFFFFFu, 16u, 8u, 0u}) {
        build_bvh8_once(boxes, n, out, limit, c_prim);
        if (out.depth <= kMaxBvh8Depth) return;
    }
}

} // namespace mpt
