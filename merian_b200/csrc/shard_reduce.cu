// shard_reduce.cu — the one exchange step of the path: combining the per-GPU accumulation buffers of a sample-sharded frame
// (include/mpt.h "multi-GPU"). NCCL over NVLink / NVSwitch, inside libmpt.so so that a C++ merian node can shard a
// frame without Python: `mpt_comm_*` wraps one ncclComm_t + a side stream, `mpt_reduce_begin` issues
//     ncclReduce(sum of the partial sums -> root)  ->  divide by the total sample count  ->  optional copy to a pinned host buffer
// on the side stream behind the render that produced the partial sums, so that the next frame's render overlaps it;
// `mpt_reduce_end` / `mpt_reduce_wait` join it. Up to four reduces may be in flight (slots): a host that blocks on frame i - 1 before it
// enqueues frame i + 1 keeps only one frame queued on the GPU, and every host hiccup then stalls the device (measured on 2 GPUs: 16-33 ms
// spikes, profiles/r02j_reduce_modes.jsonl); with slots the render stream itself waits for the reduce that last read a sum buffer. The reference has no multi-device path (render_pt.slang:113-122 loops over the samples
// of a pixel in one thread); the partition by sample index is ours, the estimator is unchanged.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2): in a process that already carries a copy (PyTorch bundles its own) that copy
// is used — two different NCCL builds in one process would clash on symbol names — otherwise the system library is loaded.
#include "../../include/mpt.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

namespace {

// the subset of nccl.h this file needs (ABI-stable since NCCL 2.0)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclFloat = 7, kNcclSum = 0;

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string error;
};

NcclApi g_nccl;
std::once_flag g_nccl_once;
std::string g_comm_error;

void load_nccl() {
    const char* override_path = getenv("MPT_NCCL_LIBRARY");
    const char* names[] = {override_path, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n) continue;
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) {
        g_nccl.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found");
        return;
    }
#define MPT_SYM(field, name)                                                                 \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.lib, name));          \
    if (!g_nccl.field) { g_nccl.error = std::string("libnccl lacks ") + name; return; }
    MPT_SYM(GetUniqueId, "ncclGetUniqueId")
    MPT_SYM(CommInitRank, "ncclCommInitRank")
    MPT_SYM(CommDestroy, "ncclCommDestroy")
    MPT_SYM(Reduce, "ncclReduce")
    MPT_SYM(AllReduce, "ncclAllReduce")
    MPT_SYM(GetErrorString, "ncclGetErrorString")
    MPT_SYM(GetVersion, "ncclGetVersion")
#undef MPT_SYM
}

const NcclApi* nccl() {
    std::call_once(g_nccl_once, load_nccl);
    return g_nccl.error.empty() ? &g_nccl : nullptr;
}

// out = sum / spp_total, the same division k_resolve applies to an unsharded frame (render_pt.slang:125-129)
__global__ void k_normalise(const float4* __restrict__ sum, float4* __restrict__ out, uint32_t npix, float spp_total) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const float4 s = sum[i];
    out[i] = make_float4(s.x / spp_total, s.y / spp_total, s.z / spp_total, s.w / spp_total);
}

} // namespace

struct mpt_comm {
    int device = 0, rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    bool owns_comm = false;
    cudaStream_t side = nullptr;
    cudaEvent_t ready = nullptr;
    static constexpr uint32_t kSlots = 4; // reduces in flight (the side stream runs them in order)
    cudaEvent_t done[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    bool pending[kSlots] = {false, false, false, false};
    std::string error;
};

namespace {
mpt_status cfail(mpt_comm* c, mpt_status s, const std::string& msg) {
    if (c) c->error = msg;
    else g_comm_error = msg;
    return s;
}
#define CCU(call)                                                                                                  \
    do {                                                                                                           \
        const cudaError_t e__ = (call);                                                                            \
        if (e__ != cudaSuccess) return cfail(c, MPT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)
#define CNC(call)                                                                                                          \
    do {                                                                                                                   \
        const ncclResult_t r__ = (call);                                                                                   \
        if (r__ != 0) return cfail(c, MPT_ERR_CUDA, std::string(#call) + ": " + nccl()->GetErrorString(r__));               \
    } while (0)

mpt_status make_comm(int device, int rank, int world, mpt_comm** out) {
    mpt_comm* c = nullptr;
    if (!out) return cfail(nullptr, MPT_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return cfail(nullptr, MPT_ERR_INVALID_ARGUMENT, "rank / world out of range");
    CCU(cudaSetDevice(device));
    c = new mpt_comm();
    c->device = device; c->rank = rank; c->world = world;
    cudaError_t e = cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming);
    for (uint32_t i = 0; i < mpt_comm::kSlots && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming);
    if (e != cudaSuccess) {
        const std::string msg = std::string("mpt_comm: ") + cudaGetErrorString(e);
        if (c->side) cudaStreamDestroy(c->side);
        if (c->ready) cudaEventDestroy(c->ready);
        for (cudaEvent_t ev : c->done)
            if (ev) cudaEventDestroy(ev);
        delete c;
        return cfail(nullptr, MPT_ERR_CUDA, msg);
    }
    *out = c;
    return MPT_OK;
}
} // namespace

extern "C" {

MPT_API const char* mpt_comm_last_error(const mpt_comm* c) { return c ? c->error.c_str() : g_comm_error.c_str(); }

MPT_API mpt_status mpt_comm_unique_id(uint8_t id[MPT_COMM_UNIQUE_ID_BYTES]) {
    mpt_comm* c = nullptr;
    if (!id) return cfail(c, MPT_ERR_INVALID_ARGUMENT, "id is NULL");
    if (!nccl()) return cfail(c, MPT_ERR_UNSUPPORTED, g_nccl.error);
    ncclUniqueId u;
    CNC(nccl()->GetUniqueId(&u));
    static_assert(sizeof(u) == MPT_COMM_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
    std::memcpy(id, &u, sizeof(u));
    return MPT_OK;
}

MPT_API mpt_status mpt_comm_create(int device, const uint8_t id[MPT_COMM_UNIQUE_ID_BYTES], int rank, int world, mpt_comm** out) {
    if (!id) return cfail(nullptr, MPT_ERR_INVALID_ARGUMENT, "id is NULL");
    if (!nccl()) return cfail(nullptr, MPT_ERR_UNSUPPORTED, g_nccl.error);
    mpt_comm* c = nullptr;
    const mpt_status s = make_comm(device, rank, world, &c);
    if (s != MPT_OK) return s;
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof(u));
    const ncclResult_t r = nccl()->CommInitRank(&c->comm, world, u, rank);
    if (r != 0) {
        const std::string msg = std::string("ncclCommInitRank: ") + nccl()->GetErrorString(r);
        mpt_comm_destroy(c);
        return cfail(nullptr, MPT_ERR_CUDA, msg);
    }
    c->owns_comm = true;
    *out = c;
    return MPT_OK;
}

MPT_API mpt_status mpt_comm_adopt(int device, void* nccl_comm, int rank, int world, mpt_comm** out) {
    if (!nccl_comm) return cfail(nullptr, MPT_ERR_INVALID_ARGUMENT, "nccl_comm is NULL");
    if (!nccl()) return cfail(nullptr, MPT_ERR_UNSUPPORTED, g_nccl.error);
    mpt_comm* c = nullptr;
    const mpt_status s = make_comm(device, rank, world, &c);
    if (s != MPT_OK) return s;
    c->comm = (ncclComm_t)nccl_comm;
    *out = c;
    return MPT_OK;
}

MPT_API void mpt_comm_destroy(mpt_comm* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->side) cudaStreamSynchronize(c->side);
    if (c->comm && c->owns_comm && nccl()) nccl()->CommDestroy(c->comm);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->ready) cudaEventDestroy(c->ready);
    for (cudaEvent_t ev : c->done)
        if (ev) cudaEventDestroy(ev);
    delete c;
}

MPT_API mpt_status mpt_reduce_begin(mpt_comm* c, const void* d_partial_sum, void* d_out, void* h_out, uint32_t width, uint32_t height, uint32_t spp_total,
                                    int root, uint32_t flags, void* render_stream, uint32_t slot) {
    if (!c || !d_partial_sum || !width || !height) return cfail(c, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    if (root < 0 || root >= c->world) return cfail(c, MPT_ERR_INVALID_ARGUMENT, "root out of range");
    if (slot >= mpt_comm::kSlots) return cfail(c, MPT_ERR_LIMIT, "slot must be < 4");
    if (c->pending[slot]) return cfail(c, MPT_ERR_INVALID_ARGUMENT, "this slot still holds a reduce: call mpt_reduce_end / mpt_reduce_wait on it first");
    const bool all = (flags & MPT_REDUCE_ALL_RANKS) != 0, is_root = all || c->rank == root;
    if (is_root && !d_out) return cfail(c, MPT_ERR_INVALID_ARGUMENT, "d_out is NULL on a rank that receives the frame");
    CCU(cudaSetDevice(c->device));
    const size_t npix = size_t(width) * height, count = npix * 4;
    CCU(cudaEventRecord(c->ready, (cudaStream_t)render_stream));
    CCU(cudaStreamWaitEvent(c->side, c->ready, 0));
    // the reduce lands in d_out on the receiving ranks; the other ranks only send (NCCL ignores recvbuff there, but wants it non-NULL)
    void* recv = is_root ? d_out : const_cast<void*>(d_partial_sum);
    if (c->world > 1) {
        if (all) CNC(nccl()->AllReduce(d_partial_sum, recv, count, kNcclFloat, kNcclSum, c->comm, c->side));
        else CNC(nccl()->Reduce(d_partial_sum, recv, count, kNcclFloat, kNcclSum, root, c->comm, c->side));
    } else if (d_out != d_partial_sum) {
        CCU(cudaMemcpyAsync(d_out, d_partial_sum, count * sizeof(float), cudaMemcpyDeviceToDevice, c->side));
    }
    if (is_root) {
        if (!(flags & MPT_REDUCE_KEEP_SUM) && spp_total > 0) {
            k_normalise<<<uint32_t((npix + 255) / 256), 256, 0, c->side>>>((const float4*)d_out, (float4*)d_out, uint32_t(npix), float(spp_total));
            CCU(cudaGetLastError());
        }
        if (h_out) CCU(cudaMemcpyAsync(h_out, d_out, count * sizeof(float), cudaMemcpyDeviceToHost, c->side));
    }
    CCU(cudaEventRecord(c->done[slot], c->side));
    c->pending[slot] = true;
    return MPT_OK;
}

MPT_API mpt_status mpt_reduce_end(mpt_comm* c, uint32_t slot, void* stream) {
    if (!c) return MPT_ERR_INVALID_ARGUMENT;
    if (slot >= mpt_comm::kSlots) return cfail(c, MPT_ERR_LIMIT, "slot must be < 4");
    if (!c->pending[slot]) return cfail(c, MPT_ERR_INVALID_ARGUMENT, "no reduce is in flight in this slot");
    CCU(cudaSetDevice(c->device));
    CCU(cudaStreamWaitEvent((cudaStream_t)stream, c->done[slot], 0));
    c->pending[slot] = false;
    return MPT_OK;
}

MPT_API mpt_status mpt_reduce_wait(mpt_comm* c, uint32_t slot) {
    if (!c) return MPT_ERR_INVALID_ARGUMENT;
    if (slot >= mpt_comm::kSlots) return cfail(c, MPT_ERR_LIMIT, "slot must be < 4");
    if (!c->pending[slot]) return cfail(c, MPT_ERR_INVALID_ARGUMENT, "no reduce is in flight in this slot");
    CCU(cudaSetDevice(c->device));
    CCU(cudaEventSynchronize(c->done[slot]));
    c->pending[slot] = false;
    return MPT_OK;
}

MPT_API mpt_status mpt_comm_info(uint32_t* nccl_version) {
    mpt_comm* c = nullptr;
    if (!nccl()) return cfail(c, MPT_ERR_UNSUPPORTED, g_nccl.error);
    int v = 0;
    CNC(nccl()->GetVersion(&v));
    if (nccl_version) *nccl_version = uint32_t(v);
    return MPT_OK;
}

} // extern "C"
