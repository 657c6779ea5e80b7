// interop.cu — Vulkan <-> CUDA sharing for the merian node (north_star: "shares its output images with the rest of the graph through
// Vulkan<->CUDA external-memory and semaphore interop"; SURVEY.md §7 step 3, §8(b) "cross-API sync options" (2)).
//
// What it replaces in the round-1 adapter: submission.sync_to_cpu_and_back + cudaMemcpy D2H + memcpy into a mapped staging buffer +
// vkCmdCopyBufferToImage. With these entry points the node allocates an exportable VkBuffer (VK_KHR_external_memory_fd,
// vkGetMemoryFdKHR -> opaque fd) and two exportable timeline semaphores (VK_KHR_external_semaphore_fd, vkGetSemaphoreFdKHR), imports
// them ONCE here, and per iteration:  queue->submit(... signal S_vk = n)  ->  mpt_wait_external_semaphore(S_vk, n, stream)  ->
// mpt_render(..., d_ptr of the imported buffer, stream)  ->  mpt_signal_external_semaphore(S_cu, n, stream)  ->  the graph's next
// submit waits S_cu = n and copies buffer -> image on the GPU. No host round trip. Reference hooks: include/merian/vk/command/
// submission.hpp:101-114 (what is replaced), include/merian/vk/command/queue.hpp:29-30 (raw submit with the node's own semaphores),
// include/merian/vk/sync/semaphore_timeline.hpp:10-29 (why merian's own TimelineSemaphore cannot be exported), extension.hpp:70-78
// (the node requests VK_KHR_external_memory_fd + VK_KHR_external_semaphore_fd in query_device_support).
//
// A second memory handle type, the POSIX fd of a CUDA virtual-memory allocation (cuMemExportToShareableHandle), goes through the same
// entry point: it lets a CUDA-only host (and the GPU test, which has no Vulkan) hand a buffer to the library across a process boundary.
#include "mpt_internal.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <fcntl.h>
#include <unistd.h>

#include <string>

struct mpt_external_memory {
    uint32_t handle_type = 0;
    cudaExternalMemory_t ext = nullptr; // MPT_EXTERNAL_MEMORY_OPAQUE_FD
    unsigned long long vmm_handle = 0;  // MPT_EXTERNAL_MEMORY_CUDA_VMM_FD: CUmemGenericAllocationHandle
    unsigned long long va = 0;          //   ... reserved + mapped address range
    uint64_t size = 0;
    void* d_ptr = nullptr;
};

struct mpt_external_semaphore {
    cudaExternalSemaphore_t ext = nullptr;
    uint32_t handle_type = 0;
};

namespace {

// Descriptor ownership, the same for every handle type: a successful import consumes the caller's descriptor, a failed one leaves it
// untouched. The CUDA runtime does not promise the second half (a failed opaque-fd memory import was seen to close the descriptor), so the
// driver only ever gets a duplicate.
void close_if_open(int fd) {
    if (fd >= 0 && fcntl(fd, F_GETFD) != -1) close(fd);
}

#define ICU(call)                                                                                                                  \
    do {                                                                                                                           \
        const cudaError_t e__ = (call);                                                                                            \
        if (e__ != cudaSuccess) return mpt_internal_fail(ctx, MPT_ERR_CUDA, (std::string(#call) + ": " + cudaGetErrorString(e__)).c_str()); \
    } while (0)

// driver entry points for the virtual-memory path, fetched through the runtime (libmpt.so does not link libcuda: it must load, and its
// host-only entry points must work, on machines without a driver)
template <class F> bool driver_fn(const char* name, F& fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return false;
    fn = reinterpret_cast<F>(p);
    return true;
}

mpt_status import_vmm(mpt_context* ctx, int fd, uint64_t size, mpt_external_memory* m) {
    CUresult (*memImport)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
    CUresult (*addrReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*memMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*memRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*addrFree)(CUdeviceptr, size_t) = nullptr;
    if (!driver_fn("cuMemImportFromShareableHandle", memImport) || !driver_fn("cuMemAddressReserve", addrReserve) || !driver_fn("cuMemMap", memMap) ||
        !driver_fn("cuMemSetAccess", setAccess) || !driver_fn("cuMemRelease", memRelease) || !driver_fn("cuMemAddressFree", addrFree))
        return mpt_internal_fail(ctx, MPT_ERR_UNSUPPORTED, "the CUDA driver lacks the virtual-memory entry points");
    CUmemGenericAllocationHandle h = 0;
    CUresult r = memImport(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    if (r != CUDA_SUCCESS) return mpt_internal_fail(ctx, MPT_ERR_CUDA, ("cuMemImportFromShareableHandle failed (" + std::to_string(int(r)) + ")").c_str());
    CUdeviceptr va = 0;
    r = addrReserve(&va, size, 0, 0, 0);
    if (r == CUDA_SUCCESS) {
        r = memMap(va, size, 0, h, 0);
        if (r == CUDA_SUCCESS) {
            CUmemAccessDesc acc{};
            acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
            acc.location.id = mpt_internal_device(ctx);
            acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
            r = setAccess(va, size, &acc, 1);
        }
        if (r != CUDA_SUCCESS) addrFree(va, size);
    }
    if (r != CUDA_SUCCESS) {
        memRelease(h);
        return mpt_internal_fail(ctx, MPT_ERR_CUDA, ("mapping the imported allocation failed (" + std::to_string(int(r)) + "): size must be a multiple of the allocation granularity").c_str());
    }
    m->vmm_handle = h;
    m->va = va;
    m->d_ptr = (void*)va;
    return MPT_OK;
}

} // namespace

extern "C" {

MPT_API mpt_status mpt_import_external_memory_fd(mpt_context* ctx, int fd, uint64_t size, uint32_t handle_type, uint32_t flags, mpt_external_memory** out,
                                                 void** d_ptr) {
    if (!ctx || !out || !d_ptr) return mpt_internal_fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    *out = nullptr;
    *d_ptr = nullptr;
    if (fd < 0 || size == 0) return mpt_internal_fail(ctx, MPT_ERR_INVALID_ARGUMENT, "bad file descriptor or size");
    if (handle_type != MPT_EXTERNAL_MEMORY_OPAQUE_FD && handle_type != MPT_EXTERNAL_MEMORY_CUDA_VMM_FD)
        return mpt_internal_fail(ctx, MPT_ERR_UNSUPPORTED, "unknown external memory handle type");
    ICU(cudaSetDevice(mpt_internal_device(ctx)));
    mpt_external_memory* m = new (std::nothrow) mpt_external_memory();
    if (!m) return mpt_internal_fail(ctx, MPT_ERR_LIMIT, "out of host memory");
    m->handle_type = handle_type;
    m->size = size;
    if (handle_type == MPT_EXTERNAL_MEMORY_OPAQUE_FD) {
        cudaExternalMemoryHandleDesc hd{};
        hd.type = cudaExternalMemoryHandleTypeOpaqueFd; // VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT
        const int dfd = dup(fd);
        if (dfd < 0) {
            delete m;
            return mpt_internal_fail(ctx, MPT_ERR_IO, "dup() of the descriptor failed");
        }
        hd.handle.fd = dfd;                             // on success the driver owns the duplicate
        hd.size = size;
        hd.flags = (flags & MPT_EXTERNAL_MEMORY_DEDICATED) ? cudaExternalMemoryDedicated : 0;
        cudaError_t e = cudaImportExternalMemory(&m->ext, &hd);
        if (e == cudaSuccess) {
            cudaExternalMemoryBufferDesc bd{};
            bd.offset = 0;
            bd.size = size;
            e = cudaExternalMemoryGetMappedBuffer(&m->d_ptr, m->ext, &bd);
            if (e != cudaSuccess) cudaDestroyExternalMemory(m->ext);
        }
        if (e != cudaSuccess) {
            delete m;
            cudaGetLastError();
            close_if_open(dfd);
            return mpt_internal_fail(ctx, MPT_ERR_CUDA, (std::string("importing the external memory failed: ") + cudaGetErrorString(e)).c_str());
        }
        close(fd);
    } else {
        const mpt_status s = import_vmm(ctx, fd, size, m);
        if (s != MPT_OK) {
            delete m;
            return s;
        }
        close(fd); // the imported handle keeps the allocation alive; like the opaque path, a successful import consumes the descriptor
    }
    *out = m;
    *d_ptr = m->d_ptr;
    return MPT_OK;
}

MPT_API mpt_status mpt_release_external_memory(mpt_context* ctx, mpt_external_memory* m) {
    if (!ctx || !m) return mpt_internal_fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    ICU(cudaSetDevice(mpt_internal_device(ctx)));
    ICU(cudaDeviceSynchronize()); // no kernel of this context may still be writing into the mapping
    if (m->handle_type == MPT_EXTERNAL_MEMORY_OPAQUE_FD) {
        if (m->d_ptr) cudaFree(m->d_ptr);
        if (m->ext) cudaDestroyExternalMemory(m->ext);
    } else {
        CUresult (*memUnmap)(CUdeviceptr, size_t) = nullptr;
        CUresult (*memRelease)(CUmemGenericAllocationHandle) = nullptr;
        CUresult (*addrFree)(CUdeviceptr, size_t) = nullptr;
        if (driver_fn("cuMemUnmap", memUnmap) && driver_fn("cuMemRelease", memRelease) && driver_fn("cuMemAddressFree", addrFree)) {
            memUnmap(m->va, m->size);
            addrFree(m->va, m->size);
            memRelease(m->vmm_handle);
        }
    }
    delete m;
    return MPT_OK;
}

MPT_API mpt_status mpt_import_external_semaphore_fd(mpt_context* ctx, int fd, uint32_t handle_type, mpt_external_semaphore** out) {
    if (!ctx || !out) return mpt_internal_fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    *out = nullptr;
    if (fd < 0) return mpt_internal_fail(ctx, MPT_ERR_INVALID_ARGUMENT, "bad file descriptor");
    if (handle_type != MPT_EXTERNAL_SEMAPHORE_OPAQUE_FD && handle_type != MPT_EXTERNAL_SEMAPHORE_TIMELINE_FD)
        return mpt_internal_fail(ctx, MPT_ERR_UNSUPPORTED, "unknown external semaphore handle type");
    ICU(cudaSetDevice(mpt_internal_device(ctx)));
    cudaExternalSemaphoreHandleDesc hd{};
    // VK_EXTERNAL_SEMAPHORE_HANDLE_TYPE_OPAQUE_FD_BIT of a binary or of a VK_SEMAPHORE_TYPE_TIMELINE semaphore
    hd.type = handle_type == MPT_EXTERNAL_SEMAPHORE_TIMELINE_FD ? cudaExternalSemaphoreHandleTypeTimelineSemaphoreFd : cudaExternalSemaphoreHandleTypeOpaqueFd;
    const int dfd = dup(fd);
    if (dfd < 0) return mpt_internal_fail(ctx, MPT_ERR_IO, "dup() of the descriptor failed");
    hd.handle.fd = dfd;
    mpt_external_semaphore* s = new (std::nothrow) mpt_external_semaphore();
    if (!s) {
        close(dfd);
        return mpt_internal_fail(ctx, MPT_ERR_LIMIT, "out of host memory");
    }
    s->handle_type = handle_type;
    const cudaError_t e = cudaImportExternalSemaphore(&s->ext, &hd);
    if (e != cudaSuccess) {
        delete s;
        cudaGetLastError();
        close_if_open(dfd);
        return mpt_internal_fail(ctx, MPT_ERR_CUDA, (std::string("importing the external semaphore failed: ") + cudaGetErrorString(e)).c_str());
    }
    close(fd);
    *out = s;
    return MPT_OK;
}

MPT_API mpt_status mpt_wait_external_semaphore(mpt_context* ctx, mpt_external_semaphore* s, uint64_t value, void* stream) {
    if (!ctx || !s) return mpt_internal_fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    ICU(cudaSetDevice(mpt_internal_device(ctx)));
    cudaExternalSemaphoreWaitParams p{};
    p.params.fence.value = value; // ignored for binary semaphores
    ICU(cudaWaitExternalSemaphoresAsync(&s->ext, &p, 1, (cudaStream_t)stream));
    return MPT_OK;
}

MPT_API mpt_status mpt_signal_external_semaphore(mpt_context* ctx, mpt_external_semaphore* s, uint64_t value, void* stream) {
    if (!ctx || !s) return mpt_internal_fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    ICU(cudaSetDevice(mpt_internal_device(ctx)));
    cudaExternalSemaphoreSignalParams p{};
    p.params.fence.value = value;
    ICU(cudaSignalExternalSemaphoresAsync(&s->ext, &p, 1, (cudaStream_t)stream));
    return MPT_OK;
}

MPT_API mpt_status mpt_release_external_semaphore(mpt_context* ctx, mpt_external_semaphore* s) {
    if (!ctx || !s) return mpt_internal_fail(ctx, MPT_ERR_INVALID_ARGUMENT, "NULL argument");
    ICU(cudaSetDevice(mpt_internal_device(ctx)));
    if (s->ext) cudaDestroyExternalSemaphore(s->ext);
    delete s;
    return MPT_OK;
}

} // extern "C"
