"""Child process of tests/test_interop.py: plays the part of the OTHER owner of a shared buffer (inside merian: the Vulkan side).
Allocates device memory with the CUDA virtual-memory API, exports it as a POSIX file descriptor (cuMemExportToShareableHandle), sends the
descriptor to the parent over the inherited Unix socket (SCM_RIGHTS), fills the buffer with a marker, and — once the parent says the
library has rendered into its own mapping of the same memory — reads the buffer back through ITS mapping and reports a digest."""
import hashlib
import socket
import struct
import sys

import numpy as np
from cuda.bindings import driver as cu


def ck(r):
    err = r[0]
    if err != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"CUDA driver error {err}")
    return r[1] if len(r) == 2 else r[1:]


def main():
    sock = socket.socket(fileno=int(sys.argv[1]))
    nbytes = int(sys.argv[2])
    ck(cu.cuInit(0))
    dev = ck(cu.cuDeviceGet(0))
    ctx = ck(cu.cuDevicePrimaryCtxRetain(dev))
    ck(cu.cuCtxSetCurrent(ctx))
    prop = cu.CUmemAllocationProp()
    prop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
    prop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    prop.location.id = 0
    prop.requestedHandleTypes = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
    gran = ck(cu.cuMemGetAllocationGranularity(prop, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM))
    size = (nbytes + gran - 1) // gran * gran
    handle = ck(cu.cuMemCreate(size, prop, 0))
    fd = int(ck(cu.cuMemExportToShareableHandle(handle, cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0)))
    va = ck(cu.cuMemAddressReserve(size, 0, 0, 0))
    ck(cu.cuMemMap(va, size, 0, handle, 0))
    acc = cu.CUmemAccessDesc()
    acc.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    acc.location.id = 0
    acc.flags = cu.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
    ck(cu.cuMemSetAccess(va, size, [acc], 1))
    ck(cu.cuMemsetD32(va, 0x7FC00000, size // 4))  # quiet NaNs: whatever the library does not write stays recognisable
    ck(cu.cuCtxSynchronize())
    socket.send_fds(sock, [struct.pack("<Q", size)], [fd])
    msg = sock.recv(16)  # the parent has rendered into its mapping
    assert msg.startswith(b"rendered"), msg
    host = np.empty(nbytes, np.uint8)
    ck(cu.cuMemcpyDtoH(host.ctypes.data, va, nbytes))
    sock.sendall(hashlib.sha256(host.tobytes()).digest())
    sock.recv(16)  # parent released its mapping
    ck(cu.cuMemUnmap(va, size))
    ck(cu.cuMemAddressFree(va, size))
    ck(cu.cuMemRelease(handle))
    sock.sendall(b"bye")


if __name__ == "__main__":
    main()
