"""External-memory / semaphore import of the C-ABI (include/mpt.h "Vulkan <-> CUDA interop"), exercised without Vulkan: a child process
owns a device allocation and exports it as a file descriptor (CUDA virtual-memory API), libmpt imports the descriptor, renders into the
imported mapping, and the child finds the frame in ITS mapping of the same memory — the import / render-into / read-back path of the
merian node, with the child in the role of the Vulkan side. Semaphores can only be created by Vulkan, so for them the test pins the
error behaviour (a descriptor that is not a semaphore is refused and stays with the caller)."""
import ctypes as C
import hashlib
import os
import socket
import struct
import subprocess
import sys

import numpy as np
import pytest

from merian_b200 import abi
from merian_b200.scene import upload_scene

from helpers import tables

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_render_into_memory_imported_from_another_process(gpu_ctx):
    import torch
    t = tables("cornell")
    upload_scene(gpu_ctx, t)
    cam = t.desc.camera.to_abi(0)
    w = h = 96
    nbytes = w * h * 16
    P = abi.RenderParams(w, h, 3, 8, 5, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0)
    want = gpu_ctx.render_device(cam, P).cpu().numpy()

    parent, child = socket.socketpair(socket.AF_UNIX, socket.SOCK_STREAM)
    proc = subprocess.Popen([sys.executable, os.path.join(HERE, "interop_child.py"), str(child.fileno()), str(nbytes)], pass_fds=[child.fileno()])
    child.close()
    try:
        msg, fds, _, _ = socket.recv_fds(parent, 8, 1)
        assert len(fds) == 1, "the child did not send a descriptor"
        size = struct.unpack("<Q", msg)[0]
        l = gpu_ctx.l
        mem, d_ptr = C.c_void_p(), C.c_void_p()
        # the opaque-fd path is the Vulkan one; a CUDA virtual-memory descriptor is a different kind of object and must be refused cleanly
        dup = os.dup(fds[0])
        st = l.mpt_import_external_memory_fd(gpu_ctx.h, dup, size, abi.EXTERNAL_MEMORY_OPAQUE_FD, 0, C.byref(mem), C.byref(d_ptr))
        opaque_accepted = st == abi.MPT_OK
        print("opaque-fd import of a CUDA VMM descriptor:", "accepted" if opaque_accepted else l.mpt_last_error(gpu_ctx.h).decode())
        if opaque_accepted:
            gpu_ctx._check(l.mpt_release_external_memory(gpu_ctx.h, mem))
        else:
            os.close(dup)  # a failed import leaves the descriptor with the caller
        st = l.mpt_import_external_memory_fd(gpu_ctx.h, fds[0], size, abi.EXTERNAL_MEMORY_CUDA_VMM_FD, 0, C.byref(mem), C.byref(d_ptr))
        gpu_ctx._check(st)
        assert d_ptr.value
        # render straight into the imported mapping (mpt_render's d_irradiance), on the current stream
        gpu_ctx._check(l.mpt_render(gpu_ctx.h, C.byref(cam), C.byref(P), None, None, d_ptr, gpu_ctx._stream()))
        torch.cuda.synchronize()
        parent.sendall(b"rendered")
        digest = parent.recv(32)
        assert digest == hashlib.sha256(want.tobytes()).digest(), "the other process does not see the frame the library rendered"
        gpu_ctx._check(l.mpt_release_external_memory(gpu_ctx.h, mem))
        parent.sendall(b"released")
        assert parent.recv(3) == b"bye"
    finally:
        parent.close()
        proc.wait(timeout=60)
    assert proc.returncode == 0


def test_semaphore_import_refuses_a_descriptor_that_is_not_a_semaphore(gpu_ctx):
    l = gpu_ctx.l
    sem = C.c_void_p()
    fd = os.open("/dev/null", os.O_RDWR)
    try:
        for kind in (abi.EXTERNAL_SEMAPHORE_OPAQUE_FD, abi.EXTERNAL_SEMAPHORE_TIMELINE_FD):
            st = l.mpt_import_external_semaphore_fd(gpu_ctx.h, fd, kind, C.byref(sem))
            assert st == abi.MPT_ERR_CUDA and not sem.value
            assert b"semaphore" in l.mpt_last_error(gpu_ctx.h)
        os.fstat(fd)  # still ours: a failed import does not consume the descriptor
        assert l.mpt_import_external_semaphore_fd(gpu_ctx.h, -1, 0, C.byref(sem)) == abi.MPT_ERR_INVALID_ARGUMENT
        assert l.mpt_import_external_semaphore_fd(gpu_ctx.h, fd, 7, C.byref(sem)) == abi.MPT_ERR_UNSUPPORTED
        mem, ptr = C.c_void_p(), C.c_void_p()
        assert l.mpt_import_external_memory_fd(gpu_ctx.h, fd, 4096, abi.EXTERNAL_MEMORY_OPAQUE_FD, 0, C.byref(mem), C.byref(ptr)) == abi.MPT_ERR_CUDA
        assert l.mpt_import_external_memory_fd(gpu_ctx.h, fd, 0, abi.EXTERNAL_MEMORY_OPAQUE_FD, 0, C.byref(mem), C.byref(ptr)) == abi.MPT_ERR_INVALID_ARGUMENT
    finally:
        os.close(fd)
