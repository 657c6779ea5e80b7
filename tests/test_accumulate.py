"""Accumulate node, plain mode: the oracle against the closed forms of accumulate.slang:103-255, and the CUDA
kernels (standalone and fused into the resolve kernel) against the oracle, bit for bit."""
import numpy as np
import pytest

from merian_b200 import abi

from helpers import render_params, tables

INF = float("inf")


def acc_params(w, h, alpha=1.0, max_history=INF, hard_clamp=INF, clear=0, flags=0, normal_reject_cos=0.8, z_reject_percent=0.1):
    return abi.AccumulateParams(w, h, alpha, max_history, hard_clamp, normal_reject_cos, z_reject_percent, clear, flags)


def f16_bits(x):
    return int(np.float16(x).view(np.uint16))


def random_inputs(w, h, seed=0, with_bad=True):
    rng = np.random.default_rng(seed)
    src = rng.random((h, w, 4), dtype=np.float32) * 4.0
    prev = rng.random((h, w, 4), dtype=np.float32) * 4.0
    if with_bad:
        src[0, 0, 1] = np.nan
        src[1, 2, 0] = np.inf
        prev[2, 1, 2] = np.nan
        prev[3, 3, 3] = -np.inf
    length = rng.integers(1, 3000, size=(h, w)).astype(np.float32)
    z = (rng.random((h, w), dtype=np.float32) * 100.0).astype(np.float32)
    hist = np.zeros((h, w, 2), np.uint32)
    hist[..., 0] = rng.integers(0, 2 ** 32, size=(h, w), dtype=np.uint64).astype(np.uint32)
    hist[..., 1] = z.astype(np.float16).view(np.uint16).astype(np.uint32) | (length.astype(np.float16).view(np.uint16).astype(np.uint32) << 16)
    tex0 = np.zeros((h, w, 2), np.uint32)
    tex0[..., 0] = hist[..., 0]
    flip = rng.random((h, w)) < 0.3  # a third of the pixels see a different surface than their history
    tex0[..., 0][flip] = rng.integers(0, 2 ** 32, size=int(flip.sum()), dtype=np.uint64).astype(np.uint32)
    tex2 = np.zeros((h, w, 4), np.float32)
    tex2[..., 2] = z * (1.0 + 0.5 * (rng.random((h, w)) < 0.2))
    return src, prev, hist, tex0, tex2


# ------------------------------------------------------------------------------------------------- oracle (CPU)
def test_clear_copies_src_and_resets_history(oracle_lib):
    from oracle.oracle_py import accumulate
    src, prev, hist, tex0, tex2 = random_inputs(8, 6)
    out, h = accumulate(acc_params(8, 6, clear=1), src, None, None, tex0, tex2)
    expect = src.copy()
    expect[0, 0] = 0  # NaN/Inf samples become 0 (accumulate.slang:128-129)
    expect[1, 2] = 0
    assert np.array_equal(out, expect)
    assert (h[..., 0] == tex0[..., 0]).all()
    assert ((h[..., 1] >> 16) == f16_bits(1.0)).all()
    z16 = np.minimum(tex2[..., 2], 65504.0).astype(np.float16).view(np.uint16)
    assert ((h[..., 1] & 0xFFFF) == z16).all()
    out0, _ = accumulate(acc_params(8, 6, alpha=0.0), src, prev, hist)  # alpha == 0 takes the same shortcut (:141)
    assert np.array_equal(out0, expect)


def test_running_mean_and_f16_history_cap(oracle_lib):
    """alpha = 1, inf history: out_N is the mean of the first N frames while the f16 counter can still count (SURVEY §8 A11:
    'N stored as f16 => saturates at 2048'); with MPT_ACC_HISTORY_F32 the counter does not saturate."""
    from oracle.oracle_py import accumulate
    w, h = 4, 2
    rng = np.random.default_rng(1)
    frames = rng.random((40, h, w, 4), dtype=np.float32)
    out = hist = None
    for i, f in enumerate(frames):
        out, hist = accumulate(acc_params(w, h, clear=int(i == 0)), f, out, hist)
    np.testing.assert_allclose(out, frames.astype(np.float64).mean(0), rtol=2e-6)
    assert ((hist[..., 1] >> 16) == f16_bits(40.0)).all()
    # the counter: 2047 -> 2048 -> 2048 (2049 is not representable in binary16 and rounds back to 2048)
    ph = np.zeros((h, w, 2), np.uint32)
    ph[..., 1] = f16_bits(2047.0) << 16
    _, h1 = accumulate(acc_params(w, h), frames[0], frames[1], ph)
    _, h2 = accumulate(acc_params(w, h), frames[0], frames[1], h1)
    assert ((h1[..., 1] >> 16) == f16_bits(2048.0)).all() and ((h2[..., 1] >> 16) == f16_bits(2048.0)).all()
    ph32 = np.zeros((h, w, 2), np.uint32)
    ph32[..., 1] = np.float32(5000.0).view(np.uint32)
    o32, h32 = accumulate(acc_params(w, h, flags=abi.ACC_HISTORY_F32), frames[0], frames[1], ph32)
    assert (h32[..., 1].view(np.float32) == 5001.0).all()
    a = np.float32(1.0) / np.float32(5001.0)
    np.testing.assert_array_equal(o32, frames[1] * (np.float32(1.0) - a) + frames[0] * a)


def test_alpha_max_history_and_hard_clamp(oracle_lib):
    from oracle.oracle_py import accumulate
    w, h = 4, 4
    src, prev, hist, _, _ = random_inputs(w, h, seed=3, with_bad=False)
    length = (hist[..., 1] >> 16).astype(np.uint16).view(np.float16).astype(np.float32)
    out, nh = accumulate(acc_params(w, h, alpha=0.9, max_history=16.0), src, prev, hist)
    new_len = np.minimum(length + 1.0, 16.0).astype(np.float32)
    a = np.maximum(np.float32(1.0) - np.float32(0.9), np.float32(1.0) / new_len)[..., None]
    np.testing.assert_array_equal(out, prev * (np.float32(1.0) - a) + src * a)
    assert np.array_equal((nh[..., 1] >> 16).astype(np.uint16).view(np.float16).astype(np.float32), new_len)
    # hard clamp scales rgb by max_l / l and the second moment by its square (:130-137)
    outc, _ = accumulate(acc_params(w, h, clear=1, hard_clamp=0.5), src)
    lum = src[..., 0] * np.float32(0.299) + src[..., 1] * np.float32(0.587) + src[..., 2] * np.float32(0.114)
    lum_c = outc[..., 0] * 0.299 + outc[..., 1] * 0.587 + outc[..., 2] * 0.114
    assert (lum_c <= 0.5 * (1 + 1e-6)).all()
    f = np.where(lum > 0.5, np.float32(0.5) / lum, np.float32(1.0))
    np.testing.assert_allclose(outc[..., 3], src[..., 3] * f * f, rtol=1e-6)


def test_gbuffer_check_rejects_other_surfaces(oracle_lib):
    from oracle.oracle_py import accumulate
    w, h = 16, 16
    src, prev, hist, tex0, tex2 = random_inputs(w, h, seed=5, with_bad=False)
    out, nh = accumulate(acc_params(w, h, flags=abi.ACC_GBUFFER_CHECK), src, prev, hist, tex0, tex2)
    reset = (nh[..., 1] >> 16) == f16_bits(1.0)
    same_normal = tex0[..., 0] == hist[..., 0]
    prev_z = (hist[..., 1] & 0xFFFF).astype(np.uint16).view(np.float16).astype(np.float32)
    same_z = np.abs(np.minimum(tex2[..., 2], 65504.0) - prev_z) <= 0.1 * np.maximum(tex2[..., 2], prev_z)
    assert reset[~same_z].all()              # depth mismatch: history dropped, out = src
    assert not reset[same_normal & same_z].any()
    assert np.array_equal(out[reset], src[reset])


# ------------------------------------------------------------------------------------------------- CUDA vs oracle
@pytest.mark.gpu
@pytest.mark.parametrize("flags,clear,alpha,max_hist,clamp", [(0, 0, 1.0, INF, INF), (0, 1, 1.0, INF, INF), (0, 0, 0.8, 32.0, 1.5),
                                                            (abi.ACC_GBUFFER_CHECK, 0, 1.0, INF, INF), (abi.ACC_HISTORY_F32, 0, 1.0, INF, 2.0),
                                                            (0, 0, 0.0, INF, INF)])
def test_cuda_accumulate_bit_exact(gpu_ctx, flags, clear, alpha, max_hist, clamp):
    import torch
    from oracle.oracle_py import accumulate
    w, h = 61, 37
    src, prev, hist, tex0, tex2 = random_inputs(w, h, seed=11)
    if flags & abi.ACC_HISTORY_F32:
        hist[..., 1] = np.random.default_rng(2).integers(1, 100000, size=(h, w)).astype(np.float32).view(np.uint32)
    A = acc_params(w, h, alpha=alpha, max_history=max_hist, hard_clamp=clamp, clear=clear, flags=flags)
    ref_out, ref_hist = accumulate(A, src, prev, hist, tex0, tex2)
    dev = gpu_ctx.torch_device
    t = lambda a: torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).to(dev)
    out = torch.empty((h, w, 4), dtype=torch.float32, device=dev)
    nh = torch.empty((h, w, 2), dtype=torch.int32, device=dev)
    gpu_ctx.accumulate_device(A, t(src), t(prev), t(hist), out, nh, t(tex0), t(tex2))
    got_out, got_hist = out.cpu().numpy(), nh.cpu().numpy().view(np.uint32)
    assert np.array_equal(got_hist, ref_hist)
    assert np.array_equal(got_out.view(np.uint32), ref_out.view(np.uint32))  # bit-exact
    # in place (prev aliases out) gives the same image
    po, ph = t(prev), t(hist)
    gpu_ctx.accumulate_device(A, t(src), po, ph, po, ph, t(tex0), t(tex2))
    assert np.array_equal(po.cpu().numpy().view(np.uint32), ref_out.view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("spp,pool", [(8, None), (8, "1")])
def test_fused_render_accumulate_equals_render_then_accumulate(gpu_ctx, spp, pool, monkeypatch):
    """mpt_render_accumulate (accumulate fused into k_resolve) == mpt_render + oracle accumulate over 3 iterations; pool = '1'
    forces several passes per call (1 Mi path slots) so the running sums go through the scratch image."""
    import torch
    from merian_b200.context import Context
    from merian_b200.scene import upload_scene
    from oracle.oracle_py import accumulate
    t = tables("cornell")
    if pool:
        monkeypatch.setenv("MPT_POOL_MSLOTS", pool)
        ctx = Context(0)
    else:
        ctx = gpu_ctx
    upload_scene(ctx, t)
    cam = t.desc.camera.to_abi(0)
    w, h = (512, 512) if pool else (96, 64)
    dev = ctx.torch_device
    acc = torch.zeros((h, w, 4), dtype=torch.float32, device=dev)
    hist = torch.zeros((h, w, 2), dtype=torch.int32, device=dev)
    ref_out = ref_hist = None
    for it in range(3):
        P = render_params(w, h, spp, frame=it)
        A = acc_params(w, h, clear=int(it == 0))
        ctx.render_accumulate_device(cam, P, A, acc, hist, acc, hist)
        frame = ctx.render_device(cam, P).cpu().numpy()
        ref_out, ref_hist = accumulate(A, frame, ref_out, ref_hist)
        assert np.array_equal(acc.cpu().numpy().view(np.uint32), ref_out.view(np.uint32)), it
        assert np.array_equal(hist.cpu().numpy().view(np.uint32), ref_hist), it
    if pool:
        ctx.close()


@pytest.mark.gpu
def test_accumulate_argument_errors(gpu_ctx):
    import torch
    from merian_b200._lib import MptError
    dev = gpu_ctx.torch_device
    z4 = torch.zeros((4, 4, 4), dtype=torch.float32, device=dev)
    z2 = torch.zeros((4, 4, 2), dtype=torch.int32, device=dev)
    with pytest.raises(MptError):  # blending needs the previous frame
        gpu_ctx.accumulate_device(acc_params(4, 4), z4, None, None, z4, z2)
    with pytest.raises(MptError):  # the f32 counter stores no depth
        gpu_ctx.accumulate_device(acc_params(4, 4, flags=abi.ACC_GBUFFER_CHECK | abi.ACC_HISTORY_F32), z4, z4, z2, z4, z2, z2, z4)
    with pytest.raises(MptError):
        gpu_ctx.accumulate_device(acc_params(4, 4, alpha=1.5), z4, z4, z2, z4, z2)
