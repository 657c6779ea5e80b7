"""PFM in / out. The checker is a numpy restatement of the format the reference reads and writes
(src/merian/io/image_io.cpp:53-168): 'PF'/'Pf', 'W H', scale (negative = little endian), one whitespace byte, bottom-up f32 rows."""
import numpy as np
import pytest

from merian_b200 import image_io
from merian_b200._lib import MptError


def np_pfm_bytes(img, big_endian=False):
    """What save_pfm of the reference writes for a top-down (H, W, C) image (RGBA -> RGB)."""
    h, w, c = img.shape
    oc = 1 if c == 1 else 3
    head = f"{'PF' if oc == 3 else 'Pf'}\n{w} {h}\n{'1' if big_endian else '-1'}\n".encode()
    body = img[::-1, :, :oc].astype(">f4" if big_endian else "<f4").tobytes()
    return head + body


def np_pfm_load(data, desired=0):
    """load_pfm of the reference (image_io.cpp:54-136) on a byte string."""
    parts = data.split(maxsplit=4)  # magic, w, h, scale, rest -- but the pixel block may start with whitespace bytes:
    magic, w, h, scale = parts[0], int(parts[1]), int(parts[2]), float(parts[3])
    sc = 3 if magic == b"PF" else 1
    n = w * h * sc * 4
    body = data[len(data) - n:]
    a = np.frombuffer(body, "<f4" if scale < 0 else ">f4").reshape(h, w, sc)[::-1].astype(np.float32)
    oc = desired or sc
    out = np.zeros((h, w, oc), np.float32)
    full = np.concatenate([a if sc == 3 else np.repeat(a, 3, -1), np.ones((h, w, 1), np.float32)], -1)
    out[...] = full[..., :oc]
    return out


@pytest.mark.parametrize("c", [1, 3, 4])
def test_save_matches_reference_format_byte_for_byte(tmp_path, c):
    rng = np.random.default_rng(c)
    img = (rng.standard_normal((7, 5, c)) * 100).astype(np.float32)
    img[0, 0, 0] = np.inf
    img[1, 1, 0] = np.nan
    p = tmp_path / "a.pfm"
    image_io.save_pfm(p, img)
    assert p.read_bytes() == np_pfm_bytes(img)
    w, h, ch = image_io.pfm_info(p)
    assert (w, h, ch) == (5, 7, 1 if c == 1 else 3)


@pytest.mark.parametrize("file_c,desired", [(3, 0), (3, 4), (3, 1), (1, 0), (1, 3), (1, 4)])
@pytest.mark.parametrize("big_endian", [False, True])
def test_load_channel_expansion_and_endianness(tmp_path, file_c, desired, big_endian):
    rng = np.random.default_rng(10 * file_c + desired)
    img = rng.random((4, 6, file_c), dtype=np.float32)
    data = np_pfm_bytes(img, big_endian)
    p = tmp_path / "b.pfm"
    p.write_bytes(data)
    got = image_io.load_pfm(p, desired)
    ref = np_pfm_load(data, desired)
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    if desired == 4:
        assert (got[..., 3] == 1.0).all()  # alpha = 1 (image_io.cpp:114-116)


def test_round_trip_rgba_drops_alpha_only(tmp_path):
    img = np.random.default_rng(0).random((9, 3, 4), dtype=np.float32)
    p = tmp_path / "c.pfm"
    image_io.save_pfm(p, img)
    back = image_io.load_pfm(p, 4)
    assert np.array_equal(back[..., :3], img[..., :3]) and (back[..., 3] == 1).all()


def test_errors(tmp_path):
    with pytest.raises(MptError, match="cannot open"):
        image_io.load_pfm(tmp_path / "missing.pfm")
    bad = tmp_path / "bad.pfm"
    bad.write_bytes(b"P6\n2 2\n255\n" + bytes(12))
    with pytest.raises(MptError, match="bad PFM magic"):
        image_io.load_pfm(bad)
    short = tmp_path / "short.pfm"
    short.write_bytes(b"PF\n4 4\n-1\n" + bytes(40))
    with pytest.raises(MptError, match="truncated"):
        image_io.load_pfm(short)
    neg = tmp_path / "neg.pfm"
    neg.write_bytes(b"PF\n-4 4\n-1\n")
    with pytest.raises(MptError, match="malformed"):
        image_io.load_pfm(neg)
    huge = tmp_path / "huge.pfm"  # a header that would ask std::vector for 16 EB: reported, not thrown across the C ABI
    huge.write_bytes(b"PF\n2000000000 2000000000\n-1\n" + bytes(16))
    with pytest.raises(MptError, match="implausible"):
        image_io.load_pfm(huge)
    hdr = tmp_path / "huge.hdr"
    hdr.write_bytes(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 2000000000 +X 2000000000\n" + bytes(16))
    with pytest.raises(MptError, match="implausible"):
        image_io.load_hdr(hdr)
    with pytest.raises(MptError, match="1, 3 or 4"):
        image_io.save_pfm(tmp_path / "two.pfm", np.zeros((2, 2, 2), np.float32))


def _rgbe(img):
    """float RGB -> RGBE bytes (Ward): the shared exponent of the largest component, 8-bit mantissas."""
    m = img.max(-1)
    e = np.where(m > 1e-32, np.ceil(np.log2(np.maximum(m, 1e-38)) + 1e-12), -128).astype(np.int32)
    scale = np.where(m > 1e-32, np.ldexp(1.0, 8 - e), 0.0)
    out = np.zeros(img.shape[:-1] + (4,), np.uint8)
    out[..., :3] = np.clip(np.floor(img * scale[..., None]), 0, 255).astype(np.uint8)
    out[..., 3] = np.where(m > 1e-32, e + 128, 0).astype(np.uint8)
    return out


def _rle_scanline(row):  # new-style RLE: per component, runs of >= 3 equal bytes
    w = len(row)
    out = bytearray([2, 2, w >> 8, w & 255])
    for k in range(4):
        comp = row[:, k].tolist()
        i = 0
        while i < w:
            run = 1
            while i + run < w and run < 127 and comp[i + run] == comp[i]:
                run += 1
            if run >= 3:
                out += bytes([128 + run, comp[i]])
                i += run
            else:
                j = i
                while j < w and j - i < 128:
                    r = 1
                    while j + r < w and r < 3 and comp[j + r] == comp[j]:
                        r += 1
                    if r >= 3:
                        break
                    j += 1
                out += bytes([j - i]) + bytes(comp[i:j])
                i = j
    return bytes(out)


@pytest.mark.parametrize("rle,w", [(False, 5), (True, 24), (False, 24)])
def test_hdr_rgbe_decoding(tmp_path, rle, w):
    rng = np.random.default_rng(w + int(rle))
    img = (rng.random((6, w, 3)) * np.array([0.01, 1.0, 3000.0])).astype(np.float32)
    img[0, :4] = 0.0            # (0,0,0,0) -> 0
    img[1, :12] = img[1, 0]     # a run for the RLE path
    rgbe = _rgbe(img)
    head = b"#?RADIANCE\n# comment\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=1.0\n\n-Y 6 +X %d\n" % w
    body = b"".join(_rle_scanline(r) for r in rgbe) if rle else rgbe.tobytes()
    p = tmp_path / "e.hdr"
    p.write_bytes(head + body)
    got = image_io.load_hdr(p, 4)
    ref = rgbe[..., :3].astype(np.float32) * np.ldexp(np.float32(1.0), rgbe[..., 3].astype(np.int32) - 136)[..., None]
    ref[rgbe[..., 3] == 0] = 0
    assert np.array_equal(got[..., :3], ref.astype(np.float32)) and (got[..., 3] == 1).all()
    assert (np.abs(got[..., :3] - img) <= img.max(-1, keepdims=True) * 2.0 ** -7 + 1e-30).all()  # 8-bit mantissa on the shared exponent
    bad = tmp_path / "bad.hdr"
    bad.write_bytes(b"#?RADIANCE\nFORMAT=32-bit_rle_xyze\n\n-Y 1 +X 1\n" + bytes(4))
    with pytest.raises(MptError, match="unsupported HDR format"):
        image_io.load_hdr(bad)
    short = tmp_path / "short.hdr"
    short.write_bytes(head + body[:len(body) // 2])
    with pytest.raises(MptError, match="truncated|corrupt"):
        image_io.load_hdr(short)


@pytest.mark.gpu
def test_env_load_and_image_write_on_device(gpu_ctx, tmp_path):
    """A lat-long .pfm loaded by libmpt lights the scene exactly like the same pixels passed to mpt_env_set; the rendered image
    written by mpt_image_write reads back bit-identically (alpha dropped)."""
    import torch
    from merian_b200 import abi, synthetic
    from merian_b200.scene import SceneTables, upload_scene
    from helpers import render_params
    t = SceneTables(synthetic.terrain(nx=60, ny=30, env_size=(128, 64)))
    upload_scene(gpu_ctx, t)
    cam = t.desc.camera.to_abi(0)
    P = render_params(96, 54, 4)
    ref = gpu_ctx.render_device(cam, P)
    env = t.desc.env
    p = tmp_path / "env.pfm"
    image_io.save_pfm(p, env.pixels)
    gpu_ctx.env_load(p, env.to_local, env.intensity)
    got = gpu_ctx.render_device(cam, P)
    assert torch.equal(got, ref)  # env_eval reads rgb only, the file's missing alpha does not matter
    out = tmp_path / "frame.pfm"
    gpu_ctx.image_write(out, got)
    back = image_io.load_pfm(out, 3)
    assert np.array_equal(back, got.cpu().numpy()[..., :3])
