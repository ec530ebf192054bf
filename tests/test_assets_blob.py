"""Build-container test (needs /root/reference/assets + PIL; skipped on the GPU box): every entry of the packed asset
blob against its source PNG decoded independently with PIL, through (1) a plain Python reader of the blob layout,
(2) the PRODUCT's loader (assets.cpp via pg2_load_texture_host: the texels the engine uploads into its device atlas),
(3) the oracle's loader (oracle/assets_blob.cpp: the texels the reference games get from IMG_Load in the stand-in).
A packer or loader bug would otherwise be invisible: oracle and engine both read the same blob."""
import ctypes
import os
import struct
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BLOB = os.path.join(ROOT, "procgen2_b200", "data", "assets.bin")
REF = os.environ.get("PG2_REFERENCE", "/root/reference")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "assets")), reason="reference assets not present")


def blob_entries():
    with open(BLOB, "rb") as f:
        assert f.read(8) == b"PG2ASSET"
        ver, n = struct.unpack("<II", f.read(8))
        assert ver == 1
        out = []
        for _ in range(n):
            raw = f.read(152)
            name = raw[:120].split(b"\0", 1)[0].decode()
            w, h, ch, _, off, zs = struct.unpack("<IIIIQQ", raw[120:])
            out.append((name, w, h, ch, off, zs))
    return out


def pil_rgba(name):
    """-> (texels u8 [h, w, 4], has_alpha) by the rule SDL_image + SDL_CreateTextureFromSurface follow: RGBA / LA /
    palette + tRNS = alpha-blended texture; RGB / opaque palette = opaque copy (A = 255)."""
    from PIL import Image
    im = Image.open(os.path.join(REF, name))
    has_alpha = im.mode in ("RGBA", "LA", "PA") or "transparency" in im.info
    a = np.asarray(im.convert("RGBA")).copy()
    if not has_alpha:
        a[..., 3] = 255
    return a, has_alpha


def test_blob_covers_every_texture_the_games_name():
    """Every texture name the engine asks for is in the blob (and the blob holds no duplicates)."""
    names = [e[0] for e in blob_entries()]
    assert len(names) == len(set(names)) and len(names) >= 200
    for n in names:
        assert os.path.exists(os.path.join(REF, n)), n


def test_blob_texels_equal_the_pngs():
    data = open(BLOB, "rb").read()
    for name, w, h, ch, off, zs in blob_entries():
        want, has_alpha = pil_rgba(name)
        assert (h, w) == want.shape[:2], name
        assert ch == (4 if has_alpha else 3), name
        raw = np.frombuffer(zlib.decompress(data[off:off + zs]), np.uint8).reshape(h, w, ch)
        np.testing.assert_array_equal(raw, want[..., :ch], err_msg=name)


def test_product_loader_equals_the_pngs():
    from procgen2_b200 import build
    build.build()
    L = ctypes.CDLL(build.ENGINE)
    L.pg2_load_texture_host.restype = ctypes.c_int64
    L.pg2_load_texture_host.argtypes = [ctypes.c_char_p, ctypes.c_char_p] + [ctypes.POINTER(ctypes.c_int32)] * 3 + [ctypes.c_void_p, ctypes.c_int64]
    for name, w, h, ch, _, _ in blob_entries():
        want, has_alpha = pil_rgba(name)
        cw, chh, cb = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        out = np.empty(w * h, np.uint32)
        n = L.pg2_load_texture_host(BLOB.encode(), name.encode(), ctypes.byref(cw), ctypes.byref(chh), ctypes.byref(cb), out.ctypes.data, out.size)
        assert n == w * h and (cw.value, chh.value) == (w, h) and cb.value == int(has_alpha), name
        np.testing.assert_array_equal(out.view(np.uint8).reshape(h, w, 4), want, err_msg=name)
    assert L.pg2_load_texture_host(BLOB.encode(), b"assets/does/not/exist.png", None, None, None, None, 0) < 0


def test_oracle_loader_equals_the_pngs():
    from oracle import build_ref
    if not build_ref.build():
        pytest.skip("oracle not buildable here")
    L = ctypes.CDLL(os.path.join(build_ref.OUT, "libpg2o_assets.so"))
    L.pg2o_blob_load.restype = ctypes.POINTER(ctypes.c_uint8)
    L.pg2o_blob_load.argtypes = [ctypes.c_char_p, ctypes.c_char_p] + [ctypes.POINTER(ctypes.c_int)] * 3
    for name, w, h, ch, _, _ in blob_entries():
        want, has_alpha = pil_rgba(name)
        cw, chh, ca = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        p = L.pg2o_blob_load(BLOB.encode(), name.encode(), ctypes.byref(cw), ctypes.byref(chh), ctypes.byref(ca))
        assert p and (cw.value, chh.value) == (w, h) and bool(ca.value) == has_alpha, name
        got = np.ctypeslib.as_array(p, shape=(h, w, 4))
        np.testing.assert_array_equal(got[..., :3], want[..., :3], err_msg=name)
        if has_alpha:
            np.testing.assert_array_equal(got[..., 3], want[..., 3], err_msg=name)
