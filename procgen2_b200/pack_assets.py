"""Asset packer: reference PNG sprite sheets -> one packed blob the engine uploads into its
device texture atlas (replaces Asset_Texture::load = IMG_Load + SDL_CreateTextureFromSurface,
/root/reference/games/coinrun/common_assets.cpp:3-17).

Run ONCE in the build container (needs PIL and the reference's assets/ directory):

    python -m procgen2_b200.pack_assets [--reference /root/reference]

The list of reachable textures is discovered by letting the compiled reference games
(oracle/_ref, asset-discovery mode of the SDL stand-in) log every IMG_Load path during
cenv_make. Pixels are decoded with PIL: RGBA / palette+tRNS -> RGBA8 (alpha-blended
texture), RGB / opaque palette -> RGB8 (opaque copy texture), then zlib-compressed.

Blob layout (little endian):
    char  magic[8] = "PG2ASSET"; u32 version = 1; u32 count;
    entry[count] = { char name[120]; u32 w, h, channels, reserved; u64 offset, zsize; }
    zlib streams ...
"""
import argparse
import ctypes
import os
import struct
import subprocess
import sys
import tempfile
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DEFAULT_OUT = os.path.join(HERE, "data", "assets.bin")


def discover(reference):
    sys.path.insert(0, ROOT)
    from oracle import build_ref
    build_ref.build()
    paths = []
    for g in build_ref.GAMES:
        with tempfile.NamedTemporaryFile("r", suffix=".log") as lg:
            code = (
                "import ctypes,os;os.chdir(%r);l=ctypes.CDLL(%r);"
                "l.cenv_make.argtypes=[ctypes.c_char_p,ctypes.c_void_p,ctypes.c_int32];"
                "assert l.cenv_make(b'',None,0)==0" % (reference, build_ref.lib_path(g))
            )
            env = dict(os.environ, PG2O_DISCOVER=lg.name)
            subprocess.check_call([sys.executable, "-c", code], env=env)
            for line in open(lg.name):
                p = line.strip()
                if p and p not in paths:
                    paths.append(p)
    return paths


def pack(reference, out, paths):
    from PIL import Image
    entries = []
    blobs = []
    for p in paths:
        im = Image.open(os.path.join(reference, p))
        has_alpha = im.mode in ("RGBA", "LA", "PA") or "transparency" in im.info
        conv = im.convert("RGBA" if has_alpha else "RGB")
        raw = conv.tobytes()
        z = zlib.compress(raw, 9)
        entries.append((p, conv.width, conv.height, 4 if has_alpha else 3, len(z)))
        blobs.append(z)
    header_size = 16 + 152 * len(entries)
    off = header_size
    with open(out, "wb") as f:
        f.write(b"PG2ASSET" + struct.pack("<II", 1, len(entries)))
        for (name, w, h, ch, zs) in entries:
            nb = name.encode()
            assert len(nb) < 120, name
            f.write(nb.ljust(120, b"\0") + struct.pack("<IIIIQQ", w, h, ch, 0, off, zs))
            off += zs
        for z in blobs:
            f.write(z)
    return entries


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=DEFAULT_OUT)
    a = ap.parse_args()
    paths = discover(a.reference)
    entries = pack(a.reference, a.out, paths)
    raw = sum(w * h * c for (_, w, h, c, _) in entries)
    print("packed %d textures, %.1f MB raw -> %.1f MB" % (len(entries), raw / 1e6, os.path.getsize(a.out) / 1e6))


if __name__ == "__main__":
    main()
