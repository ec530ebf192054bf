"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Recipe that compiles the UNMODIFIED reference games (sources stay where they lie under
/root/reference/games/<g>/) against the SDL3 stand-in in oracle/shim/ and the canonical CPU
rasteriser oracle/raster.c, into oracle/_ref/lib<Game>.so (git-ignored, travels to the GPU box).

Flags: -O3 -DNDEBUG = the reference's default Release build (games/coinrun/CMakeLists.txt:12-15);
-ffp-contract=off keeps x86-64 baseline semantics (no FMA contraction).
A per-game probe translation unit (oracle/probe/probe_<g>.cpp) is linked in to expose RNG state,
tile maps and entity floats for the parity tests; it only READS the reference's globals.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PG2_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

GAMES = {
    "maze": "Maze",
    "coinrun": "CoinRun",
    "bossfight": "BossFight",
    "chaser": "Chaser",
    "climber": "Climber",
    "caveflyer": "CaveFlyer",
    "jumper": "Jumper",
}


def lib_path(game):
    return os.path.join(OUT, "lib%s.so" % GAMES[game])


def build(games=None, force=False, verbose=False):
    if not os.path.isdir(os.path.join(REF, "games")):
        return False  # GPU box: use the prebuilt files
    os.makedirs(OUT, exist_ok=True)
    shim = os.path.join(HERE, "shim")
    alib = os.path.join(OUT, "libpg2o_assets.so")
    asrc = os.path.join(HERE, "assets_blob.cpp")
    if force or not os.path.exists(alib) or os.path.getmtime(asrc) > os.path.getmtime(alib):
        subprocess.check_call(["g++", "-std=gnu++14", "-O2", "-fPIC", "-shared", asrc, "-lz", "-Wl,-soname,libpg2o_assets.so", "-o", alib])
    for g in games or GAMES:
        out = lib_path(g)
        gdir = os.path.join(REF, "games", g)
        srcs = sorted(os.path.join(gdir, f) for f in os.listdir(gdir) if f.endswith(".cpp"))
        extra = [os.path.join(shim, "shim.cpp"), os.path.join(HERE, "raster.c")]
        probe = os.path.join(HERE, "probe", "probe_%s.cpp" % g)
        if os.path.exists(probe):
            extra.append(probe)
        deps = srcs + extra + [os.path.join(shim, "SDL3", "SDL.h"), os.path.join(HERE, "raster.h"), os.path.join(HERE, "probe", "probe_common.h")]
        if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps if os.path.exists(d)):
            continue
        objs = []
        # raster.c is C: compile separately
        ro = os.path.join(OUT, "raster_%s.o" % g)
        subprocess.check_call(["gcc", "-O3", "-DNDEBUG", "-fPIC", "-ffp-contract=off", "-c", os.path.join(HERE, "raster.c"), "-o", ro])
        objs.append(ro)
        cmd = ["g++", "-std=gnu++14", "-O3", "-DNDEBUG", "-fPIC", "-shared", "-ffp-contract=off", "-w",
               "-I" + shim, "-I" + gdir, "-I" + HERE] + srcs + [e for e in extra if e.endswith(".cpp")] + objs + ["-L" + OUT, "-lpg2o_assets", "-Wl,-rpath,$ORIGIN", "-o", out]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        os.remove(ro)
    return True


if __name__ == "__main__":
    ok = build(sys.argv[1:] or None, force=True, verbose=True)
    print("built" if ok else "reference sources not present; nothing built")
