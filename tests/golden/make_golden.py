"""Generates the committed golden vectors from the compiled reference (oracle/_ref).

    python tests/golden/make_golden.py [game ...]

Run in the build container (needs /root/reference for oracle/build_ref.py). For every game:
N_ENVS reference processes seeded SEED+i, cenv_make -> cenv_reset -> T cenv_step with a fixed
action stream (two streams: uniform and "run right / jump" biased) and reset-on-terminate.
Stored per stream: reward[T,N], terminated[T,N], CRC32 of every observation [T+1,N], the raw
reset frame and frames at a few steps, the full tile map and the MT19937 position after make
and after the first reset (level layout + RNG stream pins).
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref, ref_env  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SEED, N_ENVS, T = 4242, 6, 160
KEEP_FRAMES = (0, 1, 40, 159)


def action_streams():
    rs = np.random.RandomState(99)
    return {"uniform": rs.randint(0, 15, size=(T, N_ENVS)).astype(np.int32),
            "biased": rs.choice([6, 7, 8, 8, 7, 5, 4, 1], size=(T, N_ENVS)).astype(np.int32)}


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xffffffff


def make(game):
    out = {}
    for name, acts in action_streams().items():
        envs = [ref_env.RefEnv(game, SEED + i) for i in range(N_ENVS)]
        has_tiles = envs[0].tiles().size > 0
        if has_tiles:
            out[name + "_tiles_make"] = np.stack([e.tiles() for e in envs]).astype(np.int8)
        out[name + "_rngpos_make"] = np.array([e.rng_state()[1] for e in envs], np.int32)
        out[name + "_rngcrc_make"] = np.array([crc(e.rng_state()[0]) for e in envs], np.uint32)
        obs0 = np.stack([e.reset() for e in envs])
        if has_tiles:
            out[name + "_tiles_reset"] = np.stack([e.tiles() for e in envs]).astype(np.int8)
        out[name + "_rngpos_reset"] = np.array([e.rng_state()[1] for e in envs], np.int32)
        out[name + "_rngcrc_reset"] = np.array([crc(e.rng_state()[0]) for e in envs], np.uint32)
        crcs = np.zeros((T + 1, N_ENVS), np.uint32)
        crcs[0] = [crc(o) for o in obs0]
        rew = np.zeros((T, N_ENVS), np.float32)
        term = np.zeros((T, N_ENVS), np.bool_)
        frames = {0: obs0}
        for t in range(T):
            obs = []
            for i, e in enumerate(envs):
                o, r, d = e.step(acts[t, i])
                if d:
                    o = e.reset()
                obs.append(o)
                rew[t, i], term[t, i] = r, d
            obs = np.stack(obs)
            crcs[t + 1] = [crc(o) for o in obs]
            if t + 1 in KEEP_FRAMES:
                frames[t + 1] = obs
        out[name + "_actions"] = acts
        out[name + "_reward"] = rew
        out[name + "_terminated"] = term
        out[name + "_obs_crc"] = crcs
        for k, v in frames.items():
            out[name + "_frame%d" % k] = v
        out[name + "_rngpos_end"] = np.array([e.rng_state()[1] for e in envs], np.int32)
        out[name + "_rngcrc_end"] = np.array([crc(e.rng_state()[0]) for e in envs], np.uint32)
    out["seed"] = np.int64(SEED)
    np.savez_compressed(os.path.join(HERE, "%s.npz" % game), **out)
    print(game, "episodes ended:", {n: int(out[n + "_terminated"].sum()) for n in ("uniform", "biased")})


if __name__ == "__main__":
    build_ref.build()
    for g in (sys.argv[1:] or list(build_ref.GAMES)):
        make(g)
