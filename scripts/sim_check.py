"""Dev script: host-sim of the device source vs the live oracle, long runs with a chosen action mix.
    python scripts/sim_check.py <game> [envs] [steps] [seed] [mix]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.simlib import HostSim
from oracle import ref_env

MIXES = {
    "uniform": lambda rs, T, n: rs.randint(0, 15, size=(T, n)),
    "fire": lambda rs, T, n: np.where(rs.rand(T, n) < 0.6, 9, rs.randint(0, 15, size=(T, n))),
    "right": lambda rs, T, n: np.where(rs.rand(T, n) < 0.5, rs.randint(0, 15, size=(T, n)), rs.choice([6, 7, 8, 8, 5], size=(T, n))),
    "chase": lambda rs, T, n: np.where(rs.rand(T, n) < 0.3, rs.randint(0, 15, size=(T, n)), np.repeat(rs.choice([1, 3, 5, 7], size=(T // 8 + 1, n)), 8, axis=0)[:T]),
    "up": lambda rs, T, n: np.where(rs.rand(T, n) < 0.5, rs.randint(0, 15, size=(T, n)), rs.choice([5, 8, 2, 5, 7], size=(T, n))),
}

def main():
    game = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    seed = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    mix = sys.argv[5] if len(sys.argv) > 5 else "uniform"
    rs = np.random.RandomState(seed)
    acts = MIXES[mix](rs, T, n).astype(np.int32)
    sim = HostSim(game, n, seed)
    refs = [ref_env.RefEnv(game, seed + i) for i in range(n)]
    o = sim.reset()
    ro = np.stack([r.reset() for r in refs])
    assert np.array_equal(o, ro), ("reset frame", (o != ro).any((1, 2, 3)).nonzero())
    episodes = 0
    for t in range(T):
        o, r, d = sim.step(acts[t])
        ro, rr, rd = [], [], []
        for i, e in enumerate(refs):
            oo, w, dd = e.step(acts[t, i])
            if dd:
                oo = e.reset()
            ro.append(oo); rr.append(w); rd.append(dd)
        ro = np.stack(ro); rr = np.array(rr, np.float32); rd = np.array(rd)
        episodes += int(rd.sum())
        if not np.array_equal(r, rr) or not np.array_equal(d, rd):
            print("step", t, "reward/term mismatch", r, rr, d, rd); sys.exit(1)
        if not np.array_equal(o, ro):
            bad = (o != ro).any((1, 2, 3)).nonzero()[0]
            print("step", t, "pixel mismatch envs", bad, "npix", [(o[b] != ro[b]).any(-1).sum() for b in bad]); 
            np.save("/tmp/mismatch_mine.npy", o[bad[0]]); np.save("/tmp/mismatch_ref.npy", ro[bad[0]])
            sys.exit(1)
    print("OK", game, "envs", n, "steps", T, "episodes", episodes, "mix", mix)

main()
