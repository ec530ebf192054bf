"""Long parity soak of the host-sim build (the device headers compiled with g++, tests/hostsim) against the live reference
(oracle/_ref): every game in every distribution mode, many envs, thousands of steps with auto-reset — looks for rare
divergences the bounded test suite cannot reach. CPU only.   usage: python scripts/soak_hostsim.py [envs] [steps] [seed]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle import ref_env                                  # noqa: E402
from tests.test_hostsim_parity import SimAdapter            # noqa: E402

CASES = [("maze", None), ("maze", 0), ("maze", 2), ("coinrun", None), ("bossfight", None), ("bossfight", 0), ("chaser", None),
         ("chaser", 1), ("chaser", 2), ("climber", None), ("caveflyer", None), ("caveflyer", 0), ("caveflyer", 2),
         ("jumper", None), ("jumper", 0), ("jumper", 2)]


def soak(game, mode, n, T, seed):
    rs = np.random.RandomState(seed)
    # a mix of uniform actions and runs of one direction (so agents actually travel)
    acts = rs.randint(0, 15, size=(T, n)).astype(np.int32)
    run = rs.choice([1, 3, 5, 7, 6, 8, 2, 0], size=(T // 16 + 1, n))
    hold = rs.rand(T, n) < 0.6
    acts = np.where(hold, np.repeat(run, 16, axis=0)[:T], acts).astype(np.int32)
    sim = SimAdapter(game, n, seed, distribution_mode=-1 if mode is None else mode)
    refs = [ref_env.RefEnv(game, seed + i, mode=mode) for i in range(n)]
    np.testing.assert_array_equal(sim.reset(), np.stack([r.reset() for r in refs]))
    episodes, reward = 0, 0.0
    for t in range(T):
        o, rw, d = sim.step(acts[t])
        for i, r in enumerate(refs):
            oo, w, dd = r.step(acts[t, i])
            if dd:
                oo = r.reset(); episodes += 1
            reward += w
            assert w == rw[i] and dd == d[i], "%s mode %s step %d env %d: reward %r vs %r, done %r vs %r" % (game, mode, t, i, rw[i], w, d[i], dd)
            if not np.array_equal(o[i], oo):
                raise AssertionError("%s mode %s step %d env %d (seed %d): %d pixels differ" % (game, mode, t, i, seed + i, int((o[i] != oo).any(axis=2).sum())))
    f = sim.fields()
    for i, r in enumerate(refs):
        st, pos = r.rng_state()
        assert pos == f["mti"][i] and np.array_equal(st, f["mt"][i]), "%s mode %s: RNG state of env %d" % (game, mode, i)
        r.close()
    return episodes, reward


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 424242
    for game, mode in CASES:
        t0 = time.time()
        ep, rw = soak(game, mode, n, T, seed)
        print("%-10s mode %-4s ok: %d envs x %d steps, %d episodes, reward %.1f, %.0f s" % (game, mode, n, T, ep, rw, time.time() - t0), flush=True)
