"""Dev script (run under gpurun): parity of the CUDA engine vs oracle/_ref + quick timing."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from procgen2_b200.engine import BatchedEnv
from oracle import ref_env

def parity(game, N, T, seed):
    rs = np.random.RandomState(1)
    acts = rs.randint(0, 15, size=(T, N)).astype(np.int32)
    env = BatchedEnv(game, N, seed=seed)
    env.reset()
    o, _, _, _ = env.fetch()
    refs = [ref_env.RefEnv(game, seed + i) for i in range(N)]
    ro = np.stack([r.reset() for r in refs])
    bad = int((o != ro).any())
    if bad: print(game, "reset frame mismatch", (o != ro).any((1,2,3)).nonzero())
    for t in range(T):
        env.step(acts[t])
        o, r, te, _ = env.fetch()
        ro, rr, rt = [], [], []
        for i, e in enumerate(refs):
            oo, rw, tm = e.step(acts[t, i])
            if tm: oo = e.reset()
            ro.append(oo); rr.append(rw); rt.append(tm)
        ro = np.stack(ro)
        if (o != ro).any(): print(game, "step", t, "obs mismatch envs", (o != ro).any((1,2,3)).nonzero()[0][:8]); bad += 1
        if not np.array_equal(r, np.array(rr, np.float32)): print(game, "step", t, "reward mismatch"); bad += 1
        if not np.array_equal(te, np.array(rt)): print(game, "step", t, "term mismatch"); bad += 1
        if bad > 5: break
    print("PARITY", game, "N", N, "T", T, "bad", bad)
    env.close()
    return bad

def timing(game, N, steps=200):
    import torch
    env = BatchedEnv(game, N, seed=0)
    env.reset()
    acts = torch.randint(0, 15, (steps, N), dtype=torch.int32, device="cuda")
    for t in range(20): env.step_torch(acts[t])
    env.sync()
    t0 = time.time()
    for t in range(steps): env.step_torch(acts[t])
    env.sync()
    dt = time.time() - t0
    print("TIMING", game, "N", N, "steps/s %.3e" % (N * steps / dt), "ms/step %.3f" % (1e3 * dt / steps))
    env.close()

if __name__ == "__main__":
    games = sys.argv[1].split(",")
    bad = 0
    for g in games:
        bad += parity(g, 32, int(sys.argv[2]) if len(sys.argv) > 2 else 300, 100)
    for g in games:
        for n in (4096, 32768):
            timing(g, n)
    sys.exit(1 if bad else 0)
