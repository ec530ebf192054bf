"""Dev tool (GPU box): per-phase wall cycles of k_render from a -DPG2_PHASE_TIMERS build (PG2_ENGINE_LIB=exp_libs/phase.so).
usage: PG2_ENGINE_LIB=... python scripts/phase_probe.py game envs [steps]"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from procgen2_b200.engine import BatchedEnv, load_library
game, n = sys.argv[1], int(sys.argv[2]); steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
env = BatchedEnv(game, n, seed=0)
env.reset()
rs = np.random.RandomState(0)
for t in range(10): env.step(rs.randint(0, 15, size=n).astype(np.int32))
env.sync()
L = load_library()
L.pg2_debug_phases.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
out = (ctypes.c_uint64 * 8)()
assert L.pg2_debug_phases(env._h, out) == 0, "not a PG2_PHASE_TIMERS build"
for t in range(steps): env.step(rs.randint(0, 15, size=n).astype(np.int32))
env.sync()
L.pg2_debug_phases(env._h, out)
fr = max(out[0], 1)
names = ["frames", "ticket+begin", "build_frame", "finalize", "rasterise"]
print(game, n, "frames", out[0], " cycles/frame:", {names[k]: round(out[k] / fr) for k in range(1, 5)}, "sum", round(sum(out[1:5]) / fr))
