#!/usr/bin/env python
"""Benchmark of the hot path: env-steps/s including the 64x64x3 render (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--game coinrun] [--envs-per-gpu 4096]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # reference C++ engine (oracle/_ref) on the host cores

Workload (config.workload): BASELINE.json configs[1] — coinrun, 4096 envs per GPU, uniform-random
actions with on-device auto-reset and per-episode level regeneration. A "step" is one cenv_step
of every env of the batch. Weak scaling: every rank owns `envs_per_gpu` envs (a contiguous slice of
the global env index space, seeds = base + global index), no collective on the step path.

value    device-resident throughput: actions already in HBM, observations stay in HBM; every step is
         timed with CUDA events on the engine's stream and an L2 flush (256 MiB memset) runs
         between steps, outside the event pairs.
e2e      same metric through the host-buffer C ABI: every step copies its actions H2D from pinned memory and
         its observations / rewards / terminated flags D2H into pinned memory, inside the timed region.
         pg2_step_pipelined overlaps the D2H of step t-1 with the kernels of step t (depth-1 pipeline, two
         alternating host buffer sets); e2e.sequential is the strictly serial pg2_step + pg2_fetch pair.
roofline dominant kernel (k_render) against the measured HBM copy bandwidth in MEASURED_PEAKS.json,
         algorithmic bytes = 12 297 B per env-step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_ENV_STEP = 12288 + 4 + 4 + 1   # obs write + action read + reward write + done write
METRIC = "env-steps/sec incl. 64x64 RGB render"
UNIT = "env-steps/s"
BASE_SEED = 0


def measured_traffic(game, envs):
    """dram bytes per launch of the dominant kernel, from the committed ncu capture (profiles/traffic.json)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("%s@%d" % (game, envs))
        return (t["dram_bytes_read"] + t["dram_bytes_write"]) if t else None
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's own C++ engine (oracle/_ref) on the host cores

def _ref_proc(conn, game, seeds, action_seed):
    """Worker process: owns len(seeds) reference environments (one library copy each)."""
    sys.path.insert(0, ROOT)
    from oracle import ref_env
    envs = [ref_env.RefEnv(game, s) for s in seeds]
    for e in envs:
        e.reset()
    rs = np.random.RandomState(action_seed)
    conn.send("ready")
    while True:
        n = conn.recv()
        if n <= 0:
            break
        acts = rs.randint(0, 15, size=(n, len(envs)))
        t0 = time.perf_counter()
        for t in range(n):
            for i, e in enumerate(envs):
                if e.raw_step(acts[t, i]):
                    e.raw_reset()
        conn.send(time.perf_counter() - t0)


class ReferencePool:
    """The reference C++ engine (oracle/_ref) spread over the host cores: `procs` processes x
    `envs_per_proc` environments, env j seeded BASE_SEED + j, uniform-random actions, reset on
    terminate (game_test.py:38-40). run(n) advances every env n steps and returns env-steps/s."""

    def __init__(self, game, procs, envs_per_proc):
        import multiprocessing as mp
        ctx = mp.get_context("fork")
        self.procs, self.envs_per_proc, self.workers = procs, envs_per_proc, []
        for p in range(procs):
            parent, child = ctx.Pipe()
            seeds = [BASE_SEED + p * envs_per_proc + i for i in range(envs_per_proc)]
            w = ctx.Process(target=_ref_proc, args=(child, game, seeds, 1234 + p), daemon=True)
            w.start()
            self.workers.append((w, parent))
        for _, c in self.workers:
            assert c.recv() == "ready"

    def run(self, n):
        t0 = time.perf_counter()
        for _, c in self.workers:
            c.send(n)
        for _, c in self.workers:
            c.recv()
        wall = time.perf_counter() - t0
        return self.procs * self.envs_per_proc * n / wall, wall

    def close(self):
        for w, c in self.workers:
            c.send(0)
        for w, _ in self.workers:
            w.join(timeout=5)


def reference_available():
    from oracle import ref_env
    return ref_env.available()


def run_reference(a):
    if not reference_available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (reference sources absent)"}))
        return
    procs = os.cpu_count() or 1
    envs_per_proc = 8
    pool = ReferencePool(a.game, procs, envs_per_proc)
    for _ in range(a.warmup):
        pool.run(a.ref_inner)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        pool.run(a.ref_inner)
    wall = time.perf_counter() - t0
    pool.close()
    v = procs * envs_per_proc * a.ref_inner * a.steps / wall
    sample = "%d procs x %d envs; one timed step = %d cenv_step per env (%s, seeds %d.., uniform actions, reset on terminate)" % (
        procs, envs_per_proc, a.ref_inner, a.game, BASE_SEED)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * wall / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(a, a.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": procs, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(a, world):
    return {"workload": "BASELINE.json configs[1]: %s, %d envs per GPU, 64x64x3 uint8 obs, uniform-random actions, "
                        "auto-reset with per-episode level regeneration" % (a.game, a.envs_per_gpu),
            "game": a.game, "envs_per_gpu": a.envs_per_gpu, "global_envs": a.envs_per_gpu * world, "parallelism": "env-sharded x%d, no collective" % world,
            "l2": "flushed between timed steps (256 MiB memset outside the event pairs)", "base_seed": BASE_SEED,
            "max_episode_steps": a.max_episode_steps}


# ---------------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--game", default="coinrun")
    ap.add_argument("--envs-per-gpu", type=int, default=4096)
    ap.add_argument("--max-episode-steps", type=int, default=0,
                    help="truncate episodes (engine extension; BASELINE configs[4] stresses level generation with short episodes)")
    ap.add_argument("--ref-inner", type=int, default=50, help="env steps per env per timed sample of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if a.impl == "reference":
        if rank == 0:
            run_reference(a)
        return

    import torch
    import torch.distributed as dist
    from procgen2_b200.engine import BatchedEnv

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "WARN"     # keep stdout to the one JSON line (no "NCCL version" banner)
        dist.init_process_group("nccl", device_id=dev)

    N = a.envs_per_gpu
    env = BatchedEnv(a.game, N, seed=BASE_SEED, device=local_rank, first_env=rank * N, max_episode_steps=a.max_episode_steps)
    env.reset()
    env.sync()
    total_steps = a.warmup + a.steps
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    # synthetic uniform-random action stream, generated on device ahead of the timed region
    pool = min(total_steps, 512)
    actions = torch.randint(0, 15, (pool, N), dtype=torch.int32, device=dev, generator=gen)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.ExternalStream(env.stream_ptr, device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident arm ------------------------------------------------------------------
    with torch.cuda.stream(stream):
        for t in range(a.warmup):
            env.step_torch(actions[t % pool])
            flush.zero_()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = env.kernel_launches
    env.profile(True)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
    barrier()
    wall0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for t in range(a.steps):
            starts[t].record(stream)
            env.step_torch(actions[(a.warmup + t) % pool])
            ends[t].record(stream)
            flush.zero_()
    barrier()
    wall = time.perf_counter() - wall0
    prof, prof_steps = env.profile_read()
    env.profile(False)
    launches = env.kernel_launches - launches0
    dev_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    t_dev = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dev_ms = float(t_dev.item())

    # ---- end-to-end arm: host actions in, host observations out -----------------------------------
    # Through the C ABI with HOST buffers: every step copies its actions H2D from pinned memory and its
    # observations / rewards / terminated flags D2H into pinned memory. pg2_step_pipelined overlaps the D2H of
    # step t-1 with the kernels of step t (two alternating host buffer sets); the strictly sequential
    # pg2_step + pg2_fetch pair is timed as well (e2e.sequential).
    host_actions = actions.cpu().numpy()
    bufs = []
    for _ in range(2):
        o = torch.empty((N, 64, 64, 3), dtype=torch.uint8).pin_memory()
        r = torch.empty(N, dtype=torch.float32).pin_memory()
        d = torch.empty(N, dtype=torch.uint8).pin_memory()
        bufs.append((o.numpy(), r.numpy(), d.numpy(), (o, r, d)))
    e2e_steps = max(10, min(a.steps, 100))
    for t in range(3):
        env.step(host_actions[t % pool])
        env.fetch_into(*bufs[0][:3])
    barrier()
    e0 = time.perf_counter()
    for t in range(e2e_steps):
        env.step(host_actions[t % pool])
        env.fetch_into(*bufs[0][:3])
    barrier()
    seq_s = time.perf_counter() - e0
    for t in range(4):
        env.step_pipelined(host_actions[t % pool], *bufs[t & 1][:3])
    env.flush()
    barrier()
    e0 = time.perf_counter()
    for t in range(e2e_steps):
        env.step_pipelined(host_actions[t % pool], *bufs[t & 1][:3])
    env.flush()
    barrier()
    e2e_s = time.perf_counter() - e0
    t_e2e = torch.tensor([e2e_s, seq_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_s, seq_s = float(t_e2e[0].item()), float(t_e2e[1].item())
    sampler.stop_flag = True
    sampler.join(timeout=2)

    if rank == 0:
        value = N * world * a.steps / (dev_ms * 1e-3)
        hbm_gbs, peak_src = peaks()
        render_ms = prof["render"] / max(prof_steps, 1)
        achieved = ALG_BYTES_PER_ENV_STEP * N / (render_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(a, world),
            "e2e": {"value": N * world * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 4 * N, "d2h_bytes_per_step": N * (12288 + 4 + 1),
                    "steps": e2e_steps, "api": "pg2_step_pipelined(host actions -> pinned host obs/reward/terminated), depth-1 pipeline",
                    "sequential": N * world * e2e_steps / seq_s},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs, "traffic": measured_traffic(a.game, N),
                         "kernel": "k_render<%s>" % a.game, "kernel_ms": render_ms, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": ALG_BYTES_PER_ENV_STEP * N},
            "kernel_ms_per_step": {k: v / max(prof_steps, 1) for k, v in prof.items()},
            "clocks": sampler.summary(), "wall_s_timed_region": wall,
        }
        if not a.no_cpu_baseline and world == 1:
            procs = os.cpu_count() or 1
            if reference_available():
                pool = ReferencePool(a.game, procs, 8)
                pool.run(a.ref_inner)
                n_cpu = max(a.ref_inner, 4000)   # ~1 s wall on every host core = 10-30 s of CPU work
                v, wall_cpu = pool.run(n_cpu)
                pool.close()
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": procs, "kind": "reference",
                                        "sample": "%d procs x 8 envs x %d steps in %.1f s (%s; reference C++ engine, unmodified, + canonical "
                                                  "CPU rasteriser = oracle/_ref)" % (procs, n_cpu, wall_cpu, a.game)}
        print(json.dumps(line))
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
