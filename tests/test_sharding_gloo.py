"""N>1 host logic on CPU: two gloo ranks each own a contiguous shard of the env batch (seed = base + global env
index, exactly bench.py's `first_env = rank * envs_per_gpu`), step it with the host-sim of the device source, and
all-gather checksums; the result must equal one process stepping the whole batch. Also exercises the
max-over-ranks reduction bench.py applies to its timings. No collective is on the step path itself."""
import os
import socket
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GAME, PER_RANK, T, SEED = "coinrun", 6, 25, 77


def _actions(world):
    return np.random.RandomState(5).randint(0, 15, size=(T, PER_RANK * world)).astype(np.int32)


def _run_shard(first_env, count, acts):
    from tests.simlib import HostSim
    sim = HostSim(GAME, count, SEED + first_env)       # env i of the shard is seeded SEED + first_env + i
    sim.reset()
    out = []
    for t in range(T):
        o, r, d = sim.step(acts[t, first_env:first_env + count])
        out.append([zlib.crc32(x.tobytes()) & 0xffffffff for x in o] + [int(np.float32(v).view(np.uint32)) for v in r] + [int(v) for v in d])
    sim.close()
    return np.array(out, np.int64)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    acts = _actions(world)
    mine = torch.from_numpy(_run_shard(rank * PER_RANK, PER_RANK, acts))
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)           # bench.py: device time = max over ranks
    if rank == 0:
        q.put((torch.stack(gathered).numpy(), float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_equal_one_batch():
    import torch.multiprocessing as mp
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, tmax = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == float(world)
    whole = _run_shard(0, PER_RANK * world, _actions(world))      # [T, 3 * N]: crc | reward bits | done
    n = PER_RANK * world
    for r in range(world):
        sl = slice(r * PER_RANK, (r + 1) * PER_RANK)
        shard = gathered[r]                                        # [T, 3 * PER_RANK]
        for k in range(3):
            np.testing.assert_array_equal(shard[:, k * PER_RANK:(k + 1) * PER_RANK], whole[:, k * n:(k + 1) * n][:, sl])
