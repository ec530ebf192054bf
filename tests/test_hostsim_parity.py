"""CPU tests: the engine's device SOURCE (compiled for the host by tests/hostsim, single lane)
against the golden vectors generated from the compiled reference, and against the live
reference when oracle/_ref is available. The product path (CUDA) is covered by test_gpu_parity."""
import numpy as np
import pytest

from tests.conftest import IMPLEMENTED, golden
from tests.parity_util import check_against_golden, check_against_oracle
from tests.simlib import HostSim


class SimAdapter:
    def __init__(self, game, n, seed, **kw):
        self.sim = HostSim(game, n, seed, **kw)
        self.n = n

    def reset(self):
        return self.sim.reset()

    def step(self, actions):
        return self.sim.step(actions)

    def fields(self):
        out = {}
        for name, dt in (("mt", np.uint32), ("mti", np.int32), ("tiles", np.uint8)):
            try:
                b, esz, pe = self.sim.field(name)
            except KeyError:
                continue
            out[name] = b.view(dt).reshape(self.n, pe) if pe > 1 else b.view(dt)
        return out


@pytest.mark.parametrize("game", IMPLEMENTED)
@pytest.mark.parametrize("stream", ["uniform", "biased"])
def test_golden(game, stream):
    check_against_golden(lambda n, seed: SimAdapter(game, n, seed), golden(game), stream, read_fields=lambda a: a.fields())


@pytest.mark.parametrize("game", IMPLEMENTED)
def test_live_oracle(game, oracle_available):
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    from oracle import ref_env
    n, T, seed = 8, 300, 77
    rs = np.random.RandomState(5)
    acts = np.where(rs.rand(T, n) < 0.5, rs.randint(0, 15, size=(T, n)), rs.choice([6, 7, 8, 8, 5], size=(T, n))).astype(np.int32)
    refs = [ref_env.RefEnv(game, seed + i) for i in range(n)]
    check_against_oracle(SimAdapter(game, n, seed), refs, acts, tag=game)


@pytest.mark.parametrize("game", IMPLEMENTED)
def test_level_generation_many_seeds(game, oracle_available):
    """Level layouts + the complete MT19937 state for many seeds and consecutive resets."""
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    from oracle import ref_env
    n, seed = 24, 31337
    sim = SimAdapter(game, n, seed)
    refs = [ref_env.RefEnv(game, seed + i) for i in range(n)]
    for rnd in range(4):
        f = sim.fields()
        for i, r in enumerate(refs):
            st, pos = r.rng_state()
            assert pos == f["mti"][i], (rnd, i)
            np.testing.assert_array_equal(st, f["mt"][i])
            rt = r.tiles()
            if rt.size:
                w, h = rt.shape
                np.testing.assert_array_equal((f["tiles"][i, :w * h] & 15).reshape(w, h), rt)
        obs = sim.reset()
        np.testing.assert_array_equal(obs, np.stack([r.reset() for r in refs]))


def test_max_episode_steps_truncates():
    sim = HostSim("maze", 4, 1, max_episode_steps=5)
    sim.reset()
    for t in range(5):
        _, _, term = sim.step(np.full(4, 4, np.int32))   # action 4 = stay
    b, _, _ = sim.field("ep_steps")
    assert (b.view(np.int32) == 0).all()       # truncated at step 5 -> regenerated
    assert not term.any()                      # truncation is not termination
