"""Shared parity checks: an implementation (`host-sim` of the device source, or the CUDA engine
through its C ABI) against golden vectors / the live oracle. Bit-exact everywhere: rewards,
terminated flags, observation pixels, tile maps, MT19937 state."""
import zlib

import numpy as np


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xffffffff


def check_against_golden(impl_factory, g, stream, read_fields=None):
    """impl_factory(num_envs, seed) -> object with reset() -> obs, step(actions) -> (obs, reward, terminated)."""
    acts = g[stream + "_actions"]
    T, N = acts.shape
    seed = int(g["seed"])
    impl = impl_factory(N, seed)
    if read_fields is not None:
        check_level(read_fields(impl), g, stream + "_%s_make")
    obs = impl.reset()
    if read_fields is not None:
        check_level(read_fields(impl), g, stream + "_%s_reset")
    assert [crc(o) for o in obs] == list(g[stream + "_obs_crc"][0]), "reset frame differs"
    np.testing.assert_array_equal(obs, g[stream + "_frame0"])
    for t in range(T):
        obs, rew, term = impl.step(acts[t])
        np.testing.assert_array_equal(rew, g[stream + "_reward"][t], err_msg="reward, step %d" % t)
        np.testing.assert_array_equal(term, g[stream + "_terminated"][t], err_msg="terminated, step %d" % t)
        key = stream + "_frame%d" % (t + 1)
        if key in g:
            np.testing.assert_array_equal(obs, g[key], err_msg="pixels, step %d" % t)
        got = [crc(o) for o in obs]
        assert got == list(g[stream + "_obs_crc"][t + 1]), "observation CRC differs at step %d: envs %s" % (
            t, [i for i in range(N) if got[i] != g[stream + "_obs_crc"][t + 1][i]])
    if read_fields is not None:
        f = read_fields(impl)
        np.testing.assert_array_equal(f["mti"], g[stream + "_rngpos_end"])
        assert [crc(m) for m in f["mt"]] == list(g[stream + "_rngcrc_end"])
    return impl


def check_level(f, g, pattern):
    """Tile map + RNG state after a level generation."""
    np.testing.assert_array_equal(f["mti"], g[pattern % "rngpos"], err_msg="MT19937 position")
    assert [crc(m) for m in f["mt"]] == list(g[pattern % "rngcrc"]), "MT19937 state words differ"
    key = pattern % "tiles"
    if key in g and "tiles" in f:
        ref = g[key]
        n, w, h = ref.shape
        mine = (f["tiles"][:, :w * h] & 15).reshape(n, w, h).astype(np.int8)
        np.testing.assert_array_equal(mine, ref, err_msg="tile map")


def check_against_oracle(impl, refs, actions, tag=""):
    """Lock-step comparison with live reference environments (reset-on-terminate)."""
    obs = impl.reset()
    robs = np.stack([r.reset() for r in refs])
    np.testing.assert_array_equal(obs, robs, err_msg=tag + " reset frame")
    episodes = 0
    for t in range(actions.shape[0]):
        obs, rew, term = impl.step(actions[t])
        ro, rr, rt = [], [], []
        for i, r in enumerate(refs):
            o, w, d = r.step(actions[t, i])
            if d:
                o = r.reset()
            ro.append(o); rr.append(w); rt.append(d)
        episodes += int(np.sum(rt))
        np.testing.assert_array_equal(rew, np.array(rr, np.float32), err_msg="%s reward, step %d" % (tag, t))
        np.testing.assert_array_equal(term, np.array(rt), err_msg="%s terminated, step %d" % (tag, t))
        np.testing.assert_array_equal(obs, np.stack(ro), err_msg="%s pixels, step %d" % (tag, t))
    return episodes
