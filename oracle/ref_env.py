"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes driver for the compiled reference games in oracle/_ref/ (one loaded COPY of the library
per environment, because the reference keeps every piece of state in process globals —
/root/reference/games/coinrun/coinrun.cpp:15-57). Mirrors the call sequence of the reference's
own wrapper (cenv/cenv.py:206, 283, 338): cenv_make(options) / cenv_reset / cenv_step.
"""
import ctypes
import os
import shutil
import tempfile

import numpy as np

from . import build_ref

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DEFAULT_BLOB = os.path.join(ROOT, "procgen2_b200", "data", "assets.bin")


class _Value(ctypes.Union):
    _fields_ = [("i", ctypes.c_int32), ("f", ctypes.c_float), ("d", ctypes.c_double), ("b", ctypes.c_uint8)]


class _Buffer(ctypes.Union):
    _fields_ = [("i", ctypes.POINTER(ctypes.c_int32)), ("f", ctypes.POINTER(ctypes.c_float)),
                ("d", ctypes.POINTER(ctypes.c_double)), ("b", ctypes.POINTER(ctypes.c_uint8))]


class KeyValue(ctypes.Structure):
    _fields_ = [("key", ctypes.c_char_p), ("value_type", ctypes.c_int32), ("value_buffer_size", ctypes.c_int32),
                ("value_buffer", _Buffer)]


class Option(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("value_type", ctypes.c_int32), ("value", _Value)]


class StepData(ctypes.Structure):
    _fields_ = [("observations_size", ctypes.c_int32), ("observations", ctypes.POINTER(KeyValue)),
                ("reward", _Value), ("terminated", ctypes.c_bool), ("truncated", ctypes.c_bool),
                ("infos_size", ctypes.c_int32), ("infos", ctypes.POINTER(KeyValue))]


class RenderData(ctypes.Structure):
    _fields_ = [("value_type", ctypes.c_int32), ("value_buffer_width", ctypes.c_int32), ("value_buffer_height", ctypes.c_int32),
                ("value_buffer_channels", ctypes.c_int32), ("value_buffer", _Buffer)]


class ResetData(ctypes.Structure):
    _fields_ = [("observations_size", ctypes.c_int32), ("observations", ctypes.POINTER(KeyValue)),
                ("infos_size", ctypes.c_int32), ("infos", ctypes.POINTER(KeyValue))]


_tmpdir = None
_assets_lib = None


def available():
    return all(os.path.exists(build_ref.lib_path(g)) for g in build_ref.GAMES)


def _prepare():
    global _tmpdir, _assets_lib
    if _tmpdir is None:
        _tmpdir = tempfile.mkdtemp(prefix="pg2o_")
        os.environ.setdefault("PG2_ASSETS", DEFAULT_BLOB)
        _assets_lib = ctypes.CDLL(os.path.join(build_ref.OUT, "libpg2o_assets.so"), mode=ctypes.RTLD_GLOBAL)
    return _tmpdir


class RefEnv:
    """One reference environment (one private copy of lib<Game>.so)."""
    _count = 0

    def __init__(self, game, seed, width=None, height=None, easy_mode=None, mode=None):
        d = _prepare()
        RefEnv._count += 1
        self.game = game
        self.path = os.path.join(d, "%s_%d.so" % (game, RefEnv._count))
        shutil.copyfile(build_ref.lib_path(game), self.path)
        self.lib = ctypes.CDLL(self.path)
        self.lib.cenv_make.argtypes = [ctypes.c_char_p, ctypes.POINTER(Option), ctypes.c_int32]
        self.lib.cenv_reset.argtypes = [ctypes.POINTER(Option), ctypes.c_int32]
        self.lib.cenv_step.argtypes = [ctypes.POINTER(KeyValue), ctypes.c_int32]
        if mode is not None and game in ("coinrun", "climber"):   # their Config has a bool easy_mode instead of the enum
            easy_mode, mode = (int(mode) == 0), None
        if easy_mode is not None:   # compile-time Config::easy_mode of the generator (coinrun, climber), set through the probe
            self.probe("pg2o_set_easy_mode", None, [ctypes.c_int])(1 if easy_mode else 0)
        late_mode = mode is not None and game == "bossfight"   # its Config lives in a system cenv_make creates; only update() reads it
        if mode is not None and not late_mode:   # compile-time Config::mode (maze, chaser, jumper, caveflyer): 0 easy, 1 hard, 2 memory / extreme
            self.probe("pg2o_set_mode", None, [ctypes.c_int])(int(mode))
        opts = [(b"seed", int(seed))] + ([(b"width", int(width))] if width else []) + ([(b"height", int(height))] if height else [])
        arr = (Option * len(opts))()
        for i, (k, v) in enumerate(opts):
            arr[i].name, arr[i].value_type, arr[i].value = k, 0, _Value(i=v)
        assert self.lib.cenv_make(b"", arr, len(opts)) == 0
        if late_mode:
            self.probe("pg2o_set_mode", None, [ctypes.c_int])(int(mode))
        self.render_data = RenderData.in_dll(self.lib, "render_data")
        self.step_data = StepData.in_dll(self.lib, "step_data")
        self.reset_data = ResetData.in_dll(self.lib, "reset_data")
        self._a = ctypes.c_int32(0)
        self._kv = KeyValue(b"action", 0, 1, _Buffer(i=ctypes.pointer(self._a)))
        self._kv_ref = ctypes.byref(self._kv)
        os.unlink(self.path)  # mapping stays valid

    def close(self):
        """Unload this environment's private library copy (many-seed sweeps would otherwise keep thousands mapped)."""
        lib, self.lib = self.lib, None
        if lib is not None:
            import _ctypes
            self.step_data = self.reset_data = self.render_data = None
            _ctypes.dlclose(lib._handle)

    def _obs(self, kv):
        n = kv.value_buffer_size
        return np.ctypeslib.as_array(kv.value_buffer.b, shape=(n,)).reshape(64, 64, 3).copy()

    def render(self):
        """cenv_render: the human-mode frame [height, width, 3]."""
        self.lib.cenv_render()
        rd = self.render_data
        n = rd.value_buffer_width * rd.value_buffer_height * rd.value_buffer_channels
        return np.ctypeslib.as_array(rd.value_buffer.b, shape=(n,)).reshape(rd.value_buffer_height, rd.value_buffer_width, 3).copy()

    def reset(self, seed=None):
        if seed is None:
            assert self.lib.cenv_reset(None, 0) == 0
        else:
            opt = Option(b"seed", 0, _Value(i=int(seed)))
            assert self.lib.cenv_reset(ctypes.byref(opt), 1) == 0
        return self._obs(self.reset_data.observations[0])

    def step(self, action):
        self._a.value = int(action)
        assert self.lib.cenv_step(ctypes.byref(self._kv), 1) == 0
        sd = self.step_data
        return self._obs(sd.observations[0]), float(sd.reward.f), bool(sd.terminated)

    def raw_step(self, action):
        """cenv_step without copying the observation out (timing loops). Returns terminated."""
        self._a.value = int(action)
        self.lib.cenv_step(self._kv_ref, 1)
        return bool(self.step_data.terminated)

    def raw_reset(self):
        self.lib.cenv_reset(None, 0)

    # ---- probe accessors (oracle/probe/probe_<game>.cpp), present only when compiled in ----
    def probe(self, name, restype=ctypes.c_int, argtypes=()):
        fn = getattr(self.lib, name)
        fn.restype = restype
        fn.argtypes = list(argtypes)
        return fn

    def rng_state(self):
        """(624 state words, position) of the reference's global std::mt19937."""
        buf = (ctypes.c_uint32 * 625)()
        self.probe("pg2o_rng_state", None, [ctypes.POINTER(ctypes.c_uint32)])(buf)
        a = np.frombuffer(buf, dtype=np.uint32).copy()
        return a[:624], int(a[624])

    def tiles(self):
        dims = (ctypes.c_int * 2)()
        self.probe("pg2o_tile_dims", None, [ctypes.POINTER(ctypes.c_int)])(dims)
        w, h = dims[0], dims[1]
        buf = (ctypes.c_int32 * (w * h))()
        self.probe("pg2o_tiles", None, [ctypes.POINTER(ctypes.c_int32)])(buf)
        return np.frombuffer(buf, dtype=np.int32).reshape(w, h).copy()  # [x][y], column-major like the reference

    def floats(self, n=64):
        """Game-specific float state vector (see the probe source for the layout)."""
        buf = (ctypes.c_float * n)()
        k = self.probe("pg2o_floats", ctypes.c_int, [ctypes.POINTER(ctypes.c_float), ctypes.c_int])(buf, n)
        return np.frombuffer(buf, dtype=np.float32)[:k].copy()


def rollout(game, seed, actions, auto_reset=True, with_state=False):
    """Reference trajectory with the engine's auto-reset convention:
    make(seed) -> reset() -> for a in actions: step(a); if terminated: obs = reset()."""
    env = RefEnv(game, seed)
    obs0 = env.reset()
    obs, rew, term = [], [], []
    states = []
    for a in actions:
        o, r, t = env.step(int(a))
        if t and auto_reset:
            o = env.reset()
        obs.append(o)
        rew.append(r)
        term.append(t)
        if with_state:
            states.append(env.floats())
    out = dict(obs0=obs0, obs=np.stack(obs) if obs else np.zeros((0, 64, 64, 3), np.uint8),
               reward=np.array(rew, np.float32), terminated=np.array(term, np.bool_))
    if with_state:
        out["state"] = states
    return out
