"""Host-side mirror of the reference's Python binding for the cenv ABI.

The reference wrapper is /root/reference/cenv/cenv.py (`class CEnv(gym.Env)`, :152-380): it
dlopens one environment library, mirrors the structs of cenv.h with ctypes (:62-111), and copies
every buffer the library exposes into fresh numpy arrays. This module offers the same class name,
constructor arguments, method names, return tuples and error behaviour (any non-zero return code
raises ``Exception("Non-zero error code!")``), so `CEnv("lib/libCoinRun.so", options={"seed": 3})`
works identically against the reference's library or this project's GPU-backed drop-in — and the
reference's own unmodified cenv.py works against this project's libraries too.

Differences, all additive:
* gymnasium is optional (it is not installed in the build image): without it the spaces are small
  stand-in objects with the same attributes (`low`/`high`/`nvec`/`shape`/`dtype`).
* `step` accepts numpy arrays (`type(action) is np.array` in the reference can never be true,
  cenv/cenv.py:259): an int32 array of length num_envs steps a batched library.
* batched libraries (make option ``num_envs`` > 1) return the observation as one flat uint8 array
  of num_envs*12288 values and per-env reward/terminated/truncated in `info`.
"""
import ctypes
from ctypes import POINTER, Structure, Union, c_bool, c_char_p, c_double, c_float, c_int32, c_ubyte
from typing import Any, Dict, Optional

import numpy as np

try:  # pragma: no cover - depends on the environment
    import gymnasium as gym
    _Env = gym.Env
except Exception:  # gymnasium absent: tiny stand-ins
    gym = None

    class _Env:  # noqa: D401
        pass

CENV_VALUE_TYPE_INT, CENV_VALUE_TYPE_FLOAT, CENV_VALUE_TYPE_DOUBLE, CENV_VALUE_TYPE_BYTE = 0, 1, 2, 3
CENV_VALUE_TYPE_BOX, CENV_VALUE_TYPE_MULTI_DISCRETE = 4, 5

_NUMPY_OF_TAG = [np.int32, np.float32, np.float64, np.uint8, np.float32, np.int32]
_TAG_OF_NUMPY = {np.dtype("int32"): 0, np.dtype("float32"): 1, np.dtype("float64"): 2, np.dtype("uint8"): 3}
_TAG_OF_PYTHON = {int: CENV_VALUE_TYPE_INT, float: CENV_VALUE_TYPE_DOUBLE}   # cenv/cenv.py:39-42


class CEnv_Value(Union):
    _fields_ = [("i", c_int32), ("f", c_float), ("d", c_double), ("b", c_ubyte)]


class CEnv_Value_Buffer(Union):
    _fields_ = [("i", POINTER(c_int32)), ("f", POINTER(c_float)), ("d", POINTER(c_double)), ("b", POINTER(c_ubyte))]


class CEnv_Key_Value(Structure):
    _fields_ = [("key", c_char_p), ("value_type", c_int32), ("value_buffer_size", c_int32), ("value_buffer", CEnv_Value_Buffer)]


class CEnv_Option(Structure):
    _fields_ = [("name", c_char_p), ("value_type", c_int32), ("value", CEnv_Value)]


class CEnv_Make_Data(Structure):
    _fields_ = [("observation_spaces_size", c_int32), ("observation_spaces", POINTER(CEnv_Key_Value)),
                ("action_spaces_size", c_int32), ("action_spaces", POINTER(CEnv_Key_Value))]


class CEnv_Reset_Data(Structure):
    _fields_ = [("observations_size", c_int32), ("observations", POINTER(CEnv_Key_Value)),
                ("infos_size", c_int32), ("infos", POINTER(CEnv_Key_Value))]


class CEnv_Step_Data(Structure):
    _fields_ = [("observations_size", c_int32), ("observations", POINTER(CEnv_Key_Value)), ("reward", CEnv_Value),
                ("terminated", c_bool), ("truncated", c_bool), ("infos_size", c_int32), ("infos", POINTER(CEnv_Key_Value))]


class CEnv_Render_Data(Structure):
    _fields_ = [("value_type", c_int32), ("value_buffer_width", c_int32), ("value_buffer_height", c_int32),
                ("value_buffer_channels", c_int32), ("value_buffer", CEnv_Value_Buffer)]


class _Box:
    def __init__(self, low, high):
        self.low, self.high, self.shape, self.dtype = low, high, low.shape, low.dtype


class _MultiDiscrete:
    def __init__(self, nvec):
        self.nvec, self.shape, self.dtype = nvec, nvec.shape, nvec.dtype


def _copy_out(kv):
    """Library-owned buffer -> fresh numpy array (the caller must copy, USAGE_GUIDE.md:71)."""
    dtype = np.dtype(_NUMPY_OF_TAG[int(kv.value_type)])
    n = int(kv.value_buffer_size)
    raw = ctypes.cast(kv.value_buffer.b, POINTER(c_ubyte * (n * dtype.itemsize))).contents
    return np.frombuffer(raw, dtype=dtype, count=n).copy()


def _options(options):
    if not options:
        return None, 0, []
    arr = (CEnv_Option * len(options))()
    keep = []
    for i, (k, v) in enumerate(options.items()):
        tag = _TAG_OF_PYTHON[type(v)]
        name = k.encode("ascii")
        keep.append(name)
        arr[i].name = name
        arr[i].value_type = tag
        arr[i].value = CEnv_Value(i=int(v)) if tag == CENV_VALUE_TYPE_INT else CEnv_Value(d=float(v))
    return arr, len(options), keep


class CEnv(_Env):
    metadata = {"render_modes": ["human", "single_rgb_array"]}

    def __init__(self, lib_file_path: str, render_mode: Optional[str] = None, options: Optional[Dict[str, Any]] = None):
        self.lib = ctypes.CDLL(lib_file_path)
        self.lib.cenv_get_env_version.restype = c_int32
        self.lib.cenv_make.argtypes = [c_char_p, POINTER(CEnv_Option), c_int32]
        self.lib.cenv_make.restype = c_int32
        self.lib.cenv_reset.argtypes = [POINTER(CEnv_Option), c_int32]
        self.lib.cenv_reset.restype = c_int32
        self.lib.cenv_step.argtypes = [POINTER(CEnv_Key_Value), c_int32]
        self.lib.cenv_step.restype = c_int32
        self.lib.cenv_render.restype = c_int32
        self.lib.cenv_close.restype = None
        self.c_make_data = CEnv_Make_Data.in_dll(self.lib, "make_data")
        self.c_reset_data = CEnv_Reset_Data.in_dll(self.lib, "reset_data")
        self.c_step_data = CEnv_Step_Data.in_dll(self.lib, "step_data")
        self.c_render_data = CEnv_Render_Data.in_dll(self.lib, "render_data")

        c_options, n, _keep = _options(options)
        mode = b"" if render_mode is None else render_mode.encode("ascii")
        if self.lib.cenv_make(mode, c_options, n) != 0:
            raise Exception("Non-zero error code!")

        def spaces(count, ptr):
            out = {}
            for i in range(count):
                arr = _copy_out(ptr[i])
                if int(ptr[i].value_type) == CENV_VALUE_TYPE_MULTI_DISCRETE:
                    sp = gym.spaces.MultiDiscrete(arr) if gym else _MultiDiscrete(arr)
                else:
                    lo, hi = arr[:len(arr) // 2], arr[len(arr) // 2:]
                    sp = gym.spaces.Box(lo, hi) if gym else _Box(lo, hi)
                out[ptr[i].key.decode()] = sp
            return out

        self.observation_space = spaces(self.c_make_data.observation_spaces_size, self.c_make_data.observation_spaces)
        self.action_space = spaces(self.c_make_data.action_spaces_size, self.c_make_data.action_spaces)

    @staticmethod
    def _collect(count, ptr):
        return {ptr[i].key.decode(): _copy_out(ptr[i]) for i in range(count)}

    def step(self, action):
        keep = []
        if isinstance(action, (int, np.integer)):
            scalar = c_int32(int(action))
            kv = CEnv_Key_Value(b"action", CENV_VALUE_TYPE_INT, 1, CEnv_Value_Buffer(i=ctypes.pointer(scalar)))
            c_actions, n = ctypes.pointer(kv), 1
        elif isinstance(action, np.ndarray):
            a = np.ascontiguousarray(action)
            keep.append(a)
            buf = CEnv_Value_Buffer()
            buf.b = ctypes.cast(a.ctypes.data, POINTER(c_ubyte))
            kv = CEnv_Key_Value(b"action", _TAG_OF_NUMPY[a.dtype], len(a), buf)
            c_actions, n = ctypes.pointer(kv), 1
        elif isinstance(action, dict):
            n = len(action)
            c_actions = (CEnv_Key_Value * n)()
            for i, (k, v) in enumerate(action.items()):
                a = np.ascontiguousarray(v)
                name = k.encode("ascii")
                keep += [a, name]
                c_actions[i].key = name
                c_actions[i].value_type = _TAG_OF_NUMPY[a.dtype]
                c_actions[i].value_buffer_size = len(a)
                c_actions[i].value_buffer.b = ctypes.cast(a.ctypes.data, POINTER(c_ubyte))
        else:
            raise Exception("Unrecognized action type! Supported are: int, np.array, Dict[np.array]")
        if self.lib.cenv_step(c_actions, n) != 0:
            raise Exception("Non-zero error code!")
        sd = self.c_step_data
        observation = self._collect(sd.observations_size, sd.observations)
        info = self._collect(sd.infos_size, sd.infos)
        return observation, float(sd.reward.f), bool(sd.terminated), bool(sd.truncated), info

    def reset(self, options: Optional[Dict[str, Any]] = None):
        c_options, n, _keep = _options(options)
        if self.lib.cenv_reset(c_options, n) != 0:
            raise Exception("Non-zero error code!")
        rd = self.c_reset_data
        return self._collect(rd.observations_size, rd.observations), self._collect(rd.infos_size, rd.infos)

    def render(self):
        self.lib.cenv_render()
        rd = self.c_render_data
        kv = CEnv_Key_Value(b"frame", rd.value_type, rd.value_buffer_height * rd.value_buffer_width * rd.value_buffer_channels, rd.value_buffer)
        return _copy_out(kv).reshape(rd.value_buffer_height, rd.value_buffer_width, rd.value_buffer_channels)

    def close(self):
        self.lib.cenv_close()
