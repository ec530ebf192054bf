"""The C-ABI libraries load and export every symbol include/*.h declares (no compute: works
without a GPU), and the struct layouts match the reference's ctypes mirrors (cenv/cenv.py:62-111)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from procgen2_b200 import build
    build.build()
    return build


def declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"\b(%s\w+)\s*\(" % prefix, text)))


def test_engine_exports(built):
    lib = ctypes.CDLL(built.ENGINE)
    names = declared("pg2_engine.h", "pg2_")
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n


@pytest.mark.parametrize("game", ["maze", "coinrun", "bossfight", "chaser", "climber", "caveflyer", "jumper"])
def test_cenv_exports(built, game):
    lib = ctypes.CDLL(built.game_lib_path(game))
    for n in declared("cenv.h", "cenv_"):
        assert hasattr(lib, n), n
    for n in ("make_data", "reset_data", "step_data", "render_data"):
        ctypes.c_char.in_dll(lib, n)
    lib.cenv_get_env_version.restype = ctypes.c_int32
    assert lib.cenv_get_env_version() == 100      # games/coinrun/coinrun.cpp:9


def test_struct_layouts():
    from procgen2_b200 import cenv
    assert ctypes.sizeof(cenv.CEnv_Key_Value) == 24 and cenv.CEnv_Key_Value.value_buffer.offset == 16
    assert ctypes.sizeof(cenv.CEnv_Option) == 24 and cenv.CEnv_Option.value.offset == 16
    assert ctypes.sizeof(cenv.CEnv_Step_Data) == 40
    assert cenv.CEnv_Step_Data.terminated.offset == 24 and cenv.CEnv_Step_Data.truncated.offset == 25
    assert cenv.CEnv_Step_Data.infos.offset == 32


def test_create_fails_loudly_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from procgen2_b200.engine import BatchedEnv
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        BatchedEnv("maze", 4)


def test_vector_env_distribution_mode_names():
    """The tensor front-end takes the reference's mode names; a name that is no mode fails before anything touches the GPU."""
    from procgen2_b200.vector_env import ProcgenVectorEnv as V
    assert (V.DISTRIBUTION_MODES[None], V.DISTRIBUTION_MODES["easy"], V.DISTRIBUTION_MODES["hard"]) == (-1, 0, 1)
    assert V.DISTRIBUTION_MODES["memory"] == V.DISTRIBUTION_MODES["extreme"] == 2      # games/<g>/tilemap.h Distribution_Mode
    with pytest.raises(ValueError, match="distribution_mode"):
        V("maze", 4, distribution_mode="nightmare")
