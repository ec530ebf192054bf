"""libstdc++ emulation pins: pg2::USet iteration order == std::unordered_set<int> (real one),
checked by a small C++ program compiled on the fly."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_uset_matches_libstdcxx(tmp_path):
    exe = str(tmp_path / "test_uset")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_uset.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("OK")


def test_uset_order_regrouping_matches_libstdcxx(tmp_path):
    """pg2::USetOrder (pg2_roomgen.cuh): the warp-parallel computation of a fresh unordered_set's iteration order."""
    exe = str(tmp_path / "test_uset_order")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_uset_order.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("OK")


def test_roomgen_bit_rows_match_per_cell_restatement(tmp_path):
    """Cellular automaton and path dilation on 64-bit bit rows (pg2_roomgen.cuh) == per-cell restatements of
    room_generator.cpp on random grids."""
    exe = str(tmp_path / "test_roomgen_bits")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_roomgen_bits.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("OK")
