"""TEST INFRASTRUCTURE. Builds tests/hostsim/_build/libpg2_hostsim.so (g++, no CUDA)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build", "libpg2_hostsim.so")


def build(force=False):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    csrc = os.path.join(ROOT, "procgen2_b200", "csrc")
    deps = [os.path.join(HERE, "hostsim.cpp")]
    for d, _, files in os.walk(csrc):
        deps += [os.path.join(d, f) for f in files]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    cmd = ["g++", "-std=c++17", "-O2", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++",
           os.path.join(HERE, "hostsim.cpp"), os.path.join(csrc, "assets.cpp"), "-lz", "-ldl", "-o", OUT]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
