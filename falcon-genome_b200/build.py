"""Build libfcs_pairhmm.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(jobs: int = 0, verbose: bool = False) -> str:
    jobs = jobs or (os.cpu_count() or 4)
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), f"-j{jobs}"]
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout[-8000:] + r.stderr[-8000:])
    if r.returncode != 0:
        raise RuntimeError("building libfcs_pairhmm.so failed")
    return os.path.join(_HERE, "libfcs_pairhmm.so")


if __name__ == "__main__":
    print(build(verbose=True))
