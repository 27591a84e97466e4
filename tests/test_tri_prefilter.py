"""tri_test (division-free prefilter, csrc/exact.cuh) == the plain statement of Triangle::intersect
(reference include/triangle.hpp:23-58), decision and bits, on random and boundary-aimed inputs.  Host build of
the very header the kernels include (its arithmetic macros fall back to plain fp32 ops on the host; compiled
with -ffp-contract=off and no -march, like the oracle)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


@pytest.mark.skipif(not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")) or shutil.which("g++") is None,
                    reason="needs g++ and the CUDA headers")
def test_prefilter_matches_plain(tmp_path):
    exe = str(tmp_path / "tri_prefilter_check")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-I", CUDA_INC, "-o", exe,
                    os.path.join(ROOT, "tests", "tri_prefilter_check.cpp")], check=True)
    total = acc = 0
    for seed in (1, 2, 3):
        res = subprocess.run([exe, "2500000", str(seed)], capture_output=True, text=True)
        assert res.returncode == 0, res.stderr + res.stdout
        n, a, bad = map(int, res.stdout.split())
        assert bad == 0
        total += n; acc += a
    assert total > 25_000_000 and acc > 3_000_000   # the generator really produces hits and near-misses
