"""nka_b200/csrc/nka_hostcopy.h on the CPU: the thread pool that moves a pageable caller's vector into and out
of the pinned staging slots is plain C++ (no CUDA), so its splitting, hand-shake and re-use are checked here
with g++; the staged path as a whole is covered by tests/test_gpu_parity.py on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("threads", [0, 1, 3, 7])
def test_parallel_copy_is_exact_for_every_size_and_alignment(tmp_path, threads):
    exe = tmp_path / "hostcopy_test"
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(ROOT, "tests", "model", "hostcopy_test.cpp"),
                        "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), str(threads)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "hostcopy ok" in r.stdout, r.stdout + r.stderr
