"""nka_b200/csrc/nka_res_items.h on the CPU: the residual kernel's work-item numbering (only (band, strip) pairs that
hold cells, found again by the kernel's binary search) covers every grid cell exactly once -- plain C++, g++."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_item_numbering_covers_every_cell_once(tmp_path):
    exe = tmp_path / "res_items_test"
    r = subprocess.run(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "model", "res_items_test.cpp"), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "res items ok" in r.stdout, r.stdout + r.stderr
