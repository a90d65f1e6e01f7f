"""The C++ host mirror (gridsplines_b200/host/gridsplines.hpp) over the C ABI: compiles on the CPU box; on the GPU
box it runs config C1, two README known answers and the GridTest scenario and must print the oracle's bits."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")


def _compile(out):
    lib_dir = os.path.join(ROOT, "gridsplines_b200")
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", SRC, "-o", out, "-L" + lib_dir, "-lgs_b200",
           "-Wl,-rpath," + lib_dir]
    subprocess.check_call(cmd)


def test_cpp_mirror_compiles_and_links(tmp_path):
    _compile(str(tmp_path / "host_mirror_test"))


@pytest.mark.gpu
def test_cpp_mirror_reproduces_kats(tmp_path):
    exe = str(tmp_path / "host_mirror_test")
    _compile(exe)
    out = subprocess.check_output([exe], text=True)
    vals = {l.split()[0]: l.split()[1:] for l in out.splitlines()}
    assert float(vals["c1_grid0"][0]) == 19.163547799904077 and vals["c1_grid0"][1] == "0x403329de44c3ebb7"  # KAT R2
    assert float(vals["c1_grid99"][0]) == 6.303223056805375 and float(vals["c1_grid199"][0]) == 4.014707461585766
    assert float(vals["readme_135"][0]) == 1.718125                       # README.md:85-86
    assert abs(float(vals["readme_area_150_185"][0]) - 139.148203125) < 1e-12 * 139.148203125  # README.md:99-100
    assert vals["gridtest_dims"] == ["137", "45"]
    assert float(vals["gridtest_grid0"][0]) == 54.44328715390563          # KAT R1
    assert float(vals["gridtest_grid1"][0]) == 53.94901954479779
    assert vals["gridtest_slices_ok"] == ["1"]
