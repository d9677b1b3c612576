"""CPU suite: the packed Smith-Waterman cell recipe the CUDA kernel runs (rabbitsalign_b200/csrc/fast_cell.cuh: fast_cell,
profile_word, dir_pair, dir_word), compiled for the HOST by tests/cell_host_check.cu and compared with a plain integer
restatement of one cell of the reference (GASAL2/src/kernels/local_kernel_template.h:45-60) on 4 M random and boundary
cells over five scorings, plus every combination of four direction nibbles through the gather."""
import json
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc (host compilation of the recipe header)")
def test_cell_recipe_on_the_host_equals_plain_integer_cell():
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "cell_host_check")
    src = os.path.join(ROOT, "tests", "cell_host_check.cu")
    hdr = os.path.join(ROOT, "rabbitsalign_b200", "csrc", "fast_cell.cuh")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["nvcc", "-O1", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe, src])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["bad"] == 0 and res["checked"] > 4_000_000
