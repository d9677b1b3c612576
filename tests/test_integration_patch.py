"""CPU suite: the build-time caller-loop substitution (integration/patch_caller.py) finds exactly its eight sites in the
reference's src/pc.cpp and changes nothing else.  Skipped where /root/reference is absent (GPU box)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PC = os.path.join(os.environ.get("REF_ROOT", "/root/reference"), "src", "pc.cpp")


@pytest.mark.skipif(not os.path.exists(REF_PC), reason="reference sources not present")
def test_patch_caller_substitutes_exactly_eight_sites(tmp_path):
    out = tmp_path / "pc.cpp"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "integration", "patch_caller.py"), REF_PC, str(out)])
    a = open(REF_PC).read().splitlines()
    b = out.read_text().splitlines()
    assert len(a) == len(b)
    changed = [(x, y) for x, y in zip(a, b) if x != y]
    assert len(changed) == 8
    assert sum("rsa_ext_gasal_fail(" in y for _, y in changed) == 4
    assert sum("rsa_ext_align_gpu(aligner, " in y for _, y in changed) == 4
    for x, y in changed:
        assert ("gasal_fail(" in x) or ("aligner.align_gpu(" in x)
