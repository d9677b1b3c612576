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


def _added_lines(before_path, after_path):
    """The patchers below only INSERT lines: return them and check that nothing else moved."""
    import difflib
    a = open(before_path).read().splitlines()
    b = open(after_path).read().splitlines()
    added = []
    for tag, i1, i2, j1, j2 in difflib.SequenceMatcher(None, a, b, autojunk=False).get_opcodes():
        if tag == "equal":
            continue
        assert tag == "insert", (tag, a[i1:i2], b[j1:j2])
        added += b[j1:j2]
    return added


REF_SRC = os.path.dirname(REF_PC)


@pytest.mark.skipif(not os.path.exists(REF_PC), reason="reference sources not present")
def test_patch_hamming_only_inserts_its_call_sites(tmp_path):
    """integration/patch_hamming.py (caller half): the deferral test in extend_seed_part, the two scope
    guards, and one hamming_pass call before each of the four step1 loops -- nothing of the reference is rewritten."""
    aln, pc = tmp_path / "aln.cpp", tmp_path / "pc.cpp"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "integration", "patch_hamming.py"),
                           os.path.join(REF_SRC, "aln.cpp"), str(aln), REF_PC, str(pc)])
    added = _added_lines(os.path.join(REF_SRC, "aln.cpp"), str(aln))
    assert added[0] == '#include "hamming_glue.hpp"' and len(added) == 8
    assert sum("rsa_glue::HammingDefer hamming_scope(" in ln for ln in added) == 2
    assert sum("rsa_glue::push_pending(align_tmp_res, nam, projected_ref_start);" in ln for ln in added) == 1
    added = _added_lines(REF_PC, str(pc))
    assert added[0] == '#include "hamming_glue.hpp"' and len(added) == 5
    assert sum("rsa_glue::hamming_pass_se(thread_id, pre_records3, pre_align_tmp_results, references, aligner);" in ln for ln in added) == 2
    assert sum("rsa_glue::hamming_pass_pe(thread_id, pre_records1, pre_records2, pre_align_tmp_results, references, aligner);" in ln
               for ln in added) == 2


@pytest.mark.skipif(not os.path.exists(REF_PC), reason="reference sources not present")
def test_patch_sam_only_inserts_its_call_sites(tmp_path):
    """integration/patch_sam.py (caller half): one collector call at the top of the three text-appending
    members of class Sam, sam_begin / sam_flush around the four worker loops."""
    sam, pc = tmp_path / "sam.cpp", tmp_path / "pc.cpp"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "integration", "patch_sam.py"),
                           os.path.join(REF_SRC, "sam.cpp"), str(sam), REF_PC, str(pc)])
    added = _added_lines(os.path.join(REF_SRC, "sam.cpp"), str(sam))
    assert added[0] == '#include "sam_glue.hpp"' and len(added) == 5
    for fn in ("sam_collect_unmapped(", "sam_collect_unmapped_mate(", "sam_collect_record("):
        assert sum(("rsa_glue::" + fn) in ln for ln in added) == 1
    added = _added_lines(REF_PC, str(pc))
    assert added[0] == '#include "sam_glue.hpp"' and len(added) == 9
    assert sum("rsa_glue::sam_begin(thread_id, sam_out);" in ln for ln in added) == 4
    assert sum("rsa_glue::sam_flush(thread_id, sam_out);" in ln for ln in added) == 4
