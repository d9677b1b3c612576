"""CPU suite, SURVEY 8f rank 4: the plain restatement of the reference's SAM writer (tests/sam_format_util.py: restate)
against the reference's own `class Sam` compiled from /root/reference (oracle/_ref/libsam_ref.so), byte for byte; and
the product library exports the formatter's ABI (include/rsa_sam.h)."""
import ctypes
import os
import re

import pytest

import oracle
from rabbitsalign_b200 import sam as S
from rabbitsalign_b200.ext import LIB_PATH
from sam_format_util import VARIANTS, make_calls, restate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("variant", range(len(VARIANTS)))
def test_restatement_equals_reference_sam_writer(variant):
    kw = VARIANTS[variant]
    ref_names, calls, text, cig = make_calls(3000, seed=100 + variant)
    want = oracle.sam_reference_replay(ref_names, calls, text, cig, **kw)
    if want is None:
        pytest.skip("oracle/_ref/libsam_ref.so not built")
    got = restate(ref_names, calls, text, cig, **kw)
    assert len(want) > 100_000
    if got != want:
        gl, wl = got.split(b"\n"), want.split(b"\n")
        bad = [(i, a, b) for i, (a, b) in enumerate(zip(gl, wl)) if a != b][:3]
        raise AssertionError(f"{len(gl)} vs {len(wl)} lines; first differences: {bad}")


def test_library_exports_the_sam_abi():
    if not os.path.exists(LIB_PATH):
        pytest.skip("librsa_ext.so not built")
    lib = ctypes.CDLL(LIB_PATH)
    declared = set(re.findall(r"\b(rsa_sam_[a-z_]+)\s*\(", open(os.path.join(ROOT, "include", "rsa_sam.h")).read()))
    assert declared == set(S.SAM_ABI_SYMBOLS)
    for sym in declared:
        assert hasattr(lib, sym), sym
