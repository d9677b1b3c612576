#!/usr/bin/env python3
"""integration/patch_caller.py <reference src/pc.cpp> <out.cpp>

The edit a maintainer makes to adopt the device-side `AlignmentInfo` (optional section): at the four
caller loops of src/pc.cpp (:735-744 and its three siblings) the two calls on a GPU record,

    gasal_fail(todo_querys[i], todo_refs[i], gasal_results[i])      -> rsa_ext_gasal_fail(...)
    aligner.align_gpu(todo_querys[i], todo_refs[i], gasal_results[i]) -> rsa_ext_align_gpu(aligner, ...)

become the helpers of integration/gasal2_ssw.h (RSA_EXT_ALNINFO).  Applied at BUILD time to a copy under
integration/_build/ (git-ignored): no reference source enters this repo.  Fails if the number of sites is not 4.
"""
import re
import sys

src, out = sys.argv[1], sys.argv[2]
text = open(src).read()
args = r"\(\s*todo_querys\[i\]\s*,\s*todo_refs\[i\]\s*,\s*gasal_results\[i\]\s*\)"
text, n_fail = re.subn(r"\bif\s*\(\s*gasal_fail" + args, "if (rsa_ext_gasal_fail(todo_querys[i], todo_refs[i], gasal_results[i])", text)
text, n_aln = re.subn(r"\baligner\.align_gpu" + args, "rsa_ext_align_gpu(aligner, todo_querys[i], todo_refs[i], gasal_results[i])", text)
if n_fail != 4 or n_aln != 4:
    sys.exit(f"patch_caller.py: expected 4 + 4 call sites, found {n_fail} + {n_aln}")
open(out, "w").write(text)
