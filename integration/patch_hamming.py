#!/usr/bin/env python3
"""integration/patch_hamming.py <aln.cpp in> <aln.cpp out> <pc.cpp in> <pc.cpp out>

The edit a maintainer makes to take the Hamming shortcut of the seed extension from the device (the caller
half), applied at BUILD time to copies under integration/_build/ (git-ignored; no reference source enters
this repo).  Inputs are the copies patch_seed.py / patch_caller.py --windows produced (the edits stack).

  src/aln.cpp  extend_seed_part (:374-431): a candidate whose projected window has the read's length is left PENDING
               (rsa_glue::push_pending) instead of being decided on the spot, when deferral is on;
               align_SE_part (:95) / align_PE_part (:1372): a scope guard switches deferral on -- always for single-end
               reads, for pairs once the insert-size estimator has its 400 samples (it is fed by shortcut results until
               then, :1450-1466, so until then the host code runs unchanged);
  src/pc.cpp   before each of the four "step1" loops that turn a chunk's todo entries into (query, window) lists (:610,
               :905, :1219, :1613): one rsa_glue::hamming_pass_se / _pe call settles the chunk's pending candidates.
"""
import re
import sys

aln_in, aln_out, pc_in, pc_out = sys.argv[1:5]


def sub(text, pattern, repl, expect, what, flags=0):
    text, n = re.subn(pattern, repl, text, flags=flags)
    if n != expect:
        sys.exit(f"patch_hamming.py: expected {expect} sites for {what}, found {n}")
    return text


# ---- aln.cpp
text = open(aln_in).read()
# the DEFINITION of extend_seed_part (the forward declaration at :36 has no body)
text = sub(text,
           r"(static inline bool extend_seed_part\([^)]*\)\s*\{.*?bool gapped = true;\n)",
           r"\1"
           "    if (projected_ref_end - projected_ref_start == query.size() && consistent_nam && rsa_glue::hamming_deferred() &&\n"
           "        query.size() <= rsa_glue::kHammingMaxQuery) {\n"
           "        rsa_glue::push_pending(align_tmp_res, nam, projected_ref_start);\n"
           "        return true;\n"
           "    }\n",
           1, "extend_seed_part", flags=re.S)
text = sub(text, r"(static inline void align_SE_part\([^)]*\)\s*\{\n)",
           r"\1    rsa_glue::HammingDefer hamming_scope(true);\n", 1, "align_SE_part")
text = sub(text, r"(\ninline void align_PE_part\([^)]*\)\s*\{\n)",
           r"\1    rsa_glue::HammingDefer hamming_scope(isize_est.sample_size >= 400);\n", 1, "align_PE_part")
open(aln_out, "w").write('#include "hamming_glue.hpp"\n' + text)

# ---- pc.cpp
text = open(pc_in).read()
out, pos, n_se, n_pe = [], 0, 0, 0
for m in re.finditer(r"[ \t]*// step1 : filter nams and get todo_strings\n", text):
    head = text[m.end():text.index("{", text.index("for (", m.end()))]
    vec = re.search(r"i < (\w+)\.size\(\)", head).group(1)
    indent = re.match(r"[ \t]*", m.group(0)).group(0)
    if vec.endswith("records3"):
        call = f"{indent}rsa_glue::hamming_pass_se(thread_id, {vec}, pre_align_tmp_results, references, aligner);\n"
        n_se += 1
    else:
        vec2 = vec.replace("records1", "records2")
        call = f"{indent}rsa_glue::hamming_pass_pe(thread_id, {vec}, {vec2}, pre_align_tmp_results, references, aligner);\n"
        n_pe += 1
    out.append(text[pos:m.start()] + call)
    pos = m.start()
out.append(text[pos:])
if (n_se, n_pe) != (2, 2):
    sys.exit(f"patch_hamming.py: expected 2 + 2 step1 loops in pc.cpp, found {n_se} + {n_pe}")
open(pc_out, "w").write('#include "hamming_glue.hpp"\n' + "".join(out))
