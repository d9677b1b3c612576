#!/usr/bin/env python3
"""integration/patch_caller.py [--windows] <reference src/pc.cpp> <out.cpp>

The edits a maintainer makes to adopt the optional parts of the boundary, applied at BUILD time to a
copy under integration/_build/ (git-ignored): no reference source enters this repo.  Every substitution checks the
number of sites it expects and fails otherwise.

default (device-side AlignmentInfo, SURVEY 8f ranks 1+3): at the four caller loops of src/pc.cpp (:735-744 and its
three siblings) the two calls on a GPU record,

    gasal_fail(todo_querys[i], todo_refs[i], gasal_results[i])        -> rsa_ext_gasal_fail(...)
    aligner.align_gpu(todo_querys[i], todo_refs[i], gasal_results[i]) -> rsa_ext_align_gpu(aligner, ...)

become the helpers of integration/gasal2_ssw.h (RSA_EXT_ALNINFO).

--windows (the caller half; compile with -DRSA_EXT_WINDOWS), additionally:

  * part2_extend_seed_get_str / part2_rescue_mate_get_str (src/pc.cpp:214-242, 333-368) no longer build the window
    with substr: they record (contig, start, length) as an RsaWindow (same clipping as substr);
  * the chunk's todo list is std::vector<RsaWindow> (4 declarations, 2 parameters);
  * the helper thread's slice loop (src/pc.cpp:643-673 and its three siblings: copies of 512 strings per call)
    becomes ONE solve_ssw_on_gpu_windows call per chunk; the engine reads the windows from the genome resident in HBM.
"""
import re
import sys

argv = sys.argv[1:]
windows = "--windows" in argv
argv = [a for a in argv if a != "--windows"]
src, out = argv
text = open(src).read()


def sub(pattern, repl, expect, what, flags=0):
    global text
    text, n = re.subn(pattern, repl, text, flags=flags)
    if n != expect:
        sys.exit(f"patch_caller.py: expected {expect} sites for {what}, found {n}")


args = r"\(\s*todo_querys\[i\]\s*,\s*todo_refs\[i\]\s*,\s*gasal_results\[i\]\s*\)"
sub(r"\bif\s*\(\s*gasal_fail" + args, "if (rsa_ext_gasal_fail(todo_querys[i], todo_refs[i], gasal_results[i])", 4, "gasal_fail")
sub(r"\baligner\.align_gpu" + args, "rsa_ext_align_gpu(aligner, todo_querys[i], todo_refs[i], gasal_results[i])", 4, "align_gpu")

if windows:
    sub(r"std::vector<std::string>&\s*todo_refs", "std::vector<RsaWindow>& todo_refs", 2, "get_str parameters")
    sub(r"std::vector<std::string>\s+todo_refs;", "std::vector<RsaWindow> todo_refs;", 4, "todo_refs declarations")
    sub(r"const auto ref_segm = ref\.substr\(ref_start, ref_segm_size\);\s*todo_querys\.push_back\(query\);\s*"
        r"todo_refs\.push_back\(ref_segm\);",
        "todo_querys.push_back(query);\n    todo_refs.emplace_back(references.sequences, nam.ref_id, ref_start, ref_segm_size);",
        1, "extend-seed window")
    sub(r"std::string ref_segm = references\.sequences\[nam\.ref_id\]\.substr\(ref_start, ref_end - ref_start\);\s*"
        r"todo_querys\.push_back\(r_tmp\);\s*todo_refs\.push_back\(ref_segm\);",
        "todo_querys.push_back(r_tmp);\n    todo_refs.emplace_back(references.sequences, nam.ref_id, ref_start, ref_end - ref_start);",
        1, "mate-rescue window")
    sub(r"gpu_ssw_async = std::thread\(\[&\] \(\)\{.*?\}\);",
        "gpu_ssw_async = std::thread([&] (){\n"
        "                solve_ssw_on_gpu_windows(thread_id, gasal_results, todo_querys, todo_refs, references.sequences,\n"
        "                                         aln_params.match, aln_params.mismatch, aln_params.gap_open, aln_params.gap_extend);\n"
        "            });",
        4, "slice loops", flags=re.S)
open(out, "w").write(text)
