#!/usr/bin/env python3
"""integration/patch_seed.py <reference src/aln.cpp> <out aln.cpp> <reference-or-patched pc.cpp> <out pc.cpp>

The edit a maintainer makes to adopt GPU seeding, applied at BUILD time to copies under
integration/_build/ (git-ignored; no reference source enters this repo):

  src/aln.cpp   inside align_SE_read_part and align_PE_read_part only (the other callers of these functions are not on
                the pipeline's path): randstrobes_query / find_nams / find_nams_rescue -> rsa_glue::... (seed_glue.hpp),
                which hand out the results the GPU computed for the whole chunk;
  src/pc.cpp    before each of the eight per-read loops that call align_SE_read_part / align_PE_read_part
                (src/pc.cpp:584,710,879,995,1190,1378,1584,1752): one rsa_glue::seed_chunk_se / _pe call for the chunk.
"""
import re
import sys

aln_in, aln_out, pc_in, pc_out = sys.argv[1:5]

# ---- aln.cpp
text = open(aln_in).read()
total = {"randstrobes_query": 0, "find_nams": 0, "find_nams_rescue": 0}
for fn in ("align_SE_read_part", "align_PE_read_part"):
    m = re.search(r"\nvoid " + fn + r"\(", text)
    if not m:
        sys.exit(f"patch_seed.py: {fn} not found")
    end = text.index("\n}\n", m.start()) + 3
    body = text[m.start():end]
    body, a = re.subn(r"(?<![:\w])randstrobes_query\(", "rsa_glue::randstrobes_query(", body)
    body, c = re.subn(r"(?<![:\w])find_nams_rescue\(", "rsa_glue::find_nams_rescue(", body)
    body, b = re.subn(r"(?<![:\w])find_nams\(", "rsa_glue::find_nams(", body)
    total["randstrobes_query"] += a; total["find_nams"] += b; total["find_nams_rescue"] += c
    text = text[:m.start()] + body + text[end:]
if total != {"randstrobes_query": 2, "find_nams": 2, "find_nams_rescue": 2}:
    sys.exit(f"patch_seed.py: unexpected call sites in aln.cpp: {total}")
text = '#include "seed_glue.hpp"\n' + text
open(aln_out, "w").write(text)

# ---- pc.cpp
text = open(pc_in).read()
out, pos, n_se, n_pe = [], 0, 0, 0
for m in re.finditer(r"align_(SE|PE)_read_part\(", text):
    loop = text.rfind("for (size_t i = 0;", pos, m.start())
    if loop < 0:
        sys.exit("patch_seed.py: no loop before an align_*_read_part call")
    head = text[loop:text.index("\n", loop)]
    vec = re.search(r"i < (\w+)\.size\(\)", head).group(1)
    indent = text[text.rfind("\n", 0, loop) + 1:loop]
    if m.group(1) == "SE":
        call = f"rsa_glue::seed_chunk_se(thread_id, {vec}, index_parameters, index, map_param);\n{indent}"
        n_se += 1
    else:
        vec2 = vec.replace("records1", "records2")
        call = f"rsa_glue::seed_chunk_pe(thread_id, {vec}, {vec2}, index_parameters, index, map_param);\n{indent}"
        n_pe += 1
    out.append(text[pos:loop] + call)
    pos = loop
out.append(text[pos:])
if (n_se, n_pe) != (4, 4):
    sys.exit(f"patch_seed.py: expected 4 + 4 loops in pc.cpp, found {n_se} + {n_pe}")
open(pc_out, "w").write('#include "seed_glue.hpp"\n' + "".join(out))
