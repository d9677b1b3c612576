#!/usr/bin/env python3
"""integration/patch_sam.py <reference src/sam.cpp> <out sam.cpp> <pc.cpp in> <pc.cpp out>

The edit a maintainer makes to take the SAM text from the device (the caller half),
applied at BUILD time to copies under integration/_build/ (git-ignored; no reference source enters this repo).

  src/sam.cpp  the three member functions that append a record's text -- Sam::add_record (:141-206), Sam::add_unmapped
               (:73-86), Sam::add_unmapped_mate (:88-110) -- first offer their arguments to the chunk collector
               (rsa_glue::sam_collect_*, integration/sam_glue.hpp) and return when it took them;
  src/pc.cpp   in the four worker loops (perform_task_async_{se,pe}{,_fx}): rsa_glue::sam_begin right after the chunk's
               `Sam sam{...}` is constructed (:779, :1064, :1477, :1850), rsa_glue::sam_flush right before the chunk's string
               goes to the output buffer (:791, :1076, :1495, :1868).
"""
import re
import sys

sam_in, sam_out, pc_in, pc_out = sys.argv[1:5]


def sub(text, pattern, repl, expect, what, flags=0):
    text, n = re.subn(pattern, repl, text, flags=flags)
    if n != expect:
        sys.exit(f"patch_sam.py: expected {expect} sites for {what}, found {n}")
    return text


WRITER = "rsa_glue::SamWriter{sam_string, references, cigar_ops, tail, output_unmapped, show_details}"

text = open(sam_in).read()
text = sub(text, r"(void Sam::add_unmapped\(const KSeq& record, uint16_t flags\) \{\n)",
           r"\1    if (rsa_glue::sam_collect_unmapped(" + WRITER + r", record, flags)) return;\n", 1, "Sam::add_unmapped")
text = sub(text, r"(void Sam::add_unmapped_mate\([^)]*\) \{\n)",
           r"\1    if (rsa_glue::sam_collect_unmapped_mate(" + WRITER + r", record, flags, mate_reference_name, mate_pos)) return;\n",
           1, "Sam::add_unmapped_mate")
text = sub(text, r"(void Sam::add_record\([^)]*\) \{\n)",
           r"\1    if (rsa_glue::sam_collect_record(" + WRITER + ", query_name, flags, reference_name, pos, mapq, cigar,\n"
           "                                     mate_reference_name, mate_pos, template_len, query_sequence, qual, ed, aln_score, details)) return;\n",
           1, "Sam::add_record")
open(sam_out, "w").write('#include "sam_glue.hpp"\n' + text)

text = open(pc_in).read()
out, pos, n = [], 0, 0
for m in re.finditer(r"([ \t]*)output_buffer\.output_records\(std::move\(sam_out\), pre_chunk_index\);", text):
    ctor = text.rfind("Sam sam{", pos, m.start())
    if ctor < 0:
        sys.exit("patch_sam.py: no Sam constructor before an output_records call")
    ctor_end = text.index("};", ctor) + 2
    indent = m.group(1)
    out.append(text[pos:ctor_end] + f"\n{indent}rsa_glue::sam_begin(thread_id, sam_out);" + text[ctor_end:m.start()] +
               f"{indent}rsa_glue::sam_flush(thread_id, sam_out);\n")
    pos = m.start()
    n += 1
out.append(text[pos:])
if n != 4:
    sys.exit(f"patch_sam.py: expected 4 worker loops in pc.cpp, found {n}")
open(pc_out, "w").write('#include "sam_glue.hpp"\n' + "".join(out))
