"""Small all-paths check (a few hundred pairs, seconds): packed 8/16-lane kernels, N variant, exact kernel, redo
pass, long-CIGAR arena, device-side align_gpu, each compared with the oracle.  python tools/sanitize_check.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from parity_util import compare, oracle_arrays  # noqa: E402
from rabbitsalign_b200 import ExtensionEngine, workload as W  # noqa: E402
from rabbitsalign_b200.ext import ALNINFO_DTYPE  # noqa: E402

olib = oracle.restatement()
eng = ExtensionEngine()
ok = True
cases = [W.extension_pairs(120, seed=1), W.extension_pairs(60, seed=2, n_rate=0.02),
         W.extension_pairs(40, seed=3, read_len=300, fixed_query_len=False, indel_rate=0.02),
         W.adversarial_pairs(150, seed=4, alphabet=b"ACGTNRY"), W.adversarial_pairs(20, seed=5, max_q=500, max_t=2000),
         W.extension_pairs(30, seed=6, read_len=250, indel_rate=0.12, max_indel=2, sub_rate=0.05, fixed_query_len=False)]
for b in cases:
    aln = np.zeros(b.n, dtype=ALNINFO_DTYPE)
    eng.request_alninfo(aln)
    res = eng.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    bad = compare(eng, res, oracle_arrays(olib, b), b)
    print(b.n, "pairs", "OK" if not bad else bad[:2], eng.stats()["pairs_fast"], eng.stats()["pairs_exact"], int((res["n_ops"] > 40).sum()))
    ok &= not bad
eng.close()
sys.exit(0 if ok else 1)
