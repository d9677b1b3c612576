"""Small run of every kernel family (extension: 4/8/16-lane and flex classes, exact redo, window form, several chunks in
flight; Hamming shortcut; SAM formatter; seeding in both tiers with rescue), each checked against the oracle / the reference's
own code.  Sizes are tiny on purpose: a quick whole-product check on a GPU box (compute-sanitizer is closed on this pool, so
bounds are guarded by the parity tests and the engine's own checks instead)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import __graft_entry__ as G  # noqa: E402

G.smoke()   # packed DP (4 lanes) + traceback + exact redo, window form from the packed planes, Hamming, SAM formatter

import oracle  # noqa: E402
from rabbitsalign_b200 import ExtensionEngine, workload as W  # noqa: E402
from rabbitsalign_b200.ext import RESULT_DTYPE  # noqa: E402


def check(eng, b, tag):
    res = eng.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    score, qs, qe, rs, re, nops, pool, coff = oracle.restatement().align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    for f, exp in (("score", score), ("query_start", qs), ("query_end", qe), ("ref_start", rs), ("ref_end", re)):
        assert np.array_equal(res[f], exp), (tag, f)
    assert np.array_equal(res["n_ops"].astype(np.int32), nops), tag
    raw = pool.tobytes()
    for i in range(0, b.n, 7):
        assert eng.cigar(res, i) == raw[coff[i]:coff[i + 1]].decode(), (tag, i)


eng = ExtensionEngine(device=0)
# 8- and 16-lane groups, flex classes (variable query lengths), long CIGARs (arena), device planner with alninfo
check(eng, W.extension_pairs(600, seed=5, read_len=250, indel_rate=0.05, max_indel=4, fixed_query_len=False), "250bp indel")
check(eng, W.extension_pairs(300, seed=6, read_len=300), "300bp")
check(eng, W.extension_pairs(300, seed=8, read_len=100), "100bp")
eng.close()
# several chunks in flight (small scratch budget): slots, streams, plan-ahead
eng = ExtensionEngine(device=0, scratch_bytes=64 << 20)
check(eng, W.extension_pairs(6000, seed=9), "multi-chunk")
eng.close()

# seeding: both tiers + rescue, a few hundred reads per case
import seed_util as SU  # noqa: E402
from rabbitsalign_b200 import seed as S  # noqa: E402

if oracle.seed_reference_lib() is not None:
    for name in ("r150_repeats", "rescue_heavy", "N_rich"):
        g, rl, r = SU.CASES[name]
        contigs = W.seeding_genome(**g)
        idx = oracle.build_seed_index(contigs, rl, 4)
        buf, off = W.seeding_reads(contigs, **dict(r, n=400))
        gi = S.SeedIndexGpu(S.make_config(idx.params()), idx.randstrobes, idx.starts)
        sd = S.Seeder(gi)
        per, nams = sd.find_nams(buf, off)
        st = sd.stats()
        sd.close(); gi.close()
        SU.assert_equals_reference(idx, buf, off, per, nams)
        print("seeding", name, "ok:", st["reads"], "reads,", st["nams"], "NAMs,", st["reads_retried"], "in the warp tier")
        idx.close()
else:
    print("seeding probe skipped (no reference seeding library)")
print("probe ok")
