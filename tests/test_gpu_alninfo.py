"""GPU suite, SURVEY 8(f) "next" rows 1+3: the device-side gasal_fail + Aligner::align_gpu post-processing
(rsa_ext_request_alninfo) against the REFERENCE's own host code: gasal_fail restated in oracle/sw_oracle.c and
Aligner::align_gpu compiled from /root/reference into oracle/_ref/libssw_ref_*.so."""
import numpy as np
import pytest

import oracle
from rabbitsalign_b200 import ExtensionEngine, workload as W
from rabbitsalign_b200.ext import ALNINFO_DTYPE, alninfo_cigar_string

pytestmark = pytest.mark.gpu

CASES = [
    ("ext150", W.extension_pairs, dict(n=3000, seed=301)),
    ("ext150_var", W.extension_pairs, dict(n=2000, seed=302, fixed_query_len=False, indel_rate=0.01)),
    ("ext250_indel", W.extension_pairs, dict(n=1200, seed=303, read_len=250, indel_rate=0.03, max_indel=4, fixed_query_len=False)),
    ("ext150_N", W.extension_pairs, dict(n=2000, seed=304, n_rate=0.01)),
    ("adv_acgtn", W.adversarial_pairs, dict(n=4000, seed=305)),
    ("adv_mid", W.adversarial_pairs, dict(n=2000, seed=306, max_q=200, max_t=400)),
]


@pytest.mark.parametrize("name,gen,kw", CASES, ids=[c[0] for c in CASES])
def test_alninfo_matches_reference_host_code(name, gen, kw):
    ssw = oracle.ssw_reference()
    if ssw is None:
        pytest.skip("oracle/_ref/libssw_ref not built")
    b = gen(**kw)
    e = ExtensionEngine()
    aln = np.zeros(b.n, dtype=ALNINFO_DTYPE)
    e.request_alninfo(aln, end_bonus=10)
    res = e.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    e.request_alninfo(None)
    # reference side: oracle record -> gasal_fail -> Aligner::align_gpu
    qs, ts = b.queries(), b.targets()
    recs = oracle.restatement().align(qs, ts)
    fails = np.array([oracle.gasal_fail(q, t, r) for q, t, r in zip(qs, ts, recs)])
    # (records the gate rejects never reach align_gpu in the reference: hand it a harmless stand-in for those)
    g = dict(score=[r.score if not f else 2 for r, f in zip(recs, fails)],
             qs=[r.query_start if not f else 0 for r, f in zip(recs, fails)],
             qe=[r.query_end if not f else 0 for r, f in zip(recs, fails)],
             rs=[r.ref_start if not f else 0 for r, f in zip(recs, fails)],
             re=[r.ref_end if not f else 0 for r, f in zip(recs, fails)],
             cigar=[r.cigar_str if not f else "1M" for r, f in zip(recs, fails)])
    exp = ssw.align_gpu_packed(b.qbuf, b.qoff, b.tbuf, b.toff, g)
    bad = []
    n_ok = 0
    for i in range(b.n):
        st = int(aln["status"][i])
        if fails[i]:
            if st != 1:
                bad.append((i, "expected gasal_fail", st))
            continue
        if st == 3:
            assert int(res["n_ops"][i]) > 0  # long CIGAR: record path, nothing to compare here
            continue
        got = (int(aln["sw_score"][i]), int(aln["query_start"][i]), int(aln["query_end"][i]), int(aln["ref_start"][i]),
               int(aln["ref_end"][i]), int(aln["edit_distance"][i]), alninfo_cigar_string(aln[i]))
        want = (int(exp["score"][i]), int(exp["qs"][i]), int(exp["qe"][i]), int(exp["rs"][i]), int(exp["re"][i]),
                int(exp["ed"][i]), exp["cigar"][i])
        if st != 0 or got != want:
            bad.append((i, st, got, want, qs[i], ts[i]))
        n_ok += 1
    e.close()
    assert not bad, bad[:3]
    assert n_ok > 0.5 * b.n or name.startswith("adv")
