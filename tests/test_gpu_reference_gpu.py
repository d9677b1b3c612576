"""GPU suite: the product against the REFERENCE'S OWN GPU PATH on the same B200.

oracle/_ref/libgasal_gpu.so is GASAL2 + the reference's solve_ssw_on_gpu (src/gasal2_ssw.cpp), compiled unmodified
for sm_100a from /root/reference by oracle/Makefile.  Every record field and the CIGAR text must be identical.
Inputs stay inside what the reference handles with defined behaviour: |q| <= 496 (its per-thread column array
short2 global[500] is indexed up to ceil8(|q|)-1, local_kernel_template.h:152-154) and |t| <= 2000."""
import numpy as np
import pytest

import oracle
from rabbitsalign_b200 import workload as W

pytestmark = pytest.mark.gpu

CASES = [
    ("ext150", W.extension_pairs, dict(n=3000, seed=301)),
    ("ext150_var_N", W.extension_pairs, dict(n=2000, seed=302, fixed_query_len=False, indel_rate=0.01, n_rate=0.01)),
    ("ext250_indel", W.extension_pairs, dict(n=1500, seed=303, read_len=250, indel_rate=0.05, max_indel=4, fixed_query_len=False)),
    ("adv_acgtn", W.adversarial_pairs, dict(n=4000, seed=304)),
    ("adv_iupac", W.adversarial_pairs, dict(n=2000, seed=305, alphabet=b"ACGTNacgtnRYKMSW.-")),
    ("long", W.adversarial_pairs, dict(n=300, seed=306, max_q=496, max_t=2000)),
]


@pytest.fixture(scope="module")
def ref_gpu():
    r = oracle.reference_gpu()
    if r is None:
        pytest.skip("oracle/_ref/libgasal_gpu.so not built (needs /root/reference at build time)")
    return r


@pytest.mark.parametrize("name,gen,kw", CASES, ids=[c[0] for c in CASES])
def test_product_equals_reference_gpu_path(engine, ref_gpu, name, gen, kw):
    b = gen(**kw)
    ql = np.diff(b.qoff)
    tl = np.diff(b.toff)
    keep = np.nonzero((ql <= 496) & (tl <= 2000) & (ql > 0) & (tl > 0))[0]
    b = W.from_lists([b.queries()[i] for i in keep], [b.targets()[i] for i in keep])
    out5, texts = ref_gpu.batch(b.qbuf, b.qoff, b.tbuf, b.toff, thread_id=1, cigar_stride=2048)
    res = engine.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    bad = []
    for i in range(b.n):
        got = (int(res["score"][i]), int(res["query_start"][i]), int(res["query_end"][i]), int(res["ref_start"][i]),
               int(res["ref_end"][i]), engine.cigar(res, i))
        exp = tuple(int(x) for x in out5[i]) + (texts[i],)
        if got != exp:
            bad.append(f"pair {i}: product {got} reference-gpu {exp}")
    assert not bad, f"{len(bad)} of {b.n} differ\n" + "\n".join(bad[:5])
