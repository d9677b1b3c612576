"""Generate tests/golden/pairs_*.json from the REFERENCE's own kernels compiled for the host
(oracle/_ref/libgasal_ref*.so, built from /root/reference by oracle/Makefile).  Run in the dev container:

    python oracle/make_golden.py

Each file: {"scoring": {...}, "source": "...", "pairs": [{"q": ..., "t": ..., "res": [score, query_start,
query_end, ref_start, ref_end, cigar]}]}.  tests/test_oracle.py replays them through the C restatement;
the GPU suite replays them through the CUDA path.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from rabbitsalign_b200 import workload as W  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def dump(name, batch, ref, scoring=None, engine=None):
    """engine: ExtensionEngine limits the GPU replay needs (max_query_len / max_target_len beyond the defaults)."""
    scoring = scoring or {}
    qs, ts = batch.queries(), batch.targets()
    res = ref.align(qs, ts, **scoring)
    pairs = [{"q": q.decode("latin1"), "t": t.decode("latin1"), "res": list(r.astuple())} for q, t, r in zip(qs, ts, res)]
    doc = {"scoring": scoring, "source": os.path.basename(ref.path) + " (reference GASAL2 kernels, host build)", "pairs": pairs}
    if engine:
        doc["engine"] = engine
    json.dump(doc, open(os.path.join(OUT, f"pairs_{name}.json"), "w"))
    print(name, len(pairs))


def round2(r, r512):
    """Corners added in round 2 (tests/test_gpu_round2.py uses the same generators at larger sizes)."""
    for alpha in (b"AC", b"AAAC", b"A", b"ACGTN"):
        dump("r2_ties16_" + alpha.decode(), W.tie_dense_pairs(30, 257, 496, alpha, seed=310 + len(alpha)), r)
    dump("r2_ties8_AC", W.tie_dense_pairs(60, 100, 256, b"AC", seed=320, tmax=400), r)
    dump("r2_saturation", W.saturation_pairs(), r512, engine={"max_query_len": 512})
    # windows of 2001..2049 bases are beyond MAX_TARGET_LEN for the reference's caller, but its kernels handle them
    dump("r2_window_edges", W.window_edge_pairs(), r, engine={"max_target_len": 2100})


if __name__ == "__main__":
    r = oracle.reference()
    r512 = oracle.reference(512)
    assert r is not None and r512 is not None, "build oracle/_ref first (make -C oracle)"
    if len(sys.argv) > 1 and sys.argv[1] == "r2":
        round2(r, r512)
        sys.exit(0)
    dump("probes", W.from_lists([b"ACGTNCGTAC", b"ACGTACGTAC", b"AAAA", b"NNNN", b"ACGT", b"ACGTACGTACGT", b"acgtacgt"],
                                [b"ACGTACGTAC", b"ACGTACGTAC", b"CCCC", b"ACGT", b"TTACGTTT", b"GGACGTACGTACGTCC", b"ACGTACGT"]), r)
    dump("adversarial_acgtn", W.adversarial_pairs(400, seed=201), r)
    dump("adversarial_ac_ties", W.adversarial_pairs(400, seed=202, alphabet=b"AC"), r)
    dump("adversarial_iupac", W.adversarial_pairs(300, seed=203, alphabet=b"ACGTNacgtnRYKMSW.-"), r)
    dump("ext150", W.extension_pairs(150, seed=204), r)
    dump("ext150_N", W.extension_pairs(120, seed=205, n_rate=0.01), r)
    dump("ext250_indel", W.extension_pairs(80, seed=206, read_len=250, indel_rate=0.05, max_indel=4, fixed_query_len=False), r)
    dump("long_q500_t2000", W.adversarial_pairs(12, seed=207, max_q=500, max_t=2000), r512)
    dump("scoring_1_4_6_2", W.adversarial_pairs(300, seed=208), r, dict(match=1, mismatch=4, gap_open=6, gap_extend=2))
    round2(r, r512)
