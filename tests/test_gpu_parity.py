"""GPU suite: the CUDA path through the C ABI against the oracle, bit-exact (score, start/end, n_ops,
CIGAR text) on seeded inputs, plus size-independent properties at larger sizes."""
import numpy as np
import pytest

from rabbitsalign_b200 import workload as W
from parity_util import compare, oracle_arrays

pytestmark = pytest.mark.gpu

CASES = [
    ("adv_acgtn", W.adversarial_pairs, dict(n=6000, seed=101)),
    ("adv_ac", W.adversarial_pairs, dict(n=6000, seed=102, alphabet=b"AC")),
    ("adv_iupac", W.adversarial_pairs, dict(n=4000, seed=103, alphabet=b"ACGTNacgtnRYKMSW.-")),
    ("adv_mid", W.adversarial_pairs, dict(n=3000, seed=104, max_q=200, max_t=400)),
    ("ext150", W.extension_pairs, dict(n=3000, seed=105)),
    ("ext150_var", W.extension_pairs, dict(n=2000, seed=106, fixed_query_len=False, indel_rate=0.01)),
    ("ext250_indel", W.extension_pairs, dict(n=1500, seed=107, read_len=250, indel_rate=0.05, max_indel=4, fixed_query_len=False)),
    ("ext150_N", W.extension_pairs, dict(n=2000, seed=108, n_rate=0.01)),
    ("long", W.adversarial_pairs, dict(n=200, seed=109, max_q=500, max_t=2000)),
    ("homopolymer", W.adversarial_pairs, dict(n=4000, seed=110, alphabet=b"AAAC", max_q=120, max_t=200)),
]


@pytest.mark.parametrize("name,gen,kw", CASES, ids=[c[0] for c in CASES])
def test_exact_kernel_parity(engine_exact, oracle_lib, name, gen, kw):
    b = gen(**kw)
    res = engine_exact.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    bad = compare(engine_exact, res, oracle_arrays(oracle_lib, b), b)
    assert not bad, "\n".join(bad)


@pytest.mark.parametrize("name,gen,kw", CASES, ids=[c[0] for c in CASES])
def test_packed_kernel_parity(engine, oracle_lib, name, gen, kw):
    b = gen(**kw)
    res = engine.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    bad = compare(engine, res, oracle_arrays(oracle_lib, b), b)
    assert not bad, "\n".join(bad)
    st = engine.stats()
    assert st["kernel_launches"] > 0


def test_solve_ssw_on_gpu_shape(engine, oracle_lib):
    qs = [b"ACGTNCGTAC", b"ACGTACGTAC", b"AAAA", b"NNNN", b"ACGT"]
    ts = [b"ACGTACGTAC", b"ACGTACGTAC", b"CCCC", b"ACGT", b"TTACGTTT"]
    got = engine.solve_ssw_on_gpu(qs, ts)
    assert [g.astuple() for g in got] == [o.astuple() for o in oracle_lib.align(qs, ts)]


def test_query_too_long_is_an_error(engine):
    from rabbitsalign_b200 import ExtensionError
    with pytest.raises(ExtensionError) as ei:
        engine.solve_ssw_on_gpu([b"A" * 501], [b"A" * 600])
    assert ei.value.status == -3
    # the handle stays usable
    assert engine.solve_ssw_on_gpu([b"ACGT"], [b"ACGT"])[0].score == 8


def test_window_longer_than_max_is_flagged(engine):
    b = W.from_lists([b"ACGT" * 10, b"ACGT" * 10], [b"ACGT" * 600, b"ACGT" * 20])
    res = engine.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    assert res["status"][0] == 1 and res["n_ops"][0] == 0 and res["status"][1] == 0 and res["score"][1] == 80


def test_other_scoring(oracle_lib):
    from rabbitsalign_b200 import ExtensionEngine
    kw = dict(match=1, mismatch=4, gap_open=6, gap_extend=2)
    e = ExtensionEngine(**kw)
    b = W.adversarial_pairs(4000, seed=120, max_q=150, max_t=250)
    res = e.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    bad = compare(e, res, oracle_arrays(oracle_lib, b, **kw), b)
    e.close()
    assert not bad, "\n".join(bad)


def test_chunked_batch_equals_single_chunks(oracle_lib):
    """A batch forced through many small chunks (tiny scratch budget) gives the same records."""
    from rabbitsalign_b200 import ExtensionEngine
    b = W.extension_pairs(1500, seed=130, fixed_query_len=False)
    e = ExtensionEngine(scratch_bytes=24 << 20)
    res = e.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    bad = compare(e, res, oracle_arrays(oracle_lib, b), b)
    e.close()
    assert not bad, "\n".join(bad)


def test_resident_legs_match_submit(engine):
    b = W.extension_pairs(2000, seed=140)
    res = engine.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff).copy()
    engine.stage_resident(b.qbuf, b.qoff, b.tbuf, b.toff)
    engine.run_resident()
    engine.run_resident()  # idempotent: re-running the kernels over the same inputs
    res2 = engine.fetch_resident(b.n)
    assert res.tobytes() == res2.tobytes()
    st = engine.stats()
    assert st["dp_ms"] > 0 and st["cells"] == b.cells


def test_large_batch_properties(engine):
    """Full-size-style check without the oracle: every accepted record's CIGAR consumes exactly the
    reported spans and re-scores to the reported score (N-free input; SURVEY.md 8a invariants)."""
    import re
    b = W.fixed_pairs_fast(200_000, qlen=150, tlen=250, sub_rate=0.02, seed=150)
    res = engine.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    assert (res["status"] == 0).all() and (res["score"] > 0).all()
    # order independence: the same pairs reversed give the same records reversed
    idx = np.arange(b.n)[::-1]
    q2 = b.qbuf.reshape(b.n, 150)[idx].reshape(-1).copy()
    t2 = b.tbuf.reshape(b.n, 250)[idx].reshape(-1).copy()
    res2 = engine.align_packed(q2, b.qoff, t2, b.toff)
    assert res2[::-1].tobytes() == res.tobytes()
    rng = np.random.default_rng(1)
    q = b.qbuf.reshape(b.n, 150)
    t = b.tbuf.reshape(b.n, 250)
    for i in rng.integers(0, b.n, size=3000):
        cig = engine.cigar(res, int(i))
        ii, jj, s = int(res["ref_start"][i]), int(res["query_start"][i]), 0
        for cnt, op in re.findall(r"(\d+)([MXDI])", cig):
            cnt = int(cnt)
            if op in "MX":
                eq = q[i, jj:jj + cnt] == t[i, ii:ii + cnt]
                assert eq.all() if op == "M" else (~eq).all()
                s += 2 * cnt if op == "M" else -8 * cnt
                ii += cnt
                jj += cnt
            else:
                s -= 12 + (cnt - 1)
                if op == "D":
                    ii += cnt
                else:
                    jj += cnt
        assert s == res["score"][i] and ii == res["ref_end"][i] + 1 and jj == res["query_end"][i] + 1


def _golden_files():
    import os
    d = os.path.join(os.path.dirname(__file__), "golden")
    return sorted(f for f in os.listdir(d) if f.startswith("pairs_") and f.endswith(".json"))


@pytest.mark.parametrize("name", _golden_files())
@pytest.mark.parametrize("exact_only", [False, True], ids=["packed", "exact"])
def test_golden_vectors_from_reference_kernels(name, exact_only):
    """Committed vectors produced by the reference's own GASAL2 kernels (oracle/make_golden.py)."""
    import json
    import os
    from rabbitsalign_b200 import ExtensionEngine
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", name)))
    e = ExtensionEngine(exact_only=exact_only, **g.get("scoring", {}), **g.get("engine", {}))
    got = e.solve_ssw_on_gpu([p["q"].encode("latin1") for p in g["pairs"]], [p["t"].encode("latin1") for p in g["pairs"]])
    e.close()
    bad = [(p["q"], p["t"], list(r.astuple()), p["res"]) for p, r in zip(g["pairs"], got) if list(r.astuple()) != p["res"]]
    assert not bad, bad[:3]


def test_redo_scratch_exhaustion_is_retried(oracle_lib):
    """Every query carries a symbol outside ACGTN, so the packed kernel flags all pairs for the exact redo pass;
    their tiles (32 KB each) exceed the redo head-room, the surplus comes back with status 4 inside the engine
    and rsa_ext_wait re-submits it exact-only.  The caller must see plain, exact records."""
    from rabbitsalign_b200 import ExtensionEngine
    b = W.extension_pairs(3000, seed=160)
    q = b.qbuf.copy()
    q[b.qoff[:-1] + 75] = ord("R")  # IUPAC purine: nibble 2, equal to nothing in ACGT
    b = W.PairBatch(q, b.qoff, b.tbuf, b.toff)
    e = ExtensionEngine()
    res = e.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    st = e.stats()
    bad = compare(e, res, oracle_arrays(oracle_lib, b), b)
    e.close()
    assert not bad, "\n".join(bad)
    assert st["pairs_redo"] >= 3000  # all pairs went through the redo pass at least once


def test_concurrent_handles_from_threads(oracle_lib):
    """Several host workers, one handle each (the reference: one GASAL stream per worker thread)."""
    import threading
    from rabbitsalign_b200 import ExtensionEngine
    batches = [W.extension_pairs(700, seed=170 + k, fixed_query_len=(k % 2 == 0)) for k in range(4)]
    out = [None] * 4

    def work(k):
        e = ExtensionEngine()
        for _ in range(3):
            res = e.align_packed(batches[k].qbuf, batches[k].qoff, batches[k].tbuf, batches[k].toff)
        out[k] = compare(e, res, oracle_arrays(oracle_lib, batches[k]), batches[k])
        e.close()
    ths = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert all(o == [] for o in out), out


def _random_related_pairs(rng, shapes):
    qs, ts = [], []
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for ql, tl in shapes:
        t = acgt[rng.integers(0, 4, size=tl)]
        if tl >= ql and rng.random() < 0.8:
            a = int(rng.integers(0, tl - ql + 1))
            q = t[a:a + ql].copy()
        else:
            q = acgt[rng.integers(0, 4, size=ql)]
            k = min(ql, tl)
            q[:k] = t[:k]
        m = rng.random(ql) < 0.04
        q[m] = acgt[rng.integers(0, 4, size=int(m.sum()))]
        if ql > 12 and rng.random() < 0.5:  # one indel
            p = int(rng.integers(3, ql - 3))
            q = np.concatenate([q[:p], acgt[rng.integers(0, 4, size=2)], q[p:ql - 2]]) if rng.random() < 0.5 \
                else np.concatenate([q[:p], q[p + 2:], acgt[rng.integers(0, 4, size=2)]])
        qs.append(q.astype(np.uint8).tobytes())
        ts.append(t.astype(np.uint8).tobytes())
    return W.from_lists(qs, ts)


def test_every_query_length(engine, oracle_lib):
    """Each |q| in 1..500 (every column count C and every wide/narrow lane split of the 8- and 16-lane packed
    kernels, plus the shapes handed to the exact kernel), four windows each."""
    rng = np.random.default_rng(180)
    shapes = [(ql, int(rng.integers(max(1, ql - 20), ql + 120))) for ql in range(1, 501) for _ in range(4)]
    b = _random_related_pairs(rng, shapes)
    res = engine.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    bad = compare(engine, res, oracle_arrays(oracle_lib, b), b)
    assert not bad, "\n".join(bad)


def test_window_length_extremes(engine, oracle_lib):
    """|t| from 1 (fewer rows than lanes, partial 4-row blocks) up to the packed kernel's limit and beyond."""
    rng = np.random.default_rng(181)
    shapes = [(int(rng.integers(8, 160)), tl) for tl in list(range(1, 80)) * 2]
    shapes += [(150, tl) for tl in (1023, 1024, 1025, 1999, 2000)] + [(64, 2000), (8, 2000), (256, 1500)]
    b = _random_related_pairs(rng, shapes)
    res = engine.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    bad = compare(engine, res, oracle_arrays(oracle_lib, b), b)
    assert not bad, "\n".join(bad)


def test_single_pair_batches(engine, oracle_lib):
    for q, t in [(b"ACGTACGTACGTACGT", b"TTACGTACGTACGTACGTTT"), (b"A", b"A"), (b"ACGTACGTAC" * 25, b"ACGTACGTAC" * 30)]:
        got = engine.solve_ssw_on_gpu([q], [t])[0].astuple()
        assert got == oracle_lib.align([q], [t])[0].astuple()


def test_plan_ahead_equals_inline_slices(engine, oracle_lib):
    """A batch above the plan-ahead threshold (helper thread plans the chunks) gives the records of the same pairs
    sent in slices small enough to be planned inline; a sample is also checked against the oracle."""
    b = W.extension_pairs(90_000, seed=170, fixed_query_len=False, indel_rate=0.01, n_rate=0.002)
    whole = engine.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff).copy()
    parts = []
    step = 30_000  # <= 32768: inline planning
    for lo in range(0, b.n, step):
        part = b.slice(lo, min(b.n, lo + step))
        parts.append(engine.align_packed(part.qbuf, part.qoff, part.tbuf, part.toff).copy())
    sliced = np.concatenate(parts)
    fields = ["score", "query_start", "query_end", "ref_start", "ref_end", "n_ops", "status"]
    for f in fields:
        assert (whole[f] == sliced[f]).all(), f
    short = whole["n_ops"] <= 40
    assert (whole["rle"][short] == sliced["rle"][short]).all()
    sub = b.slice(0, 4000)
    first = engine.align_packed(sub.qbuf, sub.qoff, sub.tbuf, sub.toff)
    for f in fields:
        assert (first[f] == whole[f][:4000]).all(), f
    bad = compare(engine, first, oracle_arrays(oracle_lib, sub), sub)
    assert not bad, "\n".join(bad)


def test_large_then_small_then_large_on_one_handle(engine):
    """The helper thread is reused across batches and idle for small ones."""
    big = W.fixed_pairs_fast(70_000, qlen=150, tlen=220, sub_rate=0.02, seed=171)
    small = W.extension_pairs(300, seed=172)
    r1 = engine.align_packed(big.qbuf, big.qoff, big.tbuf, big.toff).copy()
    rs = engine.align_packed(small.qbuf, small.qoff, small.tbuf, small.toff).copy()
    r2 = engine.align_packed(big.qbuf, big.qoff, big.tbuf, big.toff).copy()
    rs2 = engine.align_packed(small.qbuf, small.qoff, small.tbuf, small.toff).copy()
    assert r1.tobytes() == r2.tobytes() and rs.tobytes() == rs2.tobytes()
    assert (r1["status"] == 0).all() and (r1["score"] > 0).all()


def test_query_too_long_inside_a_large_batch(engine):
    from rabbitsalign_b200 import ExtensionError
    b = W.fixed_pairs_fast(50_000, qlen=150, tlen=200, seed=173)
    qs = [bytes(b.qbuf[b.qoff[i]:b.qoff[i + 1]]) for i in range(40_000)] + [b"A" * 501]
    ts = [bytes(b.tbuf[b.toff[i]:b.toff[i + 1]]) for i in range(40_000)] + [b"A" * 600]
    bb = W.from_lists(qs, ts)
    with pytest.raises(ExtensionError) as ei:
        engine.align_packed(bb.qbuf, bb.qoff, bb.tbuf, bb.toff)
    assert ei.value.status == -3
    assert engine.solve_ssw_on_gpu([b"ACGT"], [b"ACGT"])[0].score == 8


def test_concurrent_workers_wide_classes_are_deterministic():
    """16 handles on 16 threads, 512-pair slices of 250-bp variable-length indel-rich pairs (column classes whose
    shared memory exceeds 48 KB, several classes per call): every record equals the single-handle baseline.
    Regression test for the per-launch cudaFuncSetAttribute race (a lost launch handed back failed-looking records)."""
    import threading
    from rabbitsalign_b200 import ExtensionEngine
    b = W.extension_pairs(48_000, seed=910, read_len=250, indel_rate=0.02, max_indel=4, sub_rate=0.02, fixed_query_len=False)
    e0 = ExtensionEngine()
    base = e0.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff).copy()
    e0.close()
    assert (base["status"] == 0).all()
    slices = [(lo, min(b.n, lo + 512)) for lo in range(0, b.n, 512)]
    parts = [b.slice(lo, hi) for lo, hi in slices]
    T = 16
    engines = [ExtensionEngine() for _ in range(T)]
    bad, errors, lock = [], [], threading.Lock()
    fields = ["score", "query_start", "query_end", "ref_start", "ref_end", "n_ops", "status"]

    def worker(w, order):
        for k in order:
            p = parts[k]
            try:
                r = engines[w].align_packed(p.qbuf, p.qoff, p.tbuf, p.toff)
            except Exception as ex:  # noqa: BLE001
                with lock:
                    errors.append(str(ex))
                return
            lo, hi = slices[k]
            diff = np.zeros(hi - lo, bool)
            for f in fields:
                diff |= r[f] != base[f][lo:hi]
            if diff.any():
                with lock:
                    bad.extend((lo + np.nonzero(diff)[0]).tolist())

    rng = np.random.default_rng(7)
    for _ in range(3):
        perm = rng.permutation(len(parts))
        th = [threading.Thread(target=worker, args=(w, perm[w::T])) for w in range(T)]
        [t.start() for t in th]
        [t.join() for t in th]
    for e in engines:
        e.close()
    assert not errors, errors[:3]
    assert not bad, f"{len(bad)} records differ from the baseline, e.g. pairs {bad[:5]}"


def test_reference_windows_equal_explicit_windows(engine, oracle_lib):
    """Windows named by (offset, length) in a resident reference give the records of the same windows sent as bytes
    (no window strings are built or copied); a sample is checked against the oracle too.
    60 000 pairs: the plan-ahead path; IUPAC symbols in the reference exercise the redo pass through the window form."""
    rng = np.random.default_rng(77)
    ref = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=3_000_000)
    ref[rng.integers(0, len(ref), size=300)] = ord("N")
    ref[rng.integers(0, len(ref), size=100)] = ord("R")
    n = 60_000
    win_len = rng.integers(150, 420, size=n).astype(np.int32)
    win_off = rng.integers(0, len(ref) - 500, size=n).astype(np.int64)
    qs, ts = [], []
    comp = {65: 84, 67: 71, 71: 67, 84: 65}
    for i in range(n):
        w = ref[win_off[i]:win_off[i] + win_len[i]]
        lo = int(rng.integers(0, max(1, len(w) - 150)))
        q = w[lo:lo + 150].copy()
        mut = rng.random(len(q)) < 0.02
        q[mut] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(mut.sum()))
        if i % 7 == 0 and len(q) > 40:  # a deletion in the read
            q = np.concatenate([q[:30], q[33:]])
        qs.append(q.tobytes())
        ts.append(w.tobytes())
    b = W.from_lists(qs, ts)
    explicit = engine.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff).copy()
    engine.set_reference(ref)
    byref = engine.align_ref_windows(b.qbuf, b.qoff, win_off, win_len)
    for f in ["score", "query_start", "query_end", "ref_start", "ref_end", "n_ops", "status"]:
        assert (explicit[f] == byref[f]).all(), f
    short = explicit["n_ops"] <= 40
    assert (explicit["rle"][short] == byref["rle"][short]).all()
    sub = b.slice(0, 3000)
    bad = compare(engine, byref[:3000], oracle_arrays(oracle_lib, sub), sub)
    assert not bad, "\n".join(bad)
    # a window outside the reference is an argument error, and the handle stays usable
    from rabbitsalign_b200 import ExtensionError
    with pytest.raises(ExtensionError):
        engine.align_ref_windows(b.qbuf[:b.qoff[1]], b.qoff[:2], np.array([len(ref) - 10]), np.array([100]))
    assert engine.solve_ssw_on_gpu([b"ACGT"], [b"ACGT"])[0].score == 8
