"""GPU suite, round 2: regression tests for the advisor's findings (poll-driven multi-chunk batches, the status-4
re-run entered through submit_ptrs / window form, byte-deterministic records with long CIGARs) and the parity corners
the round-1 review listed (16-lane path with tie-dense inputs, near-saturation scores and the packed->exact hand-off at
match*|q| = 1023/1024, |t| = 2047/2048, 5 % indel 250-bp pairs against the reference's own GPU path)."""
import time

import numpy as np
import pytest

import oracle
from rabbitsalign_b200 import ExtensionEngine, workload as W
from parity_util import compare, oracle_arrays

pytestmark = pytest.mark.gpu
FIELDS = ["score", "query_start", "query_end", "ref_start", "ref_end", "n_ops", "status"]


def test_poll_drives_a_multichunk_batch(oracle_lib):
    """the reference drives its batch with `while (poll) usleep` (src/gasal2_ssw.cpp:179).  With a
    tiny scratch budget the batch needs many more chunks than the engine has slots; poll alone must finish it."""
    b = W.extension_pairs(6000, seed=201, fixed_query_len=False)
    e = ExtensionEngine(scratch_bytes=24 << 20)
    base = e.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff).copy()
    res = e.submit(b.qbuf, b.qoff, b.tbuf, b.toff)
    t0 = time.time()
    polls = 0
    while e.poll():
        polls += 1
        time.sleep(0.0001)
        assert time.time() - t0 < 60, "poll never reached 0 (livelock)"
    e.wait()  # finalises the batch; must not have anything left to wait for
    assert polls >= 0
    assert res.tobytes() == base.tobytes()
    bad = compare(e, res, oracle_arrays(oracle_lib, b), b)
    e.close()
    assert not bad, "\n".join(bad)


def _iupac_batch(n, seed):
    b = W.extension_pairs(n, seed=seed)
    q = b.qbuf.copy()
    q[b.qoff[:-1] + 75] = ord("R")  # every pair needs the exact redo pass
    return W.PairBatch(q, b.qoff, b.tbuf, b.toff)


def test_retry_entered_through_submit_ptrs(oracle_lib):
    """the status-4 re-run used to go through rsa_ext_submit_ptrs again and copied from the handle's
    own pinned staging into itself.  Enter through submit_ptrs (what the veneer does) and check every record."""
    b = _iupac_batch(3000, 202)
    e = ExtensionEngine()
    res = e.align_ptrs(b.queries(), b.targets())
    st = e.stats()
    bad = compare(e, res, oracle_arrays(oracle_lib, b), b)
    e.close()
    assert not bad, "\n".join(bad)
    assert st["pairs_redo"] >= 3000


def test_retry_entered_through_reference_windows(oracle_lib):
    rng = np.random.default_rng(203)
    ref = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=1_000_000)
    n = 3000
    win_len = rng.integers(230, 260, size=n).astype(np.int32)
    win_off = rng.integers(0, len(ref) - 300, size=n).astype(np.int64)
    qs, ts = [], []
    for i in range(n):
        w = ref[win_off[i]:win_off[i] + win_len[i]]
        q = w[40:190].copy()
        q[75] = ord("R")
        qs.append(q.tobytes())
        ts.append(w.tobytes())
    b = W.from_lists(qs, ts)
    e = ExtensionEngine()
    e.set_reference(ref)
    res = e.align_ref_windows(b.qbuf, b.qoff, win_off, win_len)
    bad = compare(e, res, oracle_arrays(oracle_lib, b), b)
    e.close()
    assert not bad, "\n".join(bad)


def test_long_cigar_records_are_byte_deterministic(oracle_lib):
    """records with more than 40 RLE bytes used to carry a scheduling-dependent arena offset."""
    b = W.extension_pairs(4000, seed=204, read_len=400, indel_rate=0.06, max_indel=2, sub_rate=0.04, fixed_query_len=False)
    e = ExtensionEngine()
    r1 = e.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff).copy()
    assert (r1["n_ops"] > 40).sum() > 50, "the case must contain long CIGARs"
    r2 = e.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff).copy()
    assert r1.tobytes() == r2.tobytes()
    e.stage_resident(b.qbuf, b.qoff, b.tbuf, b.toff)
    e.run_resident()
    r3 = e.fetch_resident(b.n)
    assert r1.tobytes() == r3.tobytes()
    # the inline bytes are the first 40 bytes of the full string
    e.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    for i in np.nonzero(r1["n_ops"] > 40)[0][:50]:
        assert (e.full_rle(r1, int(i))[:40] == r1["rle"][i]).all()
    bad = compare(e, r1, oracle_arrays(oracle_lib, b), b)
    e.close()
    assert not bad, "\n".join(bad)


# ---- parity corners -----------------------------------------------------------------------------------------------

@pytest.mark.parametrize("alphabet", [b"AC", b"AAAC", b"A", b"ACGTN"], ids=["AC", "AAAC", "A", "ACGTN"])
def test_sixteen_lane_path_tie_dense(engine, oracle_lib, alphabet):
    """|q| 257..500 (16-lane groups) over alphabets where equal maxima and co-optimal paths are everywhere."""
    b = W.tie_dense_pairs(160, 257, 500, alphabet, seed=210 + len(alphabet))
    res = engine.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    bad = compare(engine, res, oracle_arrays(oracle_lib, b), b)
    assert not bad, "\n".join(bad)
    assert engine.stats()["pairs_fast"] == b.n


def test_near_saturation_scores_and_exact_handoff(oracle_lib):
    """Perfect and near-perfect matches at |q| in {496, 500, 511, 512}: scores 992..1022 stay on the packed kernel (the
    maximum-tracking key (score << 5 | column) is a positive s16 up to 1023), 2 * 512 = 1024 goes to the exact kernel."""
    b = W.saturation_pairs()
    e = ExtensionEngine(max_query_len=512)
    res = e.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    st = e.stats()
    bad = compare(e, res, oracle_arrays(oracle_lib, b), b)
    e.close()
    assert not bad, "\n".join(bad)
    assert res["score"].max() == 1024 and (res["score"] >= 1000).sum() >= 8
    assert st["pairs_exact"] == 6 and st["pairs_fast"] == 18  # |q| = 512 is handed to the exact kernel statically


def test_window_length_2047_and_2048(oracle_lib):
    """|t| = 2047 is the packed kernel's last row count, 2048 the exact kernel's first."""
    b = W.window_edge_pairs()
    e = ExtensionEngine(max_target_len=2100)
    res = e.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    st = e.stats()
    bad = compare(e, res, oracle_arrays(oracle_lib, b), b)
    e.close()
    assert not bad, "\n".join(bad)
    assert st["pairs_fast"] == 4 and st["pairs_exact"] == 4


def test_250bp_5pct_indel_50k_pairs_against_reference_gpu(engine):
    """BASELINE configs[3]: 250-bp reads at a 5 % indel-event rate, 50 000 pairs, every field and the CIGAR text against
    the reference's own CUDA path (GASAL2 + src/gasal2_ssw.cpp for sm_100a) on the same GPU."""
    ref_gpu = oracle.reference_gpu()
    if ref_gpu is None:
        pytest.skip("oracle/_ref/libgasal_gpu.so not built (needs /root/reference at build time)")
    b = W.extension_pairs(50_000, seed=230, read_len=250, indel_rate=0.05, max_indel=4, fixed_query_len=False)
    ql = np.diff(b.qoff)
    tl = np.diff(b.toff)
    assert (ql <= 496).all() and (tl <= 2000).all()
    res = engine.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    n_bad, first = 0, []
    for lo in range(0, b.n, 512):  # the reference takes STREAM_BATCH_SIZE = 512 pairs per call
        p = b.slice(lo, min(b.n, lo + 512))
        out5, texts = ref_gpu.batch(p.qbuf, p.qoff, p.tbuf, p.toff, thread_id=2, cigar_stride=2048)
        for k in range(p.n):
            i = lo + k
            got = (int(res["score"][i]), int(res["query_start"][i]), int(res["query_end"][i]), int(res["ref_start"][i]),
                   int(res["ref_end"][i]), engine.cigar(res, i))
            exp = tuple(int(x) for x in out5[k]) + (texts[k],)
            if got != exp:
                n_bad += 1
                if len(first) < 5:
                    first.append(f"pair {i}: product {got} reference-gpu {exp}")
    assert n_bad == 0, f"{n_bad} of {b.n} differ\n" + "\n".join(first)
    assert np.mean(res["n_ops"]) > 15  # long CIGARs, as the config intends


# ---- device-side planning (VERDICT r1 task 3) ------------------------------------------------------------------------

def _mixed_batch(seed, n_main=60_000):
    """Every routing class in one batch: packed 8-lane and 16-lane groups of many lengths, exact-kernel shapes (|q| < 8,
    |t| > 2047), pairs that are not aligned (empty query / window, window over the limit), N and IUPAC symbols."""
    rng = np.random.default_rng(seed)
    a = W.extension_pairs(n_main, seed=seed, fixed_query_len=False, indel_rate=0.01, n_rate=0.002)
    b = W.adversarial_pairs(6000, seed=seed + 1, max_q=40, max_t=90)                    # includes |q| < 8
    c = W.extension_pairs(3000, seed=seed + 2, read_len=300, fixed_query_len=False)     # 16-lane groups
    d = W.adversarial_pairs(2000, seed=seed + 3, alphabet=b"ACGTNRY")                   # redo pass
    qs = a.queries() + b.queries() + c.queries() + d.queries()
    ts = a.targets() + b.targets() + c.targets() + d.targets()
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    qs += [b"", b"ACGTACGTAC", acgt[rng.integers(0, 4, size=150)].tobytes(), acgt[rng.integers(0, 4, size=100)].tobytes()]
    ts += [b"ACGT", b"", acgt[rng.integers(0, 4, size=2300)].tobytes(), acgt[rng.integers(0, 4, size=2100)].tobytes()]
    order = rng.permutation(len(qs))
    return W.from_lists([qs[i] for i in order], [ts[i] for i in order])


def test_device_planner_equals_host_planner(oracle_lib):
    b = _mixed_batch(240)
    dev = ExtensionEngine(max_target_len=2200)
    host = ExtensionEngine(max_target_len=2200, host_plan=True)
    dev.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)   # (first call: buffers are allocated inside the timed host pass)
    host.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    rd = dev.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff).copy()
    st_dev = dev.stats()
    rh = host.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff).copy()
    st_host = host.stats()
    assert rd.tobytes() == rh.tobytes()
    for k in ("pairs_fast", "pairs_exact", "pairs_failed", "cells"):
        assert st_dev[k] == st_host[k], k
    assert st_dev["h2d_bytes"] < st_host["h2d_bytes"]          # 16 B of offsets per pair instead of the 45-byte blob
    assert st_dev["host_plan_ms"] < st_host["host_plan_ms"]
    assert (rd["status"] != 0).sum() == 3  # empty query, empty window, window over max_target_len
    sub = b.slice(0, 5000)
    first = dev.align_packed(sub.qbuf, sub.qoff, sub.tbuf, sub.toff)  # (below the device planner's threshold: host-planned)
    for f in FIELDS:
        assert (first[f] == rd[f][:5000]).all(), f
    ok = np.nonzero(first["status"] == 0)[0]
    keep = W.from_lists([sub.queries()[i] for i in ok], [sub.targets()[i] for i in ok])
    res = dev.align_packed(keep.qbuf, keep.qoff, keep.tbuf, keep.toff)
    bad = compare(dev, res, oracle_arrays(oracle_lib, keep), keep)
    dev.close(); host.close()
    assert not bad, "\n".join(bad)


def test_device_planner_many_chunks_and_exact_only():
    """Small scratch budget: the device-planned batch is cut into many chunks; exact-only engines route every pair
    through the planner's exact lists."""
    b = _mixed_batch(241, n_main=30_000)
    ref = ExtensionEngine(max_target_len=2200, host_plan=True)
    want = ref.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff).copy()
    ref.close()
    e = ExtensionEngine(max_target_len=2200, scratch_bytes=600 << 20)
    got = e.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff).copy()
    e.close()
    assert got.tobytes() == want.tobytes()
    e = ExtensionEngine(max_target_len=2200, exact_only=True)
    got = e.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff).copy()
    st = e.stats()
    e.close()
    assert st["pairs_fast"] == 0
    assert got.tobytes() == want.tobytes()


def test_packed_reference_planes_and_vectorised_window_staging(oracle_lib):
    """north_star staging: the resident reference is packed to 2 bits per base + a "not ACGT" plane on
    upload and the packed DP kernel stages windows from those planes with 128-bit loads.  (1) the planes equal a numpy
    packing; (2) window batches staged from the planes give the byte-identical records of the same batches staged from
    the ASCII copy (RSA_EXT_FLAG_ASCII_WINDOWS) and of explicit windows -- with N, IUPAC and lower-case symbols in the
    reference, windows at both ends of it, every window phase mod 64, 2047-base windows and 250-bp reads; (3) a sample
    equals the oracle."""
    rng = np.random.default_rng(2024)
    ref = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=1_000_003)
    ref[rng.integers(0, len(ref), size=400)] = ord("N")
    ref[rng.integers(0, len(ref), size=60)] = ord("R")
    ref[rng.integers(0, len(ref), size=60)] = ord("n")
    low = rng.integers(0, len(ref) - 50, size=200)
    for p in low:
        ref[p:p + 20] |= 0x20  # soft-masked stretches
    e = ExtensionEngine(max_target_len=2047)
    e.set_reference(ref)
    codes, flags = e.packed_reference(len(ref))
    nib = ref & 0xF
    code = np.full(len(ref), 0xF, np.uint8)
    for k, v in ((1, 0), (3, 1), (7, 2), (4, 3), (0xE, 4)):
        code[nib == k] = v
    two = np.where(code < 4, code, np.where(code == 4, 0, 1)).astype(np.uint32)
    pad = (-len(ref)) % 64
    two = np.concatenate([two, np.zeros(pad, np.uint32)]).reshape(-1, 16)
    want_codes = (two << (2 * np.arange(16, dtype=np.uint32))).sum(axis=1).astype(np.uint32)
    fl = np.concatenate([(code >= 4).astype(np.uint32), np.zeros(pad, np.uint32)]).reshape(-1, 32)
    want_flags = (fl << np.arange(32, dtype=np.uint32)).sum(axis=1).astype(np.uint32)
    assert (codes == want_codes).all() and (flags == want_flags).all()

    n = 40_000
    read_len = np.where(np.arange(n) % 5 == 0, 250, 150)
    win_len = (read_len + rng.integers(0, 200, size=n)).astype(np.int32)
    win_len[:64] = 2047
    win_off = rng.integers(0, len(ref) - 2100, size=n).astype(np.int64)
    win_off[:64] = 5000 + np.arange(64)            # every phase of a long window
    win_off[64], win_len[64] = 0, 300              # first bases of the reference
    win_off[65] = len(ref) - 300; win_len[65] = 300  # last bases
    win_off[66] = len(ref) - 151; win_len[66] = 151
    qs, ts = [], []
    for i in range(n):
        w = ref[win_off[i]:win_off[i] + win_len[i]]
        lo = int(rng.integers(0, max(1, len(w) - read_len[i])))
        q = w[lo:lo + read_len[i]].copy()
        mut = rng.random(len(q)) < 0.02
        q[mut] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(mut.sum()))
        if i % 6 == 0 and len(q) > 60:
            q = np.concatenate([q[:40], q[42:]])
        qs.append(q.tobytes())
        ts.append(w.tobytes())
    b = W.from_lists(qs, ts)
    packed = e.align_ref_windows(b.qbuf, b.qoff, win_off, win_len).copy()
    explicit = e.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff).copy()
    e.close()
    e2 = ExtensionEngine(ascii_windows=True, max_target_len=2047)
    e2.set_reference(ref)
    ascii_form = e2.align_ref_windows(b.qbuf, b.qoff, win_off, win_len).copy()
    short = explicit["n_ops"] <= 40
    for f in ["score", "query_start", "query_end", "ref_start", "ref_end", "n_ops", "status"]:
        assert (packed[f] == explicit[f]).all(), f
        assert (packed[f] == ascii_form[f]).all(), f
    assert (packed["rle"][short] == explicit["rle"][short]).all() and (packed["rle"][short] == ascii_form["rle"][short]).all()
    assert (packed["status"] == 0).all()
    sub = b.slice(0, 2500)
    bad = compare(e2, packed[:2500], oracle_arrays(oracle_lib, sub), sub)
    e2.close()
    assert not bad, "\n".join(bad)
