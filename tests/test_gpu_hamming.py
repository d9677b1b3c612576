"""GPU suite, SURVEY 8(f) rank 3 (first half): the Hamming shortcut on the device (rsa_ext_hamming_align /
rsa_ext_hamming_ref_windows, csrc/kernels_hamming.cuh) against the reference's own hamming_distance / hamming_align
compiled from /root/reference (oracle/_ref/libssw_ref_*.so) and against the C restatement (oracle/sw_oracle.c)."""
import numpy as np
import pytest

import oracle
from hamming_util import make_pairs, record_tuple
from rabbitsalign_b200 import ExtensionEngine
from rabbitsalign_b200.ext import alninfo_cigar_string

pytestmark = pytest.mark.gpu


def expected(qbuf, qoff, tbuf, toff, **kw):
    exp = oracle.hamming_reference(qbuf, qoff, tbuf, toff, **kw)
    if exp is None:  # GPU box without oracle/_ref: the restatement (pinned to the reference in the CPU suite)
        r = oracle.hamming_restatement(qbuf, qoff, tbuf, toff, **kw)
        exp = dict(hamming=r["hamming"], status=r["status"], score=r["score"], ed=r["ed"], qs=r["start"], qe=r["end"],
                   rs=r["start"], re=r["end"], cigar=r["cigar"])
    return exp


@pytest.mark.parametrize("seed,read_len,kw", [(11, 150, {}), (12, 250, {}), (13, 100, dict(match=1, mismatch=4, end_bonus=3)),
                                               (14, 150, dict(end_bonus=0))])
def test_hamming_shortcut_equals_reference(seed, read_len, kw):
    qbuf, qoff, tbuf, toff = make_pairs(6000, seed, read_len)
    exp = expected(qbuf, qoff, tbuf, toff, **kw)
    e = ExtensionEngine(match=kw.get("match", 2), mismatch=kw.get("mismatch", 8))
    ham, aln = e.hamming_align(qbuf, qoff, tbuf, toff, end_bonus=kw.get("end_bonus", 10))
    e.close()
    n = len(qoff) - 1
    assert (ham == exp["hamming"]).all()
    bad = []
    taken = 0
    for i in range(n):
        got = record_tuple(aln, i, alninfo_cigar_string)
        if exp["status"][i] != 0:
            want = (1, 0, 0, 0, 0, 0, 0, "")
        else:
            want = (0, int(exp["score"][i]), int(exp["ed"][i]), int(exp["qs"][i]), int(exp["qe"][i]), int(exp["rs"][i]),
                    int(exp["re"][i]), exp["cigar"][i])
            taken += 1
        if got[0] == 3:  # more runs than the inline record holds: host path, nothing to compare
            assert exp["status"][i] == 0 and exp["cigar"][i].count("X") + exp["cigar"][i].count("=") + exp["cigar"][i].count("S") > 25
            continue
        if got != want:
            bad.append((i, got, want))
    assert not bad, bad[:5]
    assert taken > n // 4 and taken < n  # both sides of the decision are exercised


def test_hamming_windows_in_the_resident_reference():
    rng = np.random.default_rng(5)
    ref = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 200_000)].copy()
    n, L = 20_000, 150
    win = rng.integers(0, len(ref) - L, n).astype(np.int64)
    win[0], win[1] = 0, len(ref) - L
    reads = np.stack([ref[w:w + L] for w in win]).copy()
    flips = rng.random(reads.shape) < rng.choice([0.0, 0.02, 0.045, 0.055, 0.2], size=(n, 1))
    reads[flips] = np.frombuffer(b"TGCA", np.uint8)[np.searchsorted(np.frombuffer(b"ACGT", np.uint8), reads[flips])]
    qbuf = np.ascontiguousarray(reads.reshape(-1))
    qoff = (np.arange(n + 1) * L).astype(np.int64)
    tbuf = np.ascontiguousarray(np.stack([ref[w:w + L] for w in win]).reshape(-1))
    e = ExtensionEngine()
    e.set_reference(ref)
    ham_w, aln_w = e.hamming_ref_windows(qbuf, qoff, win)
    ham_e, aln_e = e.hamming_align(qbuf, qoff, tbuf, qoff)
    e.close()
    assert (ham_w == ham_e).all() and aln_w.tobytes() == aln_e.tobytes()
    exp = expected(qbuf, qoff, tbuf, qoff)
    assert (ham_w == exp["hamming"]).all()
    assert ((aln_w["status"] == 0) == (exp["status"] == 0)).all()
    ok = exp["status"] == 0
    assert (aln_w["sw_score"][ok] == exp["score"][ok]).all() and (aln_w["edit_distance"][ok] == exp["ed"][ok]).all()
    assert 0.3 < ok.mean() < 0.8


def test_pinned_and_pageable_callers_get_the_same_records():
    """Pinned host arrays are copied from directly, pageable ones go through the handle's bounce buffers (and a mix of the
    two is decided per array): the records must not depend on it."""
    import torch

    def pinned(a):
        t = torch.empty(a.nbytes, dtype=torch.uint8).pin_memory()
        v = t.numpy().view(a.dtype).reshape(a.shape)
        v[...] = a
        return t, v
    qbuf, qoff, tbuf, toff = make_pairs(5000, 21, 150)
    e = ExtensionEngine()
    ham0, aln0 = e.hamming_align(qbuf, qoff, tbuf, toff)                      # all pageable
    keep = [pinned(x) for x in (qbuf, qoff, tbuf, toff)]
    ham1, aln1 = e.hamming_align(*[v for _, v in keep])                        # inputs pinned, outputs pageable
    ham2, aln2 = e.hamming_align(keep[0][1], qoff, keep[2][1], toff)           # sequences pinned, offsets pageable
    e.close()
    assert (ham0 == ham1).all() and (ham0 == ham2).all()
    assert aln0.tobytes() == aln1.tobytes() == aln2.tobytes()
