"""Synthetic (query, reference-window) pair workloads shaped like the ones the reference's pipeline
hands to `solve_ssw_on_gpu` (reference src/pc.cpp:214-242 builds the extension windows:
read + |ref_span - query_span| + up to 50 bases of flank on each side; src/pc.cpp:333-368 builds the
mate-rescue windows).  Pure numpy; used by tests and bench.py.  SURVEY.md section 8(d) lists the shapes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_CODE = np.zeros(256, np.int64)
_CODE[_ACGT] = np.arange(4)


@dataclass
class PairBatch:
    """n pairs as concatenated ASCII plus int64 offsets (offsets[i]..offsets[i+1])."""

    qbuf: np.ndarray
    qoff: np.ndarray
    tbuf: np.ndarray
    toff: np.ndarray

    @property
    def n(self) -> int:
        return len(self.qoff) - 1

    @property
    def cells(self) -> int:
        ql = np.diff(self.qoff)
        tl = np.diff(self.toff)
        return int(np.sum(ql * tl))

    def queries(self) -> List[bytes]:
        b = self.qbuf.tobytes()
        return [b[self.qoff[i]:self.qoff[i + 1]] for i in range(self.n)]

    def targets(self) -> List[bytes]:
        b = self.tbuf.tobytes()
        return [b[self.toff[i]:self.toff[i + 1]] for i in range(self.n)]

    def slice(self, lo: int, hi: int) -> "PairBatch":
        qo = self.qoff[lo:hi + 1]
        to = self.toff[lo:hi + 1]
        return PairBatch(self.qbuf[qo[0]:qo[-1]].copy(), (qo - qo[0]).copy(),
                         self.tbuf[to[0]:to[-1]].copy(), (to - to[0]).copy())


def from_lists(queries: List[bytes], targets: List[bytes]) -> PairBatch:
    def pack(seqs):
        off = np.zeros(len(seqs) + 1, dtype=np.int64)
        if seqs:
            off[1:] = np.cumsum([len(s) for s in seqs])
        buf = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy()
        return buf, off
    qb, qo = pack(queries)
    tb, to = pack(targets)
    return PairBatch(qb, qo, tb, to)


def _mutate(rng: np.random.Generator, seq: np.ndarray, sub_rate: float, indel_rate: float,
            max_indel: int) -> np.ndarray:
    """Apply substitutions and indel events to one read (small python loop over events only)."""
    s = seq.copy()
    if sub_rate > 0:
        m = rng.random(len(s)) < sub_rate
        k = int(m.sum())
        if k:
            # substitute with a different base
            shift = rng.integers(1, 4, size=k)
            s[m] = _ACGT[(_CODE[s[m]] + shift) % 4]
    if indel_rate > 0:
        n_ev = rng.binomial(len(s), indel_rate)
        if n_ev:
            pos = np.sort(rng.integers(1, max(2, len(s) - 1), size=n_ev))[::-1]
            out = s.tolist()
            for p in pos:
                ln = int(rng.integers(1, max_indel + 1))
                if rng.random() < 0.5:
                    del out[p:p + ln]
                else:
                    out[p:p] = _ACGT[rng.integers(0, 4, size=ln)].tolist()
            s = np.array(out, dtype=np.uint8)
    return s


def extension_pairs(n: int, read_len: int = 150, sub_rate: float = 0.01, indel_rate: float = 0.002,
                    max_indel: int = 3, rescue_frac: float = 0.05, n_rate: float = 0.0,
                    seed: int = 43, fixed_query_len: bool = True) -> PairBatch:
    """Pairs distributed like the reference's todo lists (SURVEY.md 8a workload facts):
    windows = read_len + diff + ext_left + ext_right with ext in [0, 50] (mostly 50), and a
    `rescue_frac` share of mate-rescue windows (read_len + ~200..250)."""
    rng = np.random.default_rng(seed)
    qs: List[np.ndarray] = []
    ts: List[np.ndarray] = []
    for _ in range(n):
        rescue = rng.random() < rescue_frac
        if rescue:
            left = int(rng.integers(0, 150))
            right = int(rng.integers(50, 100))
        else:
            left = 50 if rng.random() < 0.9 else int(rng.integers(0, 50))
            right = 50 if rng.random() < 0.9 else int(rng.integers(0, 50))
        core = _ACGT[rng.integers(0, 4, size=read_len)]
        read = _mutate(rng, core, sub_rate, indel_rate, max_indel)
        if fixed_query_len:
            # sequencers emit fixed-length reads: trim/extend to read_len
            if len(read) >= read_len:
                read = read[:read_len]
            else:
                read = np.concatenate([read, _ACGT[rng.integers(0, 4, size=read_len - len(read))]])
        win = np.concatenate([_ACGT[rng.integers(0, 4, size=left)], core,
                              _ACGT[rng.integers(0, 4, size=right)]])
        if n_rate > 0:
            m = rng.random(len(read)) < n_rate
            read = read.copy()
            read[m] = ord("N")
        qs.append(read)
        ts.append(win)
    qoff = np.zeros(n + 1, np.int64); qoff[1:] = np.cumsum([len(x) for x in qs])
    toff = np.zeros(n + 1, np.int64); toff[1:] = np.cumsum([len(x) for x in ts])
    return PairBatch(np.concatenate(qs), qoff, np.concatenate(ts), toff)


def fixed_pairs_fast(n: int, qlen: int = 150, tlen: int = 250, sub_rate: float = 0.01,
                     seed: int = 43) -> PairBatch:
    """Vectorised generator for the microbenchmark shape: fixed |q| x |t|,
    substitutions only, read placed at a random offset of the window.  Scales to millions of pairs."""
    rng = np.random.default_rng(seed)
    t = _ACGT[rng.integers(0, 4, size=(n, tlen), dtype=np.uint8)]
    start = rng.integers(0, tlen - qlen + 1, size=n)
    idx = start[:, None] + np.arange(qlen)[None, :]
    q = np.take_along_axis(t, idx, axis=1)
    if sub_rate > 0:
        m = rng.random((n, qlen)) < sub_rate
        shift = rng.integers(1, 4, size=(n, qlen), dtype=np.uint8)
        code = np.zeros(256, np.uint8)
        code[_ACGT] = np.arange(4, dtype=np.uint8)
        q = np.where(m, _ACGT[(code[q] + shift) % 4], q)
    qoff = np.arange(n + 1, dtype=np.int64) * qlen
    toff = np.arange(n + 1, dtype=np.int64) * tlen
    return PairBatch(np.ascontiguousarray(q).reshape(-1), qoff, np.ascontiguousarray(t).reshape(-1), toff)


def adversarial_pairs(n: int, seed: int = 7, max_q: int = 60, max_t: int = 90,
                      alphabet: bytes = b"ACGTN") -> PairBatch:
    """Short random pairs over small alphabets: dense in ties, zero scores, N cells and walks that
    leave the matrix (SURVEY.md 8a rules 3-9)."""
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    qs, ts = [], []
    for _ in range(n):
        ql = int(rng.integers(1, max_q + 1))
        tl = int(rng.integers(1, max_t + 1))
        if rng.random() < 0.5 and tl >= 4:
            t = alpha[rng.integers(0, len(alpha), size=tl)]
            a = int(rng.integers(0, tl))
            q = t[a:a + ql].copy()
            if len(q) == 0:
                q = alpha[rng.integers(0, len(alpha), size=1)]
            m = rng.random(len(q)) < 0.15
            q[m] = alpha[rng.integers(0, len(alpha), size=int(m.sum()))]
            if rng.random() < 0.5 and len(q) > 4:
                p = int(rng.integers(1, len(q) - 1))
                q = np.delete(q, slice(p, p + int(rng.integers(1, 4))))
        else:
            q = alpha[rng.integers(0, len(alpha), size=ql)]
            t = alpha[rng.integers(0, len(alpha), size=tl)]
        qs.append(q.astype(np.uint8))
        ts.append(t.astype(np.uint8))
    qoff = np.zeros(n + 1, np.int64); qoff[1:] = np.cumsum([len(x) for x in qs])
    toff = np.zeros(n + 1, np.int64); toff[1:] = np.cumsum([len(x) for x in ts])
    return PairBatch(np.concatenate(qs), qoff, np.concatenate(ts), toff)


def extension_pairs_fast(n: int, read_len: int = 150, sub_rate: float = 0.01, indel_rate: float = 0.002,
                         max_indel: int = 3, rescue_frac: float = 0.05, seed: int = 43,
                         chunk: int = 1 << 16) -> PairBatch:
    """Vectorised version of extension_pairs for bench-sized batches (BASELINE.json configs[1] shape):
    fixed-length reads; windows = 0..50 + read_len + 0..50 (90% of flanks are the full 50, as
    src/pc.cpp:231-239 clips them only at contig ends), `rescue_frac` mate-rescue windows with longer
    flanks; per read substitutions at `sub_rate` and at most one indel event (probability
    read_len*indel_rate, length 1..max_indel)."""
    rng = np.random.default_rng(seed)
    qparts, tparts, tlens = [], [], []
    maxw = read_len + 150 + 100 + max_indel
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        rescue = rng.random(m) < rescue_frac
        left = np.where(rng.random(m) < 0.9, 50, rng.integers(0, 50, size=m))
        right = np.where(rng.random(m) < 0.9, 50, rng.integers(0, 50, size=m))
        left = np.where(rescue, rng.integers(0, 150, size=m), left)
        right = np.where(rescue, rng.integers(50, 100, size=m), right)
        tlen = left + read_len + right
        T = _ACGT[rng.integers(0, 4, size=(m, maxw), dtype=np.uint8)]
        # read = window[left + j + shift(j)]
        ev = rng.random(m) < min(1.0, read_len * indel_rate)
        is_del = rng.random(m) < 0.5
        k = rng.integers(1, max_indel + 1, size=m)
        p = rng.integers(1, read_len - max_indel - 1, size=m)
        j = np.arange(read_len)[None, :]
        shift = np.zeros((m, read_len), dtype=np.int64)
        dmask = (ev & is_del)[:, None] & (j >= p[:, None])
        shift = np.where(dmask, k[:, None], shift)                      # deletion: skip k window bases
        imask = (ev & ~is_del)[:, None] & (j >= (p + k)[:, None])
        shift = np.where(imask, -k[:, None], shift)                     # insertion: k extra read bases
        idx = left[:, None] + j + shift
        Q = np.take_along_axis(T, idx, axis=1)
        ins = (ev & ~is_del)[:, None] & (j >= p[:, None]) & (j < (p + k)[:, None])
        Q = np.where(ins, _ACGT[rng.integers(0, 4, size=(m, read_len), dtype=np.uint8)], Q)
        if sub_rate > 0:
            sm = rng.random((m, read_len)) < sub_rate
            sh = rng.integers(1, 4, size=(m, read_len))
            Q = np.where(sm, _ACGT[(_CODE[Q] + sh) % 4], Q)
        keep = np.arange(maxw)[None, :] < tlen[:, None]
        tparts.append(T[keep])
        qparts.append(np.ascontiguousarray(Q, dtype=np.uint8).reshape(-1))
        tlens.append(tlen)
    tl = np.concatenate(tlens)
    qoff = np.arange(n + 1, dtype=np.int64) * read_len
    toff = np.zeros(n + 1, np.int64)
    toff[1:] = np.cumsum(tl)
    return PairBatch(np.concatenate(qparts), qoff, np.concatenate(tparts).astype(np.uint8), toff)


# ---- round-2 parity corners (shared by the GPU tests and the golden-vector generator) -------------------------------

def tie_dense_pairs(n: int, qmin: int, qmax: int, alphabet: bytes, seed: int, tmax: int = 700) -> PairBatch:
    """Related pairs over a tiny alphabet: equal maxima and co-optimal paths everywhere (first-maximum rule,
    H-source priority, gap-extension ties).  qmin/qmax select the lane geometry (257..500: 16-lane groups)."""
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    qs, ts = [], []
    for _ in range(n):
        ql = int(rng.integers(qmin, qmax + 1))
        tl = int(rng.integers(max(1, ql - 40), min(tmax, ql + 200)))
        t = alpha[rng.integers(0, len(alpha), size=tl)]
        a = int(rng.integers(0, max(1, tl - ql)))
        q = t[a:a + ql].copy()
        if len(q) < ql:
            q = np.concatenate([q, alpha[rng.integers(0, len(alpha), size=ql - len(q))]])
        m = rng.random(ql) < 0.05
        q[m] = alpha[rng.integers(0, len(alpha), size=int(m.sum()))]
        if rng.random() < 0.5 and ql > 12:
            p = int(rng.integers(5, ql - 5))
            q = np.concatenate([q[:p], q[p + 2:], alpha[rng.integers(0, len(alpha), size=2)]])
        qs.append(q.astype(np.uint8).tobytes())
        ts.append(t.astype(np.uint8).tobytes())
    return from_lists(qs, ts)


def saturation_pairs(seed: int = 220, qlens=(496, 500, 511, 512)) -> PairBatch:
    """Perfect / near-perfect matches of long queries: scores 2*|q| - small, i.e. 992..1024 at the default scoring --
    the top of the packed kernel's key range (score << 5 | column in a positive s16: score <= 1023) and the hand-off to
    the exact kernel at 1024.  Six variants per length: perfect, one mismatch, one deletion, homopolymer, AC repeat,
    leading N."""
    rng = np.random.default_rng(seed)
    qs, ts = [], []
    for ql in qlens:
        for variant in range(6):
            t = _ACGT[rng.integers(0, 4, size=ql + 60)]
            q = t[30:30 + ql].copy()
            if variant == 1:
                q[ql // 2] = _ACGT[(_CODE[q[ql // 2]] + 1) % 4]
            if variant == 2:
                q = np.concatenate([q[:100], q[101:], _ACGT[rng.integers(0, 4, size=1)]])
            if variant == 3:
                t = np.full(ql + 60, ord("A"), np.uint8)
                q = t[:ql].copy()
            if variant == 4:
                t = np.tile(np.frombuffer(b"AC", dtype=np.uint8), (ql + 60) // 2)
                q = t[:ql].copy()
            if variant == 5:
                q[0] = ord("N")
            qs.append(q.astype(np.uint8).tobytes())
            ts.append(t.astype(np.uint8).tobytes())
    return from_lists(qs, ts)


def window_edge_pairs(seed: int = 221, tlens=(2046, 2047, 2048, 2049), qlens=(150, 300)) -> PairBatch:
    """Windows at the packed kernel's row limit (2047) and just beyond; the alignment ends in the last rows."""
    rng = np.random.default_rng(seed)
    qs, ts = [], []
    for tl in tlens:
        for ql in qlens:
            t = _ACGT[rng.integers(0, 4, size=tl)]
            a = tl - ql - 3
            q = t[a:a + ql].copy()
            q[ql // 3] = ord("A") if q[ql // 3] != ord("A") else ord("C")
            qs.append(q.tobytes())
            ts.append(t.tobytes())
    return from_lists(qs, ts)


# ---- seeding workloads: synthetic genome with injected repeats + simulated reads -------------------

_COMP = np.zeros(256, np.uint8)
_COMP[_ACGT] = np.frombuffer(b"TGCA", dtype=np.uint8)
_COMP[ord("N")] = ord("N")


def seeding_genome(n_contigs: int = 3, contig_len: int = 1_000_000, seed: int = 5, repeat_families: int = 4,
                   copies_per_contig: int = 10, family_len: int = 2000, divergence: float = 0.02,
                   low_complexity: int = 20) -> List[np.ndarray]:
    """Random contigs with (a) repeat families copied into every contig at `divergence` (BASELINE configs[2]: "random +
    injected repeats"), so reads from them hit many loci on several reference sequences, and (b) a few low-complexity
    stretches (homopolymers, dinucleotide repeats) where equal s-mer hashes exercise the syncmer tie rules."""
    rng = np.random.default_rng(seed)
    contigs = [_ACGT[rng.integers(0, 4, size=contig_len)] for _ in range(n_contigs)]
    fams = [_ACGT[rng.integers(0, 4, size=family_len)] for _ in range(repeat_families)]
    for c in contigs:
        for fam in fams:
            for _ in range(copies_per_contig):
                if len(c) <= family_len + 10:
                    continue
                p = int(rng.integers(0, len(c) - family_len - 1))
                seg = fam.copy()
                m = rng.random(family_len) < divergence
                seg[m] = _ACGT[rng.integers(0, 4, size=int(m.sum()))]
                c[p:p + family_len] = seg
        for _ in range(low_complexity):
            ln = int(rng.integers(30, 120))
            if len(c) <= ln + 10:
                continue
            p = int(rng.integers(0, len(c) - ln - 1))
            unit = _ACGT[rng.integers(0, 4, size=int(rng.integers(1, 4)))]
            c[p:p + ln] = np.resize(unit, ln)
    return contigs


def seeding_reads(contigs: List[np.ndarray], n: int, read_len: int = 150, seed: int = 6, sub_rate: float = 0.01,
                  indel_rate: float = 0.002, n_rate: float = 0.0, junk_frac: float = 0.01,
                  vary_len: bool = False):
    """Reads sampled from the contigs (both strands), mutated; `junk_frac` of them are random (no locus), some are
    shorter than a seed.  Returns (buf uint8, offsets int64)."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        ln = read_len if not vary_len else int(rng.integers(max(8, read_len // 3), read_len + 1))
        if rng.random() < junk_frac:
            s = _ACGT[rng.integers(0, 4, size=ln)] if rng.random() < 0.7 else _ACGT[rng.integers(0, 4, size=int(rng.integers(1, 30)))]
        else:
            c = contigs[int(rng.integers(0, len(contigs)))]
            ln = min(ln, len(c))
            p = int(rng.integers(0, len(c) - ln + 1))
            s = _mutate(rng, c[p:p + ln], sub_rate, indel_rate, 3)
            if rng.random() < 0.5:
                s = _COMP[s][::-1]
        if n_rate > 0 and len(s):
            m = rng.random(len(s)) < n_rate
            s = s.copy()
            s[m] = ord("N")
        out.append(np.ascontiguousarray(s, dtype=np.uint8))
    off = np.zeros(n + 1, np.int64)
    off[1:] = np.cumsum([len(x) for x in out])
    return (np.concatenate(out) if out else np.zeros(0, np.uint8)), off
