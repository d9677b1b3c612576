"""Inputs for the Hamming-shortcut tests (reference src/aln.cpp:391-404): (read, equally long window) pairs around the
5 % decision boundary, with mismatches clustered at the ends (soft clips, end-bonus ties), N bases, lower case, empty
and unequal-length pairs."""
import numpy as np

from oracle import pack_strings

ALPHA = np.frombuffer(b"ACGT", np.uint8)


def make_pairs(n: int, seed: int, read_len: int = 150, var_len: bool = True):
    rng = np.random.default_rng(seed)
    qs, ts = [], []
    for i in range(n):
        L = int(rng.integers(1, 501)) if (var_len and i % 3 == 0) else read_len
        t = ALPHA[rng.integers(0, 4, L)].copy()
        q = t.copy()
        kind = i % 8
        # number of mismatches around 5 % of the length (both sides of the float boundary), or far away
        base = L * 0.05
        k = int(max(0, min(L, round(base + rng.integers(-3, 4))))) if kind < 5 else int(rng.integers(0, max(1, L // 3)))
        if kind == 1:      # clustered at the start: soft clip / end-bonus decision
            pos = rng.choice(min(L, max(k, 12)), size=min(k, min(L, max(k, 12))), replace=False)
        elif kind == 2:    # clustered at the end
            w = min(L, max(k, 12))
            pos = L - 1 - rng.choice(w, size=min(k, w), replace=False)
        else:
            pos = rng.choice(L, size=k, replace=False)
        for p in pos:
            q[p] = ALPHA[(int(np.searchsorted(ALPHA, q[p])) + 1 + int(rng.integers(0, 3))) % 4] if q[p] in ALPHA else ord("A")
        if kind == 3 and L > 4:   # N bases compare as ordinary characters here (std::string ==)
            q[int(rng.integers(0, L))] = ord("N")
            t[int(rng.integers(0, L))] = ord("N")
        if kind == 4 and L > 4:   # lower case differs from upper case
            p = int(rng.integers(0, L))
            q[p] = q[p] | 0x20
        qs.append(q.tobytes())
        ts.append(t.tobytes())
    # corner cases: empty read, unequal lengths, all mismatches, single base
    qs += [b"", b"ACGT", b"AAAAAAAAAA", b"A", b"C"]
    ts += [b"", b"ACGTA", b"CCCCCCCCCC", b"A", b"A"]
    qbuf, qoff = pack_strings(qs)
    tbuf, toff = pack_strings(ts)
    return qbuf, qoff, tbuf, toff


def record_tuple(aln, i, cigar_string):
    return (int(aln["status"][i]), int(aln["sw_score"][i]), int(aln["edit_distance"][i]), int(aln["query_start"][i]),
            int(aln["query_end"][i]), int(aln["ref_start"][i]), int(aln["ref_end"][i]), cigar_string(aln[i]))
