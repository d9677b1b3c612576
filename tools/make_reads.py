"""Deterministic synthetic reference + reads for the end-to-end SAM checks (SURVEY.md 8d cfg1/cfg4 shapes).

    python tools/make_reads.py OUTDIR --ref-len 5386 --reads 10000 --read-len 150 [--paired] [--indel 0.004]

Writes OUTDIR/ref.fa and OUTDIR/reads_1.fq (and reads_2.fq).  Pure numpy with fixed seeds so that the dev
container (golden generation) and the GPU box (product run) produce byte-identical inputs.
"""
import argparse
import os

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.zeros(256, np.uint8)
COMP[ACGT] = np.frombuffer(b"TGCA", dtype=np.uint8)
COMP[ord("N")] = ord("N")
CODE = np.zeros(256, np.int64)
CODE[ACGT] = np.arange(4)


def mutate(rng, s, sub, indel, max_indel, n_rate):
    s = s.copy()
    m = rng.random(len(s)) < sub
    if m.any():
        s[m] = ACGT[(CODE[s[m]] + rng.integers(1, 4, size=int(m.sum()))) % 4]
    n_ev = rng.binomial(len(s), indel)
    if n_ev:
        out = s.tolist()
        for p in np.sort(rng.integers(1, len(s) - 1, size=n_ev))[::-1]:
            ln = int(rng.integers(1, max_indel + 1))
            if rng.random() < 0.5:
                del out[p:p + ln]
            else:
                out[p:p] = ACGT[rng.integers(0, 4, size=ln)].tolist()
        s = np.array(out, dtype=np.uint8)
    if n_rate > 0:
        m = rng.random(len(s)) < n_rate
        s[m] = ord("N")
    return s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("outdir")
    ap.add_argument("--ref-len", type=int, default=5386)
    ap.add_argument("--contigs", type=int, default=1)
    ap.add_argument("--reads", type=int, default=10000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--paired", action="store_true")
    ap.add_argument("--sub", type=float, default=0.01)
    ap.add_argument("--indel", type=float, default=0.003)
    ap.add_argument("--max-indel", type=int, default=3)
    ap.add_argument("--n-rate", type=float, default=0.0)
    ap.add_argument("--insert-mean", type=float, default=300.0)
    ap.add_argument("--insert-sd", type=float, default=30.0)
    ap.add_argument("--seed", type=int, default=42)
    a = ap.parse_args()
    os.makedirs(a.outdir, exist_ok=True)
    rng = np.random.default_rng(a.seed)
    contigs = [ACGT[rng.integers(0, 4, size=a.ref_len // a.contigs)] for _ in range(a.contigs)]
    with open(os.path.join(a.outdir, "ref.fa"), "wb") as f:
        for i, c in enumerate(contigs):
            f.write(b">contig%d\n" % i)
            b = c.tobytes()
            for k in range(0, len(b), 60):
                f.write(b[k:k + 60] + b"\n")
    rng = np.random.default_rng(a.seed + 1)
    f1 = open(os.path.join(a.outdir, "reads_1.fq"), "wb")
    f2 = open(os.path.join(a.outdir, "reads_2.fq"), "wb") if a.paired else None
    for i in range(a.reads):
        c = contigs[int(rng.integers(0, len(contigs)))]
        if a.paired:
            ins = int(max(a.read_len + 10, rng.normal(a.insert_mean, a.insert_sd)))
            ins = min(ins, len(c) - 1)
            p = int(rng.integers(0, len(c) - ins))
            frag = c[p:p + ins]
            if rng.random() < 0.5:
                frag = COMP[frag][::-1]
            r1 = mutate(rng, frag[:a.read_len], a.sub, a.indel, a.max_indel, a.n_rate)
            r2 = mutate(rng, COMP[frag[-a.read_len:]][::-1], a.sub, a.indel, a.max_indel, a.n_rate)
            f1.write(b"@r%d/1\n" % i + r1.tobytes() + b"\n+\n" + b"I" * len(r1) + b"\n")
            f2.write(b"@r%d/2\n" % i + r2.tobytes() + b"\n+\n" + b"I" * len(r2) + b"\n")
        else:
            p = int(rng.integers(0, len(c) - a.read_len))
            s = c[p:p + a.read_len]
            if rng.random() < 0.5:
                s = COMP[s][::-1]
            r1 = mutate(rng, s, a.sub, a.indel, a.max_indel, a.n_rate)
            f1.write(b"@r%d\n" % i + r1.tobytes() + b"\n+\n" + b"I" * len(r1) + b"\n")
    f1.close()
    if f2:
        f2.close()


if __name__ == "__main__":
    main()
