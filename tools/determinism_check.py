"""Stress: the same pairs through different call shapes / threads / scratch histories must give identical records.

    python tools/determinism_check.py [pairs] [rounds] [threads]

Baseline = one handle, whole batch.  Then `rounds` passes of 512-pair slices in shuffled order on `threads` handles
concurrently (the reference pipeline's call pattern).  Any record that differs from the baseline is printed with the
oracle's answer for that pair."""
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from rabbitsalign_b200 import ExtensionEngine, workload as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 6
T = int(sys.argv[3]) if len(sys.argv) > 3 else 16
b = W.extension_pairs(n, seed=900, read_len=250, indel_rate=0.02, max_indel=4, sub_rate=0.02, fixed_query_len=False)
FIELDS = ["score", "query_start", "query_end", "ref_start", "ref_end", "n_ops", "status"]
base_eng = ExtensionEngine()
base = base_eng.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff).copy()
print("baseline: status!=0:", int((base["status"] != 0).sum()), "n_ops>40:", int((base["n_ops"] > 40).sum()), flush=True)
slices = [(lo, min(b.n, lo + 512)) for lo in range(0, b.n, 512)]
parts = [b.slice(lo, hi) for lo, hi in slices]
engines = [ExtensionEngine() for _ in range(T)]
bad = {}
lock = threading.Lock()


errors = []


def worker(w, order):
    e = engines[w]
    for k in order:
        p = parts[k]
        try:
            r = e.align_packed(p.qbuf, p.qoff, p.tbuf, p.toff)
        except Exception as ex:  # noqa: BLE001
            with lock:
                errors.append((w, k, str(ex)))
            continue
        lo, hi = slices[k]
        ref = base[lo:hi]
        diff = np.zeros(hi - lo, bool)
        for f in FIELDS:
            diff |= r[f] != ref[f]
        short = ref["n_ops"] <= 40
        diff |= short & (r["rle"] != ref["rle"]).any(axis=1)
        for i in np.nonzero(diff)[0]:
            with lock:
                bad.setdefault(lo + int(i), []).append((w, {f: int(r[f][i]) for f in FIELDS}, e.cigar(r, int(i))))


rng = np.random.default_rng(5)
for rd in range(rounds):
    perm = rng.permutation(len(parts))
    th = [threading.Thread(target=worker, args=(w, perm[w::T])) for w in range(T)]
    [t.start() for t in th]
    [t.join() for t in th]
    print(f"round {rd}: {len(bad)} pairs differ so far", flush=True)
if bad:
    olib = oracle.restatement()
    for i in sorted(bad)[:10]:
        sub = b.slice(i, i + 1)
        o = olib.align_packed(sub.qbuf, sub.qoff, sub.tbuf, sub.toff)
        raw = o[6].tobytes()
        print("pair", i, "qlen", int(b.qoff[i + 1] - b.qoff[i]), "tlen", int(b.toff[i + 1] - b.toff[i]))
        print("   baseline", {f: int(base[f][i]) for f in FIELDS}, base_eng.cigar(base, i))
        print("   oracle  ", [int(x[0]) for x in o[:6]], raw[o[7][0]:o[7][1]].decode())
        for w, rec, cg in bad[i][:3]:
            print("   sliced w%d" % w, rec, cg)
for w, k, msg in errors[:5]:
    print("call error: worker", w, "slice", k, msg)
print("RESULT", "deterministic" if not bad and not errors else f"{len(bad)} nondeterministic pairs, {len(errors)} failed calls")
sys.exit(1 if bad or errors else 0)
