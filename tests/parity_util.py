"""Shared comparison helpers: CUDA engine records vs oracle arrays, bit-exact."""
import numpy as np


def oracle_arrays(olib, b, **score_kw):
    score, qs, qe, rs, re, nops, pool, coff = olib.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff, **score_kw)
    raw = pool.tobytes()
    cig = [raw[coff[i]:coff[i + 1]].decode() for i in range(b.n)]
    return dict(score=score, query_start=qs, query_end=qe, ref_start=rs, ref_end=re, n_ops=nops, cigar=cig)


def compare(engine, res, ora, b, max_report=5):
    """Returns a list of human-readable mismatch lines (empty = parity)."""
    bad = []
    n = len(res)
    fields = ["score", "query_start", "query_end", "ref_start", "ref_end", "n_ops"]
    mism = np.zeros(n, bool)
    for f in fields:
        mism |= res[f].astype(np.int64) != ora[f].astype(np.int64)
    mism |= res["status"] != 0
    idx = set(np.nonzero(mism)[0].tolist())
    # CIGAR text for every pair (cheap enough at test sizes)
    for i in range(n):
        if i in idx:
            continue
        if engine.cigar(res, i) != ora["cigar"][i]:
            idx.add(i)
    qs_, ts_ = None, None
    for i in sorted(idx)[:max_report]:
        if qs_ is None:
            qs_, ts_ = b.queries(), b.targets()
        got = tuple(int(res[f][i]) for f in fields) + (int(res["status"][i]),)
        try:
            gc = engine.cigar(res, i)
        except Exception as ex:  # noqa: BLE001
            gc = f"<{ex}>"
        exp = tuple(int(ora[f][i]) for f in fields)
        bad.append(f"pair {i}: got {got} {gc} expected {exp} {ora['cigar'][i]} q={qs_[i].decode(errors='replace')} "
                   f"t={ts_[i].decode(errors='replace')}")
    if idx:
        bad.insert(0, f"{len(idx)} of {n} pairs differ")
    return bad
