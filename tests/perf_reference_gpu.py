"""Not a test (not collected): the reference's GPU path timed on this box, next to the product, same pairs.

    python tests/perf_reference_gpu.py [--pairs 131072] [--workers 16]

The reference path (oracle/_ref/libgasal_gpu.so = GASAL2 + src/gasal2_ssw.cpp for sm_100a) is driven as
src/pc.cpp drives it: blocking 512-pair slices, one GASAL stream per worker thread.  Prints one JSON line with the
extension GCUPS of (a) the reference GPU path with 1 and W worker threads, (b) the product through the same call shape
(512-pair slices per worker, C ABI), (c) the product with the whole batch in one call."""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from rabbitsalign_b200 import ExtensionEngine, workload as W  # noqa: E402


def run_threads(fn, workers):
    th = [threading.Thread(target=fn, args=(w,)) for w in range(workers)]
    t0 = time.perf_counter()
    [t.start() for t in th]
    [t.join() for t in th]
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=131072)
    ap.add_argument("--workers", type=int, default=16)
    a = ap.parse_args()
    ref = oracle.reference_gpu()
    if ref is None:
        print(json.dumps({"unavailable": "oracle/_ref/libgasal_gpu.so not built"}))
        return
    b = W.extension_pairs_fast(a.pairs, seed=43)
    out = {"pairs": b.n, "cells": b.cells, "slice": ref.slice_size, "workers": a.workers,
           "workload": "bench.py batch shape (150 bp reads, windows read+flanks / mate-rescue)"}
    shards = [b.slice(w * b.n // a.workers, (w + 1) * b.n // a.workers) for w in range(a.workers)]

    # (a) reference GPU path
    ref.batch(*_arr(b.slice(0, 1024)), thread_id=0)  # context, GASAL storage
    t0 = time.perf_counter()
    ref.batch(*_arr(b), thread_id=0)
    dt1 = time.perf_counter() - t0
    for w in range(a.workers):
        ref.batch(*_arr(shards[w].slice(0, 512)), thread_id=w)  # per-worker storage
    dtw = run_threads(lambda w: ref.batch(*_arr(shards[w]), thread_id=w), a.workers)
    out["reference_gpu"] = {"gcups_1_worker": b.cells / dt1 / 1e9, "gcups_%d_workers" % a.workers: b.cells / dtw / 1e9,
                            "s_1_worker": dt1, "s_workers": dtw}

    # (b) product, same call shape
    engines = [ExtensionEngine() for _ in range(a.workers)]

    def sliced(w):
        e, s = engines[w], shards[w]
        for lo in range(0, s.n, 512):
            p = s.slice(lo, min(s.n, lo + 512))
            e.align_packed(p.qbuf, p.qoff, p.tbuf, p.toff)
    parts = [[shards[w].slice(lo, min(shards[w].n, lo + 512)) for lo in range(0, shards[w].n, 512)] for w in range(a.workers)]

    def sliced_pre(w):
        e = engines[w]
        for p in parts[w]:
            e.align_packed(p.qbuf, p.qoff, p.tbuf, p.toff)
    sliced_pre(0)
    t0 = time.perf_counter()
    for w in range(a.workers):
        sliced_pre_one = parts[w]
        for p in sliced_pre_one:
            engines[0].align_packed(p.qbuf, p.qoff, p.tbuf, p.toff)
    d1 = time.perf_counter() - t0
    run_threads(sliced_pre, a.workers)
    dW = run_threads(sliced_pre, a.workers)
    out["product_sliced"] = {"gcups_1_worker": b.cells / d1 / 1e9, "gcups_%d_workers" % a.workers: b.cells / dW / 1e9}

    # (c) product, one call
    engines[0].align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    t0 = time.perf_counter()
    engines[0].align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    d = time.perf_counter() - t0
    out["product_one_call"] = {"gcups": b.cells / d / 1e9, "note": "pageable numpy buffers (bench.py e2e uses pinned)"}
    for e in engines:
        e.close()
    print(json.dumps(out))


def _arr(b):
    return b.qbuf, b.qoff, b.tbuf, b.toff


if __name__ == "__main__":
    main()
