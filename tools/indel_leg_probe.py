"""Per-kernel view of bench.py's leg_250bp_5pct_indel (BASELINE configs[3]): routing, DP / traceback device time of the
serialised engine, and (under `ncu --metrics gpu__time_duration.sum`) the duration of every class launch."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rabbitsalign_b200 import ExtensionEngine, workload as W  # noqa: E402

u = W.extension_pairs(8192, seed=44, read_len=250, indel_rate=0.05, max_indel=4, fixed_query_len=False)
reps = 16
qbuf = np.tile(u.qbuf, reps)
tbuf = np.tile(u.tbuf, reps)
qoff = np.concatenate([u.qoff[:-1] + k * int(u.qoff[-1]) for k in range(reps)] + [np.array([reps * int(u.qoff[-1])])]).astype(np.int64)
toff = np.concatenate([u.toff[:-1] + k * int(u.toff[-1]) for k in range(reps)] + [np.array([reps * int(u.toff[-1])])]).astype(np.int64)
cells = float(u.cells) * reps
ql = np.diff(u.qoff)
out = {"qlen_hist": {int(k): int(v) for k, v in zip(*np.unique((ql + 7) // 8, return_counts=True))}, "cells": cells}
for serial in (True, False):
    eng = ExtensionEngine(serialize=serial)
    eng.stage_resident(qbuf, qoff, tbuf, toff)
    for _ in range(4):
        eng.run_resident()
    st = eng.stats()
    out["serial" if serial else "pipelined"] = {k: st[k] for k in ("dp_ms", "tb_ms", "kernel_launches", "pairs_fast", "pairs_exact", "pairs_redo")}
    if serial:
        out["dp_gcups_serial"] = cells / (st["dp_ms"] * 1e-3) / 1e9
    eng.close()
print(json.dumps(out))
