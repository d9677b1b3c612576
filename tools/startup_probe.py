"""Where does a worker's first solve_ssw_on_gpu call spend its time?  T threads, each: create a handle, then
three 512-pair calls (the reference's STREAM_BATCH_SIZE slice); prints per-phase wall times (max over threads).
python tools/startup_probe.py [threads]"""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rabbitsalign_b200 import ExtensionEngine, workload as W  # noqa: E402
from rabbitsalign_b200 import ext  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 16
batches = [W.extension_pairs_fast(512, seed=100 + i) for i in range(T)]
t0 = time.perf_counter()
ndev = ext.load_library().rsa_ext_device_count()
t_init = time.perf_counter() - t0
rows = [None] * T
bar = threading.Barrier(T)


def worker(i):
    b = batches[i]
    bar.wait()
    ts = [time.perf_counter()]
    eng = ExtensionEngine()
    ts.append(time.perf_counter())
    for _ in range(4):
        eng.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
        ts.append(time.perf_counter())
    rows[i] = [ts[k + 1] - ts[k] for k in range(len(ts) - 1)]
    bar.wait()
    eng.close()


t1 = time.perf_counter()
th = [threading.Thread(target=worker, args=(i,)) for i in range(T)]
[t.start() for t in th]
[t.join() for t in th]
wall = time.perf_counter() - t1
names = ["create", "call1", "call2", "call3", "call4"]
out = {"threads": T, "device_count_s": round(t_init, 3), "wall_s": round(wall, 3)}
for k, nme in enumerate(names):
    out[nme + "_max_ms"] = round(1e3 * max(r[k] for r in rows), 2)
    out[nme + "_mean_ms"] = round(1e3 * sum(r[k] for r in rows) / T, 2)
print(json.dumps(out))
