"""Host-side timeline of one end-to-end submit/wait of the bench batch from pinned host buffers
(RSA_EXT_TRACE laps: plan, enqueue, retire)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rabbitsalign_b200 import ExtensionEngine, workload as W
from rabbitsalign_b200.ext import RESULT_DTYPE
b = W.extension_pairs_fast(1048576, seed=5)
def pinned(a):
    t = torch.empty(a.nbytes, dtype=torch.uint8).pin_memory()
    v = t.numpy().view(a.dtype).reshape(a.shape)
    v[...] = a
    return t, v
keep = [pinned(x) for x in (b.qbuf, b.qoff, b.tbuf, b.toff)]
tres = torch.empty(b.n * RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
res = tres.numpy().view(RESULT_DTYPE)
eng = ExtensionEngine()
def step():
    eng.submit(keep[0][1], keep[1][1], keep[2][1], keep[3][1], res)
    eng.wait()
for _ in range(3):
    step()
os.environ["RSA_EXT_TRACE"] = "1"
t0 = time.perf_counter()
step()
print("step ms", (time.perf_counter() - t0) * 1e3, file=sys.stderr)
os.environ.pop("RSA_EXT_TRACE")
eng.close()
