"""Where the 0.44 ms of one blocking 512-pair call goes: device timeline of the kernels (resident run, CUDA events)
next to the wall time of the whole call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rabbitsalign_b200 import ExtensionEngine, workload as W
for n in (512, 4096):
    b = W.extension_pairs_fast(n, seed=5)
    eng = ExtensionEngine(serialize=True)
    eng.stage_resident(b.qbuf, b.qoff, b.tbuf, b.toff)
    for _ in range(5):
        eng.run_resident()
    os.environ["RSA_EXT_TRACE"] = "1"
    st = eng.stats()
    os.environ.pop("RSA_EXT_TRACE")
    for _ in range(20):
        eng.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    t0 = time.perf_counter()
    for _ in range(200):
        eng.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
    dt = (time.perf_counter() - t0) / 200
    print(f"n={n}: whole call {dt*1e6:.0f} us; kernels dp {st['dp_ms']*1e3:.0f} us, traceback {st['tb_ms']*1e3:.0f} us", file=sys.stderr)
    eng.close()
