"""Per-chunk DP / traceback timeline of a resident run (overlapped and serialized engines), via RSA_EXT_TRACE."""
import os, sys
sys.path.insert(0, "/root/repo")
from rabbitsalign_b200 import ExtensionEngine, workload as W
b = W.extension_pairs_fast(1048576, seed=5)
for flags in (False, True):
    eng = ExtensionEngine(serialize=flags)
    eng.stage_resident(b.qbuf, b.qoff, b.tbuf, b.toff)
    for _ in range(3):
        eng.run_resident()
    os.environ["RSA_EXT_TRACE"] = "1"
    print("serialize", flags, file=sys.stderr)
    st = eng.stats()
    os.environ.pop("RSA_EXT_TRACE")
    eng.close()
