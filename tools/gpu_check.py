"""First-contact GPU diagnostic: environment facts + staged parity (exact kernel, then packed kernel),
reporting every stage instead of stopping at the first failure.  Run under gpurun."""
import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=60).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return f"<{e}>"


def main():
    print("== env ==")
    print(sh("nvidia-smi -L"))
    print("nproc", sh("nproc"), "|", sh("lscpu | grep 'Model name'"))
    print("reference present:", os.path.isdir("/root/reference"), "| ncu:", sh("which ncu"))
    import numpy as np
    import oracle
    from rabbitsalign_b200 import ExtensionEngine, workload as W
    from parity_util import compare, oracle_arrays
    olib = oracle.restatement()
    cases = [
        ("tiny", lambda: W.from_lists([b"ACGTNCGTAC", b"ACGTACGTAC", b"AAAA", b"NNNN", b"ACGT", b"ACGTACGTACGT"],
                                      [b"ACGTACGTAC", b"ACGTACGTAC", b"CCCC", b"ACGT", b"TTACGTTT", b"GGACGTACGTACGTCC"])),
        ("adv_acgtn", lambda: W.adversarial_pairs(3000, seed=101)),
        ("adv_ac", lambda: W.adversarial_pairs(3000, seed=102, alphabet=b"AC")),
        ("adv_mid", lambda: W.adversarial_pairs(1500, seed=104, max_q=200, max_t=400)),
        ("ext150", lambda: W.extension_pairs(2000, seed=105)),
        ("ext150_var", lambda: W.extension_pairs(1500, seed=106, fixed_query_len=False, indel_rate=0.01)),
        ("ext250_indel", lambda: W.extension_pairs(800, seed=107, read_len=250, indel_rate=0.05, max_indel=4, fixed_query_len=False)),
        ("ext150_N", lambda: W.extension_pairs(1500, seed=108, n_rate=0.01)),
        ("iupac", lambda: W.adversarial_pairs(2000, seed=103, alphabet=b"ACGTNacgtnRYKMSW.-")),
    ]
    for mode in ("exact", "packed"):
        print(f"== {mode} ==")
        try:
            eng = ExtensionEngine(device=0, exact_only=(mode == "exact"))
        except Exception:
            traceback.print_exc()
            continue
        for name, mk in cases:
            try:
                b = mk()
                t0 = time.time()
                res = eng.align_packed(b.qbuf, b.qoff, b.tbuf, b.toff)
                dt = time.time() - t0
                bad = compare(eng, res, oracle_arrays(olib, b), b, max_report=4)
                st = eng.stats()
                print(f"{name}: n={b.n} {'OK' if not bad else 'MISMATCH'} {dt*1e3:.1f} ms fast={st['pairs_fast']} "
                      f"exact={st['pairs_exact']} launches={st['kernel_launches']}")
                for line in bad:
                    print("   ", line[:600])
            except Exception:
                traceback.print_exc()
        eng.close()


if __name__ == "__main__":
    main()
