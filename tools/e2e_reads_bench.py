"""End-to-end reads/s of the reference's host pipeline (FASTQ in -> SAM out, wall clock) with the extension on
(a) the B200 engine (integration/_build/rabbitsalign_b200) and (b) the reference's CPU SSW path
(integration/_build/rabbitsalign_cpussw), same inputs, same thread count.  BASELINE.json metric (i).

    python tools/e2e_reads_bench.py [--ref-len 20000000] [--reads 400000] [--threads N] [--paired]

Prints one JSON line.  The pipeline around the boundary (seeding, NAMs, SAM) is the reference's unmodified host
code, so this number is bounded by the host; it is reported, not optimised, in this round.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = os.path.join(ROOT, "integration", "_build")


def md5_nopg(path):
    h = hashlib.md5()
    for line in open(path, "rb"):
        if not line.startswith(b"@PG"):
            h.update(line)
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-len", type=int, default=20_000_000)
    ap.add_argument("--contigs", type=int, default=4)
    ap.add_argument("--reads", type=int, default=400_000)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--paired", action="store_true")
    ap.add_argument("--batch", type=int, default=0, help="unused (STREAM_BATCH_SIZE is a compile-time macro of the veneer)")
    a = ap.parse_args()
    out = {"ref_len": a.ref_len, "reads": a.reads * (2 if a.paired else 1), "paired": a.paired, "threads": a.threads}
    with tempfile.TemporaryDirectory() as d:
        t0 = time.time()
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_reads.py"), d, "--ref-len", str(a.ref_len),
                               "--contigs", str(a.contigs), "--reads", str(a.reads), "--seed", "77"] + (["--paired"] if a.paired else []))
        out["gen_s"] = round(time.time() - t0, 1)
        files = [os.path.join(d, "ref.fa"), os.path.join(d, "reads_1.fq")] + ([os.path.join(d, "reads_2.fq")] if a.paired else [])
        for name in ("rabbitsalign_cpussw", "rabbitsalign_b200"):
            exe = os.path.join(B, name)
            if not os.path.exists(exe):
                out[name] = "not built"
                continue
            sam = os.path.join(d, name + ".sam")
            t0 = time.time()
            r = subprocess.run([exe, "-t", str(a.threads), "-o", sam] + files, capture_output=True, text=True)
            dt = time.time() - t0
            if r.returncode != 0:
                out[name] = {"error": r.stderr[-500:]}
                continue
            mapping = [ln for ln in r.stderr.splitlines() if "Total time" in ln or "indexing" in ln.lower()]
            out[name] = {"wall_s": round(dt, 2), "reads_per_s_wall": round(out["reads"] / dt), "sam_md5": md5_nopg(sam),
                         "stderr_times": mapping[-8:]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
