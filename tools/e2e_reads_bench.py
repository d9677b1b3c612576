"""End-to-end reads/s of the reference's host pipeline (FASTQ in -> SAM out, wall clock) with the extension on
(a) the B200 engine (integration/_build/rabbitsalign_b200) and (b) the reference's CPU SSW path
(integration/_build/rabbitsalign_cpussw), same inputs, same thread count.  BASELINE.json metric (i).

    python tools/e2e_reads_bench.py [--ref-len 20000000] [--reads 400000] [--threads N] [--paired] [--subsets a,b]

Prints one JSON line.  With --subsets (read counts <= --reads, single-end) every binary also runs on the first a, b, ...
reads of the same file and the line carries "mapping_s_per_mreads": the slope of the pipeline's own "Total time
mapping" between the smallest and the largest run, i.e. the steady-state cost with process start-up (CUDA context,
index load) taken out.  The pipeline around the boundary (seeding, NAMs, SAM) is the reference's unmodified host
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
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--sub", type=float, default=0.01)
    ap.add_argument("--indel", type=float, default=0.003)
    ap.add_argument("--max-indel", type=int, default=3)
    ap.add_argument("--binaries", default="", help="comma-separated binaries under integration/_build (default: all)")
    ap.add_argument("--repeat", type=int, default=1, help="runs per binary; the best wall time is reported, all are listed")
    ap.add_argument("--subsets", default="", help="comma-separated read counts to also run (prefixes of the read file)")
    ap.add_argument("--workdir", default="", help="where inputs and SAM files live (default: /dev/shm when it has room -- a 10 M-read "
                                                  "run writes 3.5 GB of SAM per binary and a slow scratch disk then times the disk -- else the "
                                                  "system temp directory)")
    a = ap.parse_args()
    out = {"ref_len": a.ref_len, "reads": a.reads * (2 if a.paired else 1), "paired": a.paired, "threads": a.threads,
           "read_len": a.read_len, "sub": a.sub, "indel": a.indel, "max_indel": a.max_indel}
    workdir = a.workdir or None
    if workdir is None and os.path.isdir("/dev/shm"):
        st = os.statvfs("/dev/shm")
        need = a.reads * (2 if a.paired else 1) * (a.read_len * 5 + 400) + 2 * a.ref_len   # FASTQ + one SAM + FASTA, generous
        if st.f_bavail * st.f_frsize > 2 * need:
            workdir = "/dev/shm"
    out["workdir"] = workdir or tempfile.gettempdir()
    with tempfile.TemporaryDirectory(dir=workdir) as d:
        t0 = time.time()
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_reads.py"), d, "--ref-len", str(a.ref_len),
                               "--contigs", str(a.contigs), "--reads", str(a.reads), "--seed", "77", "--read-len", str(a.read_len),
                               "--sub", str(a.sub), "--indel", str(a.indel), "--max-indel", str(a.max_indel)] + (["--paired"] if a.paired else []))
        out["gen_s"] = round(time.time() - t0, 1)
        files = [os.path.join(d, "ref.fa"), os.path.join(d, "reads_1.fq")] + ([os.path.join(d, "reads_2.fq")] if a.paired else [])
        def run(exe, fq_files, tag):
            sam = os.path.join(d, tag + ".sam")
            t0 = time.time()
            r = subprocess.run([exe, "-t", str(a.threads), "-o", sam, files[0]] + fq_files, capture_output=True, text=True)
            dt = time.time() - t0
            if r.returncode != 0:
                return {"error": r.stderr[-500:]}
            mapping = [ln for ln in r.stderr.splitlines() if "Total time" in ln or "indexing" in ln.lower()]
            res = {"wall_s": round(dt, 2), "sam_md5": md5_nopg(sam), "stderr_times": mapping[-8:]}
            veneer = [ln for ln in r.stderr.splitlines() if ln.startswith("[rsa_ext veneer]")]   # RSA_EXT_STATS=1
            if veneer:
                res["veneer_stats"] = veneer
            cost = [ln for ln in r.stderr.splitlines() if ln.startswith("cost time1")]  # the *_timed builds: per-worker phase timers
            if cost:
                res["worker_phase_timers"] = cost[:32]
            for ln in mapping:
                if "Total time mapping" in ln:
                    res["mapping_s"] = float(ln.split(":")[1].split()[0])
            os.remove(sam)
            return res

        subsets = [int(x) for x in a.subsets.split(",") if x] if not a.paired else []
        sub_files = {}
        for k in subsets:
            path = os.path.join(d, "sub_%d.fq" % k)
            with open(files[1], "rb") as f, open(path, "wb") as g:
                for _ in range(4 * k):
                    g.write(f.readline())
            sub_files[k] = path
        names = a.binaries.split(",") if a.binaries else ["rabbitsalign_cpussw", "rabbitsalign_gasalgpu", "rabbitsalign_b200",
                                                          "rabbitsalign_b200_alninfo", "rabbitsalign_b200_big", "rabbitsalign_b200_win",
                                                          "rabbitsalign_b200_gpuseed"]
        for name in names:
            exe = os.path.join(B, name)
            if not os.path.exists(exe):
                out[name] = "not built"
                continue
            res = run(exe, files[1:], name)
            walls = [res.get("wall_s")]
            for _ in range(a.repeat - 1):
                r2 = run(exe, files[1:], name)
                walls.append(r2.get("wall_s"))
                if "error" not in r2 and ("error" in res or r2["wall_s"] < res["wall_s"]):
                    r2_md5_same = r2.get("sam_md5") == res.get("sam_md5")
                    res = r2
                    res["sam_md5_stable"] = r2_md5_same
            if a.repeat > 1:
                res["wall_s_runs"] = walls
            if "error" not in res:
                res["reads_per_s_wall"] = round(out["reads"] / res["wall_s"])
                if res.get("mapping_s"):
                    res["reads_per_s_mapping"] = round(out["reads"] / res["mapping_s"])
            out[name] = res
            pts = [(out["reads"], res.get("mapping_s"))]
            for k in subsets:
                rk = run(exe, [sub_files[k]], "%s_%d" % (name, k))
                res["subset_%d" % k] = {kk: rk.get(kk) for kk in ("wall_s", "mapping_s", "error") if kk in rk}
                pts.append((k, rk.get("mapping_s")))
            pts = sorted(p for p in pts if p[1] is not None)
            if len(pts) >= 2 and pts[-1][0] > pts[0][0]:
                slope = (pts[-1][1] - pts[0][1]) / (pts[-1][0] - pts[0][0]) * 1e6
                res["mapping_s_per_mreads"] = round(slope, 3)
                res["startup_s"] = round(pts[0][1] - slope * pts[0][0] / 1e6, 2)
                res["steady_reads_per_s"] = round(1e6 / slope) if slope > 0 else None
    print(json.dumps(out))


if __name__ == "__main__":
    main()
