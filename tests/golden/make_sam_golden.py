"""Generate tests/golden/sam_golden.json -- run in the dev container (needs /root/reference built through
integration/build.sh).  For each config: synthetic inputs from tools/make_reads.py (deterministic), the
reference's host pipeline with the reference's own GASAL2 kernels compiled for the host
(integration/_build/rabbitsalign_gasalref) -> md5 of the SAM without its @PG line (the only line holding the
command line, reference src/main.cpp:97).  The same pipeline with the extension forced onto the reference's CPU
SSW path (rabbitsalign_cpussw) is recorded too: the product must match the FIRST md5, not the second.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
B = os.path.join(ROOT, "integration", "_build")

CONFIGS = {
    # BASELINE.json configs[0]: phiX-sized reference, 10k x 150 bp single-end
    "cfg1_phix_se150": dict(reads=["--ref-len", "5386", "--reads", "10000", "--seed", "42"], paired=False, threads=4),
    # configs[3] shape: 250 bp reads, indel-rich
    "cfg4_se250_indel": dict(reads=["--ref-len", "200000", "--contigs", "2", "--reads", "6000", "--read-len", "250",
                                    "--indel", "0.02", "--max-indel", "4", "--seed", "7"], paired=False, threads=4),
    # reads carrying N: the GASAL path and the SSW path disagree on >20% of the records here
    "se150_N": dict(reads=["--ref-len", "100000", "--reads", "6000", "--n-rate", "0.01", "--seed", "9"], paired=False, threads=4),
    # configs[1] shape, scaled down: paired-end with mate rescue (one worker: the non-FX PE path keeps one
    # insert-size estimator per worker, reference src/pc.cpp:1581,1748)
    "pe150": dict(reads=["--ref-len", "1000000", "--contigs", "4", "--reads", "10000", "--paired", "--seed", "11"], paired=True, threads=1),
    # the reference's SHIPPED build configuration (RabbitFX reader + OPT_NUMA_CLOSE, build.sh:49): several 4-MiB
    # chunks per file, 4 workers (FX-PE keeps one insert-size estimator per chunk, so it is thread-count independent)
    "fx_pe150_t4": dict(reads=["--ref-len", "2000000", "--contigs", "4", "--reads", "50000", "--paired", "--seed", "13"],
                        paired=True, threads=4, fx=True),
    "fx_se150_t4": dict(reads=["--ref-len", "1000000", "--contigs", "2", "--reads", "60000", "--indel", "0.006", "--seed", "15"],
                        paired=False, threads=4, fx=True),
}


def md5_file(path, skip_pg=False):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for line in f:
            if skip_pg and line.startswith(b"@PG"):
                continue
            h.update(line)
    return h.hexdigest()


def run_config(name, cfg, binary, workdir):
    d = os.path.join(workdir, name)
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_reads.py"), d] + cfg["reads"])
    files = ["ref.fa", "reads_1.fq"] + (["reads_2.fq"] if cfg["paired"] else [])
    out = os.path.join(d, os.path.basename(binary) + ".sam")
    subprocess.check_call([binary, "-t", str(cfg["threads"]), "-o", out] + [os.path.join(d, f) for f in files],
                          stderr=subprocess.DEVNULL)
    return {"inputs": {f: md5_file(os.path.join(d, f)) for f in files}, "sam_md5": md5_file(out, skip_pg=True),
            "records": sum(1 for ln in open(out, "rb") if not ln.startswith(b"@"))}


if __name__ == "__main__":
    gold = {}
    with tempfile.TemporaryDirectory() as wd:
        for name, cfg in CONFIGS.items():
            pre = "rabbitsalign_fx_" if cfg.get("fx") else "rabbitsalign_"
            g = run_config(name, cfg, os.path.join(B, pre + "gasalref"), wd)
            c = run_config(name, cfg, os.path.join(B, pre + "cpussw"), wd)
            gold[name] = {"make_reads_args": cfg["reads"], "paired": cfg["paired"], "threads": cfg["threads"],
                          "fx": bool(cfg.get("fx")),
                          "inputs": g["inputs"], "records": g["records"], "sam_md5_gasal_semantics": g["sam_md5"],
                          "sam_md5_cpu_ssw_path": c["sam_md5"]}
            print(name, gold[name]["sam_md5_gasal_semantics"], gold[name]["sam_md5_cpu_ssw_path"])
    json.dump(gold, open(os.path.join(ROOT, "tests", "golden", "sam_golden.json"), "w"), indent=1)
