#!/usr/bin/env python
"""bench.py -- extension hot path benchmark (contract: see the task brief / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs P]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (packed/exact DP kernels + traceback -> 64-byte records) over one
synthetic batch shaped like BASELINE.json configs[1] (150 bp reads vs their NAM windows, 1% error, 5%
mate-rescue windows).  Metric: extension GCUPS, cells = sum |q|*|t|.

  value      device-resident: inputs already in HBM, K x rsa_ext_run_resident, CUDA events on the engine's
             compute stream, max over ranks.
  e2e        the same batch through the C ABI from pinned HOST buffers (rsa_ext_submit + rsa_ext_wait):
             H2D of the ASCII + plan, kernels, D2H of the records, every step, wall clock.
  roofline   the DP kernel is bound by the integer/DPX issue rate, so that is the roof reported: DP-phase
             GCUPS against the ALU-pipe issue rate measured live by tools/dpx_microbench x 64 cells / 12 instructions per
             packed cell pair; the kernel's own bare-recipe ceiling, the HBM view and the staging rates ride along.
  cpu_baseline / --impl reference
             the reference's own CPU extension path (Aligner::align = SSW, compiled from /root/reference
             into oracle/_ref) on the host cores, bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "extension_gcups"
UNIT = "GCUPS"


def workload_desc(pairs, read_len):
    return (f"BASELINE configs[1]-shaped extension batch: {pairs} (query,window) pairs per GPU per step, "
            f"{read_len} bp reads, windows read+flanks(0..50) (95%) / mate-rescue windows (5%), 1% substitutions, "
            f"0.2%/bp indel events; cells = sum |q|*|t|")


def config_dict(pairs, read_len):
    """Identical in both arms (the driver compares the dicts); everything arm-specific lives outside `config`."""
    return {"workload": workload_desc(pairs, read_len), "pairs_per_gpu_per_step": pairs, "read_len": read_len,
            "scoring": "match 2, mismatch 8, gap open 12, gap extend 1 (strobealign defaults)",
            "l2": "inputs + direction tiles of one step (> 20 GB) exceed the 126 MB L2; no flush needed",
            "note": "BASELINE configs[4] (extension microbench) is 10 M pairs of this shape; a step is 1 048 576 of them so "
                    "that warm-up, the CPU baseline and the extra legs fit a default run (GCUPS is size-normalised; "
                    "--pairs 10485760 runs the full size)"}


def make_batch(pairs, read_len, seed):
    from rabbitsalign_b200 import workload as W
    return W.extension_pairs_fast(pairs, read_len=read_len, seed=seed)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []   # (arrival time, csv line)
        self.t_begin = None

    def mark_begin(self):
        """Start of the timed region: only samples arriving after this count (nvidia-smi is started earlier, during
        warm-up, because its own start-up can exceed a short timed region, especially with 8 ranks on a box)."""
        self.t_begin = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        inside = [ln for t, ln in self.lines if self.t_begin is None or t >= self.t_begin]
        if not inside and self.lines:  # region shorter than the sampling period: the sample closest to it
            inside = [self.lines[-1][1]]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(batch, threads, steps=1, warmup=0):
    """Time the reference's CPU extension path on `batch`; returns (gcups, kind, seconds_per_step)."""
    import oracle
    ssw = oracle.ssw_reference()
    if ssw is not None:
        kind = "reference"

        def run():
            ssw.align_packed(batch.qbuf, batch.qoff, batch.tbuf, batch.toff, threads=threads)
    else:
        kind = "port"
        olib = oracle.restatement()
        bounds = np.linspace(0, batch.n, threads + 1).astype(int)
        parts = [batch.slice(int(bounds[k]), int(bounds[k + 1])) for k in range(threads) if bounds[k + 1] > bounds[k]]

        def run():
            ths = [threading.Thread(target=olib.align_packed, args=(p.qbuf, p.qoff, p.tbuf, p.toff)) for p in parts]
            [t.start() for t in ths]
            [t.join() for t in ths]
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / steps
    return batch.cells / dt / 1e9, kind, dt


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:  # noqa: BLE001
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def issue_ceiling():
    """Live integer/DPX issue-rate ceiling: tools/dpx_microbench runs the bare packed cell recipe (no memory,
    no shuffles) on every SM; returns cells/clk/SM and the chip GCUPS it reached, or None."""
    exe = os.path.join(ROOT, "tools", "dpx_microbench")
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe, "--quick"], capture_output=True, text=True, timeout=120).stdout
    except Exception:  # noqa: BLE001
        return None
    best = None
    alu = {}
    for ln in out.splitlines():
        try:
            r = json.loads(ln)
        except ValueError:
            continue
        if r.get("test", "").startswith("SW cell recipe"):
            if best is None or r["chip_gcups"] > best["chip_gcups"]:
                best = r
        # CUDA-event-timed issue rates of the ALU-pipe DPX instructions (whole chip, warp-instructions per second)
        if r.get("test") in ("VIMNMX3.S16x2", "VIADDMNMX.S16x2", "LOP3") and "chip_warp_ginstr_per_s" in r:
            alu[r["test"]] = max(alu.get(r["test"], 0.0), r["chip_warp_ginstr_per_s"])
    if best is None or "VIADDMNMX.S16x2" not in alu:
        return None
    return {"recipe_chip_gcups": best["chip_gcups"], "alu_warp_ginstr_per_s": alu["VIADDMNMX.S16x2"], "alu_rates": alu}


_PINNED_KEEP = []


def _pinned_copy(a):
    """A pinned host copy of `a` (the torch tensor that owns the memory is kept alive for the process)."""
    import torch
    t = torch.empty(a.nbytes, dtype=torch.uint8).pin_memory()
    v = t.numpy().view(a.dtype).reshape(a.shape)
    v[...] = a
    _PINNED_KEEP.append(t)
    return v


def leg_250bp_indel(device, scratch_gb):
    """BASELINE configs[3]: 250-bp reads at a 5 % indel-event rate (wide windows, ~30-run CIGARs).  8192 distinct pairs
    (scalar generator) tiled 64x to 524 288 pairs (four chunks, so the end-to-end leg overlaps copies and kernels like the
    main workload); resident (CUDA events) and end to end (submit/wait from pinned host memory, wall)."""
    import torch
    from rabbitsalign_b200 import ExtensionEngine, workload as W
    from rabbitsalign_b200.ext import RESULT_DTYPE
    u = W.extension_pairs(8192, seed=44, read_len=250, indel_rate=0.05, max_indel=4, fixed_query_len=False)
    reps = 64
    qbuf = _pinned_copy(np.tile(u.qbuf, reps))
    tbuf = _pinned_copy(np.tile(u.tbuf, reps))
    qoff = np.concatenate([u.qoff[:-1] + k * int(u.qoff[-1]) for k in range(reps)] + [np.array([reps * int(u.qoff[-1])])]).astype(np.int64)
    toff = np.concatenate([u.toff[:-1] + k * int(u.toff[-1]) for k in range(reps)] + [np.array([reps * int(u.toff[-1])])]).astype(np.int64)
    cells = float(u.cells) * reps
    n = u.n * reps
    eng = ExtensionEngine(device=device, scratch_bytes=scratch_gb << 30)
    eng.stage_resident(qbuf, qoff, tbuf, toff)
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", device))
    for _ in range(3):
        eng.run_resident()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5):
        eng.run_resident()
    e1.record(stream)
    e1.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res = eng.fetch_resident(n)
    results = _pinned_copy(np.zeros(n, dtype=RESULT_DTYPE))
    eng.submit(qbuf, qoff, tbuf, toff, results); eng.wait()
    t0 = time.perf_counter()
    for _ in range(3):
        eng.submit(qbuf, qoff, tbuf, toff, results); eng.wait()
    dt = (time.perf_counter() - t0) / 3
    out = {"pairs": n, "distinct_pairs": u.n, "mean_cigar_runs": float(np.mean(res["n_ops"])),
           "mean_window": float(np.mean(np.diff(u.toff))), "resident_gcups": cells / (ms * 1e-3) / 1e9,
           "e2e_gcups": cells / dt / 1e9, "records_equal": bool(results.tobytes() == res.tobytes())}
    eng.close()
    return out


def leg_seeding(device, cores, peaks):
    """SURVEY 8f rank 2, the second kernel family: randstrobe seeding + index lookup + NAM merge (+ rescue) for 1 M reads of
    150 bp against a 20 Mb genome with injected repeats.  The index is INPUT DATA built by the reference's own host code
    (StrobemerIndex::populate, out of scope) through oracle/_ref/libseed_ref.so -- the same library that provides this
    leg's CPU baseline (the reference's randstrobes_query + find_nams + find_nams_rescue on all host cores).  The kernels
    and the C ABI measured are the product's (include/rsa_seed.h)."""
    import oracle
    import torch
    from rabbitsalign_b200 import seed as S, workload as W
    if oracle.seed_reference_lib() is None:
        return None
    contigs = W.seeding_genome(n_contigs=4, contig_len=5_000_000, seed=41, repeat_families=8, copies_per_contig=20)
    idx = oracle.build_seed_index(contigs, 150, cores)
    small, soff = W.seeding_reads(contigs, 50_000, seed=42)
    reps = 20
    buf = np.tile(small, reps)
    off = np.concatenate([soff[:-1] + k * int(soff[-1]) for k in range(reps)] + [np.array([reps * int(soff[-1])])]).astype(np.int64)
    n = len(off) - 1
    tb = torch.empty(buf.nbytes, dtype=torch.uint8).pin_memory()
    pbuf = tb.numpy(); pbuf[...] = buf
    gi = S.SeedIndexGpu(S.make_config(idx.params(), device=device), idx.randstrobes, idx.starts)
    sd = S.Seeder(gi)
    sd.stage(pbuf, off)
    for _ in range(3):
        sd.run_staged()
    ms, ms_large = [], []
    for _ in range(5):
        sd.run_staged()
        st = sd.stats()
        ms.append(st["kernel_ms"]); ms_large.append(st["kernel_ms_large"])
    st = sd.stats()
    sd.find_nams(pbuf, off, copy=False)
    t0 = time.perf_counter()
    for _ in range(3):
        per, nams = sd.find_nams(pbuf, off, copy=False)  # H2D of the reads, kernels, D2H of the NAMs into pinned memory
    dt = (time.perf_counter() - t0) / 3
    st2 = sd.stats()
    sub = 200_000
    t0 = time.perf_counter()
    idx.time_find_nams(buf, np.ascontiguousarray(off[:sub + 1]), cores)
    t_cpu = time.perf_counter() - t0
    kms = float(np.median(ms))
    # algorithmic HBM bytes per read: its bases in, per query randstrobe one 32-byte sector of bucket starts, one of index
    # entries for the search, one for the filter probe, 16 bytes per hit entry read, 40 bytes per NAM + 16 per read out
    n_rs = 2 * 26  # ~26 randstrobes per strand at 150 bp (k 20, s 16)
    alg_bytes = float(off[-1]) + n * n_rs * 96.0 + 40.0 * st["nams"] + 16.0 * n
    out = {"reads": n, "genome_bp": 20_000_000, "index_entries": idx.n_randstrobes, "nams": st["nams"],
           "reads_rescued": st["reads_rescued"], "reads_large_tier": st["reads_retried"], "reads_failed": st["reads_failed"],
           "kernel_ms": kms, "kernel_ms_large_tier": float(np.median(ms_large)),
           "reads_per_s_resident": n / (kms * 1e-3), "nams_per_s_resident": st["nams"] / (kms * 1e-3),
           "reads_per_s_e2e": n / dt, "h2d_bytes": st2["h2d_bytes"], "d2h_bytes": st2["d2h_bytes"],
           "cpu_baseline": {"kind": "reference", "cores": cores, "reads_per_s": sub / t_cpu, "sample": f"first {sub} reads, one pass"},
           "roofline": {"bound": "hbm", "achieved": alg_bytes / (kms * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                        "frac": alg_bytes / (kms * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": None,
                        "note": "random 16-32 byte index probes: latency x sectors bound, not streaming bandwidth; algorithmic "
                                "bytes = read bases + 96 B per query randstrobe (bucket starts, entry, filter probe) + NAMs out"}}
    sd.close(); gi.close(); idx.close()
    return out


def leg_hamming(device, peaks):
    """the Hamming shortcut (hamming_distance + 5 % test + hamming_align, reference src/aln.cpp:391-404,
    src/aligner.cpp:219-302) for 1 M (read, equally long window) pairs of 150 bp through rsa_ext_hamming_align: blocking
    call from pinned host memory (H2D + kernel + D2H); CPU baseline = the reference's own functions, one thread."""
    import oracle
    import torch
    from rabbitsalign_b200 import ExtensionEngine
    rng = np.random.default_rng(7)
    n, L = 1 << 20, 150
    t = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n * L)]
    q = t.copy()
    rate = rng.choice([0.0, 0.01, 0.03, 0.08], size=n).repeat(L)
    flip = rng.random(n * L) < rate
    q[flip] = np.frombuffer(b"TGCA", np.uint8)[np.searchsorted(np.frombuffer(b"ACGT", np.uint8), q[flip])]
    off = (np.arange(n + 1, dtype=np.int64) * L)
    tq = torch.empty(q.nbytes, dtype=torch.uint8).pin_memory(); pq = tq.numpy(); pq[...] = q
    tt = torch.empty(t.nbytes, dtype=torch.uint8).pin_memory(); pt = tt.numpy(); pt[...] = t
    eng = ExtensionEngine(device=device)
    ham, aln = eng.hamming_align(pq, off, pt, off)
    t0 = time.perf_counter()
    for _ in range(3):
        ham, aln = eng.hamming_align(pq, off, pt, off)
    dt = (time.perf_counter() - t0) / 3
    kms = eng.stats()["dp_ms"]
    eng.close()
    sub = 1 << 17
    t0 = time.perf_counter()
    ref = oracle.hamming_reference(q, off[:sub + 1], t, off[:sub + 1])
    t_cpu = time.perf_counter() - t0
    out = {"pairs": n, "read_len": L, "shortcut_taken": int((aln["status"] == 0).sum()), "pairs_per_s_e2e": n / dt,
           "h2d_bytes": int(2 * n * L + 16 * n), "d2h_bytes": int(132 * n),
           "kernel_ms": kms,
           "roofline": {"bound": "hbm", "achieved": (2 * n * L + 16 * n + 132 * n) / (kms * 1e-3) / 1e9 if kms > 0 else None,
                        "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                        "frac": (2 * n * L + 16 * n + 132 * n) / (kms * 1e-3) / 1e9 / peaks["hbm_gbs"] if kms > 0 else None, "traffic": None,
                        "note": "kernel only (CUDA events): 300 B of bases + 16 B of offsets in, 132 B out per pair; the blocking call "
                                "around it is bound by its two PCIe copies"}}
    if ref is not None:
        same = bool((ham[:sub] == ref["hamming"]).all() and ((aln["status"][:sub] == 0) == (ref["status"] == 0)).all() and
                    (aln["sw_score"][:sub][ref["status"] == 0] == ref["score"][ref["status"] == 0]).all())
        out["cpu_baseline"] = {"kind": "reference", "cores": 1, "pairs_per_s": sub / t_cpu, "sample": f"first {sub} pairs, one thread"}
        out["equals_reference_on_sample"] = same
    return out


def leg_sam_format(device, peaks):
    """SAM text of 1 M single-end records (150-bp reads, half on the reverse strand) through
    rsa_sam_format (H2D of descriptors + read text, two kernels, D2H of the SAM text); CPU baseline = the reference's
    own class Sam (oracle/_ref/libsam_ref.so), one thread, same records; the texts must be identical."""
    import oracle
    from rabbitsalign_b200 import sam as S
    rng = np.random.default_rng(8)
    n, L = 1 << 20, 150
    names = np.char.add("read", np.char.zfill(np.arange(n).astype(str), 8)).astype("S12")
    name_bytes = np.frombuffer(names.tobytes(), np.uint8)
    seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n * L)]
    qual = rng.integers(35, 74, n * L).astype(np.uint8)
    text = np.concatenate([name_bytes, seq, qual])
    cig = np.zeros(3 * n, np.uint32)
    cig[0::3] = (70 << 4) | 7; cig[1::3] = (1 << 4) | 8; cig[2::3] = (79 << 4) | 7
    calls = np.zeros(n, oracle.SAM_CALL_DTYPE)
    calls["kind"] = 0; calls["is_primary"] = 1; calls["mapq1"] = 60
    a = calls["a1"]
    a["ref_id"] = rng.integers(0, 4, n); a["ref_start"] = rng.integers(0, 25_000_000, n); a["edit_distance"] = 1
    a["score"] = 290; a["length"] = L; a["is_rc"] = rng.integers(0, 2, n); a["cigar_off"] = 3 * np.arange(n); a["n_cigar"] = 3
    r = calls["r1"]
    r["name_off"] = 12 * np.arange(n); r["name_len"] = 12
    r["seq_off"] = 12 * n + L * np.arange(n); r["seq_len"] = L
    r["qual_off"] = 12 * n + L * n + L * np.arange(n); r["qual_len"] = L
    rec = np.zeros(n, S.RECORD_DTYPE)
    rec["kind"] = S.KIND_ALIGNED; rec["flags"] = np.where(a["is_rc"] != 0, 0x10, 0); rec["ref_id"] = a["ref_id"]
    rec["pos"] = a["ref_start"]; rec["mapq"] = 60; rec["mate_ref"] = -1; rec["mate_pos"] = 0xFFFFFFFF; rec["tlen"] = 0
    rec["edit_distance"] = 1; rec["score"] = 290; rec["cigar_off"] = a["cigar_off"]; rec["n_cigar"] = 3
    for k in ("name", "seq", "qual"):
        rec[k + "_off"] = r[k + "_off"]; rec[k + "_len"] = r[k + "_len"]
    ref_names = [b"contig1", b"contig2", b"contig3", b"contig4"]
    import torch

    def pinned(a):
        t = torch.empty(a.nbytes, dtype=torch.uint8).pin_memory()
        v = t.numpy().view(a.dtype).reshape(a.shape)
        v[...] = a
        return t, v
    keep = []
    for name in ("rec", "text", "cig"):
        t, v = pinned({"rec": rec, "text": text, "cig": cig}[name])
        keep.append(t)
        if name == "rec":
            rec = v
        elif name == "text":
            text = v
        else:
            cig = v
    tout = torch.empty(int(text.nbytes) + 200 * n, dtype=torch.uint8).pin_memory()
    f = S.SamFormatter(ref_names, device=device)
    got = f.format(rec, text, cig, out=tout.numpy())
    t0 = time.perf_counter()
    for _ in range(3):
        got = f.format(rec, text, cig, out=tout.numpy())
    dt = (time.perf_counter() - t0) / 3
    kms = f.kernel_ms()
    got = got.tobytes()
    f.close()
    out = {"records": n, "read_len": L, "sam_bytes": len(got), "records_per_s_e2e": n / dt, "sam_gb_per_s_e2e": len(got) / dt / 1e9,
           "h2d_bytes": int(rec.nbytes + text.nbytes + cig.nbytes), "d2h_bytes": len(got),
           "kernel_ms": kms,
           "roofline": {"bound": "hbm", "achieved": (rec.nbytes + text.nbytes + cig.nbytes + len(got)) / (kms * 1e-3) / 1e9 if kms > 0 else None,
                        "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                        "frac": (rec.nbytes + text.nbytes + cig.nbytes + len(got)) / (kms * 1e-3) / 1e9 / peaks["hbm_gbs"] if kms > 0 else None,
                        "traffic": None,
                        "note": "kernels only (length + scan + writer, CUDA events): algorithmic bytes = descriptors + read text + CIGAR ops in, "
                                "SAM text out; the blocking call around them is bound by its two PCIe copies (pinned host memory)"}}
    t0 = time.perf_counter()
    want = oracle.sam_reference_replay(ref_names, calls, text, cig)
    t_cpu = time.perf_counter() - t0
    if want is not None:
        out["cpu_baseline"] = {"kind": "reference", "cores": 1, "records_per_s": n / t_cpu, "sample": "all records, one thread"}
        out["text_identical_to_reference"] = bool(got == want)
    return out


def pipeline_block(threads):
    """BASELINE metric (i), end-to-end reads/s: the reference's host pipeline (integration/_build, compiled from the
    reference by integration/build.sh) with the reference's own GPU path vs this engine, same synthetic FASTQ/FASTA,
    same threads.  A small job (1.5 M single-end reads, 50 Mb): process start-up weighs in (CUDA initialisation inside a
    busy process varies by seconds), so every binary runs twice, the better run counts, and the pipeline's own
    "Total time mapping" is reported next to the wall clock.  Runs at BASELINE scale: profiles/r2_e2e_*.json."""
    exe = os.path.join(ROOT, "tools", "e2e_reads_bench.py")
    # the reference as shipped; this engine behind the unmodified caller; the full device path (windows by offset, GPU seeding,
    # Hamming shortcut and SAM text on the device -- INTEGRATION.md)
    full = "rabbitsalign_b200_gpusam" if os.path.exists(os.path.join(ROOT, "integration", "_build", "rabbitsalign_b200_gpusam")) \
        else "rabbitsalign_b200_gpuseed"
    bins = ["rabbitsalign_gasalgpu", "rabbitsalign_b200_big", full]
    if full != "rabbitsalign_b200_gpuseed":
        bins.append("rabbitsalign_b200_gpuseed")   # reported beside it (same SAM bytes; the speed-ups below use `full`)
    if not all(os.path.exists(os.path.join(ROOT, "integration", "_build", b)) for b in bins[:2]):
        return None
    try:
        r = subprocess.run([sys.executable, exe, "--ref-len", "50000000", "--reads", "1500000", "--threads", str(threads),
                            "--binaries", ",".join(bins), "--repeat", "2"], capture_output=True, text=True, timeout=420)
        d = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as ex:  # noqa: BLE001
        return {"error": str(ex)[:200]}
    out = {"reads": d["reads"], "ref_len": d["ref_len"], "threads": d["threads"]}
    for b in bins:
        if isinstance(d.get(b), dict):
            out[b] = {k: d[b].get(k) for k in ("wall_s", "mapping_s", "reads_per_s_wall", "reads_per_s_mapping", "sam_md5")}
    try:
        out["sam_identical"] = len({out[b]["sam_md5"] for b in bins if b in out}) == 1
        out["mapping_speedup_vs_reference_gpu_build"] = out[bins[0]]["mapping_s"] / out[bins[2]]["mapping_s"]
        out["wall_speedup_vs_reference_gpu_build"] = out[bins[0]]["wall_s"] / out[bins[2]]["wall_s"]
    except Exception:  # noqa: BLE001
        pass
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100, help="timed steps (100 x ~20 ms: a 2-s timed region per leg)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1 << 20, help="pairs per GPU per step")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--cpu-sample", type=int, default=1 << 17, help="pairs in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the 250-bp, seeding, Hamming, SAM and pipeline legs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = max(args.gpus, world) if world > 1 else args.gpus
    cores = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm (CPU) -------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = min(args.pairs, args.cpu_sample)
        b = make_batch(sample, args.read_len, seed=43)
        g, kind, dt = cpu_reference_run(b, cores, steps=max(1, args.steps), warmup=min(args.warmup, 1))
        line = {
            "impl": "reference", "metric": METRIC, "value": g, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "s16/u8 SSE2", "data": "synthetic",
            "config": config_dict(args.pairs, args.read_len),
            "detail": {"note": "reference CPU path (Aligner::align: SSW striped SW + banded traceback + end bonus) on all "
                               "host cores; no GPU involved, so the value does not grow with --gpus",
                       "build": "oracle/_ref/libssw_ref_v3.so: the reference's ext/ssw + src/aligner.cpp compiled "
                                "-O3 -march=x86-64-v3 (the reference's build.sh:49 uses -march=native on its build host; "
                                "the library is built in the dev container and must run on the GPU box)"},
            "cpu_baseline": {"value": g, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{sample} pairs of the step's batch per step"},
            "e2e": {"value": g, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm ---------------------------
    affinity = None
    if world > 1 and hasattr(os, "sched_setaffinity") and os.environ.get("RSA_BENCH_NO_AFFINITY") != "1":
        # one rank per GPU on one box: keep each rank (its main thread, the engine's plan-ahead thread, the pinned staging
        # it first-touches) on its own contiguous slice of the host cores, so that eight ranks do not migrate across
        # sockets while each streams ~0.5 GB per step to its GPU
        try:
            cpus = sorted(os.sched_getaffinity(0))
            per = max(1, len(cpus) // world)
            mine = cpus[local_rank * per:(local_rank + 1) * per] or cpus
            os.sched_setaffinity(0, mine)
            affinity = [mine[0], mine[-1]]
        except OSError:
            affinity = None
    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device; this benchmark has no CPU path"}))
        return 1
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        # NCCL announces itself on stdout ("NCCL version ..."); keep stdout clean for the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    from rabbitsalign_b200 import ExtensionEngine
    from rabbitsalign_b200.ext import RESULT_DTYPE

    from rabbitsalign_b200 import sharding
    batch = make_batch(args.pairs, args.read_len, seed=sharding.rank_seed(43, rank))
    scratch_gb = int(os.environ.get("RSA_EXT_SCRATCH_GB", "0"))  # tuning knob (experiments)
    eng = ExtensionEngine(device=local_rank, scratch_bytes=scratch_gb << 30)
    sampler = ClockSampler(local_rank)  # started now: nvidia-smi needs up to a second before its first sample
    sampler.start()

    # pinned host copies (the C ABI copies straight from/to pinned memory)
    def pinned(a):
        t = torch.empty(a.nbytes, dtype=torch.uint8).pin_memory()
        v = t.numpy().view(a.dtype).reshape(a.shape)
        v[...] = a
        return t, v
    keep = []
    tq, qbuf = pinned(batch.qbuf); keep.append(tq)
    tt, tbuf = pinned(batch.tbuf); keep.append(tt)
    tqo, qoff = pinned(batch.qoff); keep.append(tqo)
    tto, toff = pinned(batch.toff); keep.append(tto)
    tres = torch.empty(batch.n * RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    results = tres.numpy().view(RESULT_DTYPE)

    # ---- device-resident leg: `value` ------------------------------------------------------------------
    eng.stage_resident(qbuf, qoff, tbuf, toff)
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local_rank))
    warmup = max(3, args.warmup)  # the timing rules ask for at least three warm-up steps
    for _ in range(warmup):
        eng.run_resident()
    torch.cuda.synchronize()
    barrier()
    sampler.mark_begin()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        eng.run_resident()
    e1.record(stream)
    e1.synchronize()
    barrier()
    ms_total = e0.elapsed_time(e1)
    st = eng.stats()  # dp_ms / tb_ms of the last step (events around the kernels)
    clocks = sampler.stop()
    launches_per_step = st["kernel_launches"]
    t_max, total_cells = sharding.reduce_step(ms_total, float(batch.cells), dist, device="cuda")
    ms_step = t_max / args.steps
    value = total_cells / (ms_step * 1e-3) / 1e9
    res_resident = eng.fetch_resident(batch.n)

    # ---- clean per-kernel durations for the roofline: same batch, kernels serialised on one stream --------------
    eng_s = ExtensionEngine(device=local_rank, serialize=True, scratch_bytes=scratch_gb << 30)
    eng_s.stage_resident(qbuf, qoff, tbuf, toff)
    for _ in range(3):
        eng_s.run_resident()
    st_serial = eng_s.stats()
    eng_s.close()

    # ---- end-to-end leg through the C ABI from host buffers ----------------------------------------------
    for _ in range(2):
        eng.submit(qbuf, qoff, tbuf, toff, results)
        eng.wait()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.submit(qbuf, qoff, tbuf, toff, results)
        eng.wait()
    dt = time.perf_counter() - t0
    st_e2e = eng.stats()
    dt_max, _ = sharding.reduce_step(dt, 0.0, dist, device="cuda")
    e2e_value = total_cells / (dt_max / args.steps) / 1e9
    # ---- same end-to-end leg with the device-side gasal_fail + align_gpu post-processing (SURVEY 8f rows 1+3) ----
    from rabbitsalign_b200.ext import ALNINFO_DTYPE
    taln = torch.empty(batch.n * ALNINFO_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    aln = taln.numpy().view(ALNINFO_DTYPE)
    eng.request_alninfo(aln, end_bonus=10)
    eng.submit(qbuf, qoff, tbuf, toff, results)
    eng.wait()
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps // 2)):
        eng.submit(qbuf, qoff, tbuf, toff, results)
        eng.wait()
    dt_aln = (time.perf_counter() - t0) / max(1, args.steps // 2)
    eng.request_alninfo(None)
    aln_stats = {"gcups": batch.cells / dt_aln / 1e9, "ms_per_step": dt_aln * 1e3,
                 "accepted": int((aln["status"] == 0).sum()), "gasal_fail": int((aln["status"] == 1).sum()),
                 "long_cigar": int((aln["status"] == 3).sum())}

    # ---- end-to-end with the windows named by (offset, length) in a reference resident in HBM: the
    #      step's concatenated window buffer plays the reference, so the records must equal the explicit form's
    win_stats = None
    if len(tbuf) < (1 << 32):
        eng.set_reference(tbuf)
        win_off = np.ascontiguousarray(toff[:-1])
        win_len = np.diff(toff).astype(np.int32)
        lib = eng.lib

        def win_step():
            rc = lib.rsa_ext_submit_ref_windows(eng.h, batch.n, qbuf.ctypes.data, qoff.ctypes.data, win_off.ctypes.data,
                                                win_len.ctypes.data, results.ctypes.data)
            assert rc == 0, rc
            eng.wait()
        win_step()
        t0 = time.perf_counter()
        for _ in range(max(1, args.steps // 2)):
            win_step()
        dt_win = (time.perf_counter() - t0) / max(1, args.steps // 2)
        dt_win_max, _ = sharding.reduce_step(dt_win, 0.0, dist, device="cuda")   # every rank runs this leg: max over ranks
        win_stats = {"gcups": batch.cells / dt_win / 1e9, "ms_per_step": dt_win * 1e3,
                     "all_ranks_gcups": total_cells / dt_win_max / 1e9,
                     "h2d_bytes_per_step": eng.stats()["h2d_bytes"],
                     "records_equal_explicit_form": bool(results.tobytes() == res_resident.tobytes())}

    # ---- the reference's own call shape: blocking 512-pair slices (STREAM_BATCH_SIZE, src/pc.cpp:644-672) -----
    n_slices = min(200, batch.n // 512)
    t0 = time.perf_counter()
    for k in range(n_slices):
        eng.submit_raw(512, qbuf.ctypes.data, qoff.ctypes.data + 8 * 512 * k, tbuf.ctypes.data,
                       toff.ctypes.data + 8 * 512 * k, results.ctypes.data + 64 * 512 * k)
        eng.wait()
    slice_dt = (time.perf_counter() - t0) / max(1, n_slices)
    eng.submit(qbuf, qoff, tbuf, toff, results)  # restore the full-batch records for the check below
    eng.wait()
    same = bool(results.tobytes() == res_resident.tobytes())
    ok = bool((results["status"] == 0).all() and (results["score"] > 0).mean() > 0.99)

    if rank != 0:
        eng.close()
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline --------------------------------------------------------------------------------------
    peaks, peak_kind = measured_peaks()
    dp_ms = st_serial["dp_ms"]     # DP kernels of one step, not overlapped with anything (CUDA events on their stream)
    tb_ms = st_serial["tb_ms"]
    # algorithmic HBM bytes of the DP phase per pair: ASCII in, direction nibbles out, 16-byte end record
    ql = np.diff(batch.qoff); tl = np.diff(batch.toff)
    dp_bytes = float(np.sum(ql + tl) + np.sum(ql * tl) / 2 + 16 * batch.n)
    hbm_achieved = dp_bytes / (dp_ms * 1e-3) / 1e9 if dp_ms > 0 else None
    # measured DRAM traffic of that kernel: ncu --set full capture of this round (profiles/*_traffic.json), scaled from
    # bytes per cell to the bytes of one average launch of this run
    traffic = None
    n_dp_launches = max(1, (launches_per_step - 0) // 4)  # per chunk: packed DP, redo, group traceback, traceback
    try:
        tj = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_traffic.json"))[-1]
        traffic = json.load(open(os.path.join(ROOT, "profiles", tj)))["dram_bytes_per_cell"] * batch.cells / n_dp_launches
    except Exception:  # noqa: BLE001
        pass
    # The binding roof of the DP kernel is the integer/DPX issue rate, so that is what `roofline` reports:
    #   achieved = DP-phase GCUPS (CUDA events around the DP kernels, serialised engine: nothing overlaps them);
    #   peak     = measured ALU-pipe issue rate (tools/dpx_microbench: VIADDMNMX.S16x2 warp-instructions per second over
    #              the whole chip, CUDA events, this run) x 64 cells per warp-instruction (32 lanes x s16x2)
    #              / 12 instructions per packed cell pair.  The 12 is SURVEY 8d's floor: 15 integer ops per cell with
    #              traceback (1 compare/select, 1 diag add, 3 max for H incl. the 0 floor, 1 shared tmp-GapOE, 2 x (add+max)
    #              for E'/F', 4 direction predicates, 1 running max); DPX fuses add+max and the 3-way max+relu, leaving
    #              ~12 instructions for two packed cells.  Recipe-independent: nothing of this kernel enters the peak.
    #   traffic  = DRAM bytes per DP launch from the ncu --set full capture (bytes per cell x cells of an average launch).
    kInstrFloor = 12
    ceil = issue_ceiling()
    dp_gcups = batch.cells / (dp_ms * 1e-3) / 1e9 if dp_ms > 0 else None
    hbm = {"achieved": hbm_achieved, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
           "frac": (hbm_achieved / peaks["hbm_gbs"]) if hbm_achieved else None,
           "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
           "note": "algorithmic bytes of the DP phase (ASCII in + 4-bit direction tiles + 16-byte end records) per second; "
                   "far from the HBM roof by design"}
    # window-staging phase (north_star): the staging is fused into the DP kernel (ASCII -> codes/profiles in shared memory
    # and registers), so its bytes move at the DP kernel's pace; the dominant stream is the direction tiles going out
    staging = {"ascii_in_gbs": float(np.sum(ql + tl)) / (dp_ms * 1e-3) / 1e9 if dp_ms > 0 else None,
               "direction_tiles_out_gbs": float(np.sum(ql * tl) / 2) / (dp_ms * 1e-3) / 1e9 if dp_ms > 0 else None,
               "note": "fused into fast_dp_kernel: no separate staging kernel exists; rates = bytes / DP-phase time"}
    roofline = None
    if ceil and dp_gcups:
        peak = ceil["alu_warp_ginstr_per_s"] * 64.0 / kInstrFloor
        roofline = {"bound": "int/DPX issue", "kernel": "fast_dp_kernel<L,C> (packed s16x2 DP + direction nibbles)",
                    "achieved": dp_gcups, "peak": peak, "unit": "GCUPS", "frac": dp_gcups / peak, "traffic": traffic,
                    "algorithmic_bytes_per_launch": dp_bytes / n_dp_launches, "launches_per_step": n_dp_launches,
                    "peak_source": "tools/dpx_microbench in this run: VIADDMNMX.S16x2 issue rate (CUDA events, whole chip) "
                                   f"x 64 cells / {kInstrFloor} instructions per packed cell pair (SURVEY 8d)",
                    "alu_warp_ginstr_per_s": ceil["alu_warp_ginstr_per_s"], "alu_rates_ginstr_per_s": ceil["alu_rates"],
                    "instr_per_packed_pair_floor": kInstrFloor, "dp_ms": dp_ms, "tb_ms": tb_ms,
                    "own_recipe_ceiling": {"chip_gcups": ceil["recipe_chip_gcups"], "frac": dp_gcups / ceil["recipe_chip_gcups"],
                                           "note": "rsa::fast_cell on registers only (no memory, shuffles, skew): loop overhead "
                                                   "of the kernel, not distance from the hardware"},
                    "hbm": hbm, "staging": staging}
    else:
        roofline = {"bound": "int/DPX issue", "achieved": dp_gcups, "peak": None, "unit": "GCUPS", "frac": None,
                    "traffic": traffic, "note": "tools/dpx_microbench unavailable: no live ALU issue rate", "hbm": hbm,
                    "staging": staging}

    # ---- extra legs (rank 0, N=1 only): BASELINE configs[3] and the end-to-end pipeline ----------------------------------
    leg250 = None
    pipe = None
    seeding = None
    hamming = None
    samfmt = None
    if n_gpus == 1 and not args.no_extra_legs:
        try:
            leg250 = leg_250bp_indel(local_rank, scratch_gb)
        except Exception as ex:  # noqa: BLE001
            leg250 = {"error": str(ex)[:200]}
        eng.close()  # the pipeline binaries need the GPU memory and the host cores
        eng = None
        try:
            seeding = leg_seeding(local_rank, cores, peaks)
        except Exception as ex:  # noqa: BLE001
            seeding = {"error": str(ex)[:300]}
        for name, fn in (("hamming", leg_hamming), ("samfmt", leg_sam_format)):
            try:
                res_leg = fn(local_rank, peaks)
            except Exception as ex:  # noqa: BLE001
                res_leg = {"error": str(ex)[:300]}
            if name == "hamming":
                hamming = res_leg
            else:
                samfmt = res_leg
        pipe = pipeline_block(cores)

    # ---- CPU baseline (rank 0, N=1 only) ------------------------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline and n_gpus == 1:
        sample = min(batch.n, args.cpu_sample)
        g, kind, _ = cpu_reference_run(batch.slice(0, sample), cores)
        cpu = {"value": g, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"first {sample} pairs of the step's batch, one pass"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": warmup,
        "warmup_requested": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "s16x2", "data": "synthetic",
        "config": config_dict(args.pairs, args.read_len),
        "detail": {"cells_per_gpu_per_step": batch.cells, "pairs_per_s": batch.n * n_gpus / (ms_step * 1e-3),
                   "routing": {"packed": st["pairs_fast"], "exact": st["pairs_exact"], "failed": st["pairs_failed"],
                               "redo_last_chunk": st["pairs_redo"]},
                   "resident_equals_e2e_records": same, "records_sane": ok, "rank0_cpu_affinity": affinity,
                   "slice512_one_worker": {"us_per_call": slice_dt * 1e6, "pairs_per_s": 512 / slice_dt if slice_dt > 0 else None},
                   "e2e_with_device_align_gpu": aln_stats, "e2e_windows_in_resident_reference": win_stats,
                   "leg_250bp_5pct_indel": leg250, "seeding": seeding, "hamming_shortcut": hamming,
                   "sam_format": samfmt, "pipeline": pipe},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": st_e2e["h2d_bytes"],
                "d2h_bytes_per_step": st_e2e["d2h_bytes"],
                "ms_per_step": dt_max / args.steps * 1e3, "host_plan_ms_per_step": st_e2e["host_plan_ms"]},
        "gpu_launches": int(launches_per_step * args.steps),
        "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if eng is not None:
        eng.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
