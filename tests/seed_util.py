"""Helpers of the seeding parity tests: the host-compiled check harness (tests/seed_host_check.cu), the comparison with
the reference's seeding path (oracle/_ref/libseed_ref.so) and the shared cases."""
import ctypes as C
import os
import subprocess

import numpy as np

import oracle
from rabbitsalign_b200 import seed as S, workload as W

HERE = os.path.dirname(os.path.abspath(__file__))
FIELDS = ["query_start", "query_end", "query_prev_hit_startpos", "ref_start", "ref_end", "ref_prev_hit_startpos",
          "n_hits", "ref_id", "score"]

# name -> (genome kwargs, read_len for the index profile, reads kwargs)
CASES = {
    "r150_repeats": (dict(n_contigs=3, contig_len=600_000, seed=5), 150, dict(n=6000, seed=6, n_rate=0.0005)),
    "r100": (dict(n_contigs=2, contig_len=400_000, seed=7, repeat_families=2), 100, dict(n=4000, read_len=100, seed=8)),
    "r250_indel": (dict(n_contigs=2, contig_len=400_000, seed=9), 250, dict(n=3000, read_len=250, seed=10, indel_rate=0.01, sub_rate=0.02)),
    "r400": (dict(n_contigs=2, contig_len=300_000, seed=11), 400, dict(n=1500, read_len=400, seed=12)),
    "r50_short": (dict(n_contigs=2, contig_len=200_000, seed=13), 50, dict(n=3000, read_len=50, seed=14, vary_len=True)),
    "many_contigs": (dict(n_contigs=40, contig_len=30_000, seed=15, repeat_families=6, copies_per_contig=2, family_len=600), 150,
                     dict(n=4000, seed=16)),
    "N_rich": (dict(n_contigs=2, contig_len=300_000, seed=17), 150, dict(n=3000, seed=18, n_rate=0.02)),
    # a genome that is mostly repeats: most randstrobes are filtered -> nonrepetitive fraction < 0.7 -> rescue mode
    "rescue_heavy": (dict(n_contigs=2, contig_len=300_000, seed=19, repeat_families=2, copies_per_contig=120, family_len=1200,
                          divergence=0.005), 150, dict(n=3000, seed=20)),
    "low_complexity": (dict(n_contigs=2, contig_len=100_000, seed=21, low_complexity=400), 150, dict(n=3000, seed=22)),
}


def make_case(name):
    g, rl, r = CASES[name]
    contigs = W.seeding_genome(**g)
    idx = oracle.build_seed_index(contigs, rl, 4)
    buf, off = W.seeding_reads(contigs, **r)
    return idx, buf, off


def host_harness():
    """Build (once) and load tests/_build/libseed_host_check.so: the product's per-read seeding code compiled for the host."""
    out = os.path.join(HERE, "_build", "libseed_host_check.so")
    src = os.path.join(HERE, "seed_host_check.cu")
    hdr = os.path.join(HERE, "..", "rabbitsalign_b200", "csrc", "kernels_seed.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared", "-o", out, src])
    lib = C.CDLL(out)
    lib.seed_host_check.restype = C.c_int64
    lib.seed_host_check.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_int64]
    return lib


def host_seed(lib, idx, buf, off, cfg=None):
    """Small tier for every read, large tier for the reads that overflow it (what rsa_seed_find_nams does on the GPU)."""
    n = len(off) - 1
    cfg = cfg or S.make_config(idx.params())
    per = np.zeros(n, S.READ_DTYPE)
    nams = np.zeros(n * 128 + 1024, S.NAM_DTYPE)
    tot = lib.seed_host_check(C.byref(cfg), idx.randstrobes.ctypes.data, idx.n_randstrobes, idx.starts.ctypes.data, n,
                              buf.ctypes.data, off.ctypes.data, 0, per.ctypes.data, nams.ctypes.data, len(nams))
    assert tot >= 0
    nams = nams[:tot]
    bad = np.nonzero(per["flags"] & S.READ_FAILED)[0]
    if len(bad):
        lens = np.diff(off)[bad]
        o2 = np.zeros(len(bad) + 1, np.int64)
        o2[1:] = np.cumsum(lens)
        b2 = np.concatenate([buf[off[i]:off[i + 1]] for i in bad])
        per2 = np.zeros(len(bad), S.READ_DTYPE)
        n2 = np.zeros(len(bad) * 40000, S.NAM_DTYPE)
        t2 = lib.seed_host_check(C.byref(cfg), idx.randstrobes.ctypes.data, idx.n_randstrobes, idx.starts.ctypes.data, len(bad),
                                 b2.ctypes.data, o2.ctypes.data, 1, per2.ctypes.data, n2.ctypes.data, len(n2))
        assert t2 >= 0
        per2["nam_off"] += len(nams)
        nams = np.concatenate([nams, n2[:t2]])
        per[bad] = per2
    return per, nams, len(bad)


def assert_equals_reference(idx, buf, off, per, nams, rescue_level=2):
    """Every field of every NAM, in the reference's order; the nonrepetitive fraction; the rescue decision."""
    cnt, frac, resc, ref = idx.find_nams(buf, off, rescue_level=rescue_level)
    assert int((per["flags"] & S.READ_FAILED).sum()) == 0
    bad = np.nonzero(per["n_nams"] != cnt)[0]
    assert len(bad) == 0, f"NAM counts differ for reads {bad[:10]}"
    assert (per["nonrepetitive_fraction"] == frac).all()
    assert ((per["flags"] & S.READ_RESCUED) == resc).all()
    ordered = S.apply_group_order(per, nams, idx.map_order)
    for f in FIELDS:
        d = np.nonzero(ordered[f] != ref[f])[0]
        assert len(d) == 0, f"{f} differs at NAMs {d[:10]}"
    assert ((ordered["flags"] & 1) == ref["is_rc"]).all()
    return cnt, resc
