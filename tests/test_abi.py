"""CPU suite: the C-ABI library loads, exports every symbol include/rsa_ext.h declares, refuses to run
without a GPU, and its host-only logic (planning, CIGAR text) behaves."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from rabbitsalign_b200 import ext, workload as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_every_declared_symbol():
    lib = ext.load_library()
    header = open(os.path.join(ROOT, "include", "rsa_ext.h")).read()
    declared = set(re.findall(r"\b(rsa_ext_[a-z_]+)\s*\(", header))
    assert declared == set(ext.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.rsa_ext_version() >= 1


def test_exports_every_declared_seeding_symbol():
    from rabbitsalign_b200 import seed
    lib = ext.load_library()
    header = open(os.path.join(ROOT, "include", "rsa_seed.h")).read()
    declared = set(re.findall(r"\b(rsa_seed_[a-z_]+)\s*\(", header))
    assert declared == set(seed.SEED_ABI_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert C.sizeof(seed.SeedConfig) == 56 and seed.NAM_DTYPE.itemsize == 40 and seed.READ_DTYPE.itemsize == 16


def test_seeding_has_no_cpu_fallback():
    import torch
    from rabbitsalign_b200 import seed
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = seed.make_config(dict(bits=8, filter_cutoff=10, k=20, s=16, t_syncmer=3, w_min=5, w_max=11, max_dist=80, q=255))
    with pytest.raises(seed.SeedError) as ei:
        seed.SeedIndexGpu(cfg, np.zeros(16, np.uint8), np.zeros(257, np.uint64))
    assert "no CPU path" in str(ei.value)


def test_result_record_is_64_bytes():
    assert ext.RESULT_DTYPE.itemsize == 64


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ext.ExtensionError) as ei:
        ext.ExtensionEngine()
    assert ei.value.status == -2 and "no CPU path" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "rabbitsalign_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "oracle/" not in src and "liboracle" not in src, f


def test_rle_to_text_matches_oracle_decode(oracle_lib):
    rng = np.random.default_rng(5)
    olib = oracle_lib.lib
    olib.rsa_oracle_cigar_text.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    for _ in range(300):
        n = int(rng.integers(1, 60))
        rle = ((rng.integers(1, 64, size=n) << 2) | rng.integers(0, 4, size=n)).astype(np.uint8)
        buf = C.create_string_buffer(2048)
        w = olib.rsa_oracle_cigar_text(rle.ctypes.data, n, C.cast(buf, C.c_void_p), 2048)
        assert ext.rle_to_text(rle, n) == buf.raw[:w].decode()
    assert ext.rle_to_text(np.array([(63 << 2) | 0, (5 << 2) | 0, (1 << 2) | 1], np.uint8), 3) == "1X68M"


def test_planner_routes_by_shape():
    b = W.extension_pairs(500, seed=3)
    p = ext.plan_debug(b.qoff, b.toff)
    assert p["pairs"] == 500 and p["fast_pairs"] == 500 and p["exact_pairs"] == 0 and p["fast_classes"] == 1
    assert p["groups"] % 4 == 0 and p["groups"] >= 250
    p = ext.plan_debug(b.qoff, b.toff, exact_only=True)
    assert p["fast_pairs"] == 0 and p["exact_pairs"] == 500
    # tiny queries go to the exact kernel, 257..512-base queries to the 16-lane packed kernel; empty strings
    # are failed records
    q = [b"ACGT", b"A" * 300, b"", b"ACGTACGTAC"]
    t = [b"ACGTTT", b"A" * 400, b"ACGT", b""]
    bb = W.from_lists(q, t)
    p = ext.plan_debug(bb.qoff, bb.toff)
    assert p["exact_pairs"] == 1 and p["failed"] == 2 and p["fast_pairs"] == 1


def test_planner_chunks_by_scratch_budget():
    b = W.extension_pairs(300, seed=4)
    full = ext.plan_debug(b.qoff, b.toff)
    small = ext.plan_debug(b.qoff, b.toff, scratch_cap=full["scratch_bytes"] // 3)
    assert 0 < small["pairs"] < 300 and small["scratch_bytes"] <= full["scratch_bytes"] // 3 + 65536


def test_device_planner_host_pass_agrees_with_host_planner():
    """The host pass of the device planner (chunk cut, routing counts, group slots) against the full host planner on the
    same batch; its scratch bound must cover what the host planner lays out and stay within a few percent of it."""
    b = W.extension_pairs(3000, seed=6, fixed_query_len=False, indel_rate=0.01)
    hp = ext.plan_debug(b.qoff, b.toff)
    sp = ext.scan_debug(b.qoff, b.toff)
    for k in ("pairs", "fast_pairs", "exact_pairs", "failed", "fast_classes"):
        assert sp[k] == hp[k], k
    assert sp["group_slots"] == hp["groups"]
    assert hp["scratch_bytes"] <= sp["scratch_bound"] <= hp["scratch_bytes"] * 1.15 + (1 << 20)
    q = [b"ACGT", b"A" * 300, b"", b"ACGTACGTAC", b"ACGT" * 30]
    t = [b"ACGTTT", b"A" * 400, b"ACGT", b"", b"ACGT" * 600]
    bb = W.from_lists(q, t)
    sp = ext.scan_debug(bb.qoff, bb.toff)
    assert sp["exact_pairs"] == 1 and sp["failed"] == 3 and sp["fast_pairs"] == 1
    # chunk cut by the scratch budget
    small = ext.scan_debug(b.qoff, b.toff, scratch_cap=hp["scratch_bytes"] // 3)
    assert 0 < small["pairs"] < 3000 and small["scratch_bound"] <= hp["scratch_bytes"] // 3


def test_device_planner_host_pass_is_cheap():
    b = W.extension_pairs_fast(131072, seed=8)
    sp = ext.scan_debug(b.qoff, b.toff, time_reps=20)
    hp = ext.plan_debug(b.qoff, b.toff, time_reps=5)
    assert sp["pairs"] == 131072
    # (a timing on a shared container: the margin is wide on purpose; typical ratio 0.25-0.45)
    assert sp["scan_ns"] < hp["plan_ns"] * 0.8, (sp["scan_ns"], hp["plan_ns"])
    print("host pass ns/pair", sp["scan_ns"] / 131072, "host planner ns/pair", hp["plan_ns"] / 131072)
