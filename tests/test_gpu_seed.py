"""GPU suite: the seeding kernels through the C ABI (include/rsa_seed.h: rsa_seed_index_upload / rsa_seed_find_nams)
against the reference's own seeding path compiled from /root/reference (oracle/_ref/libseed_ref.so), on the reference's
own index arrays: every NAM field, the NAM order, the nonrepetitive fraction and the rescue decision, bit-exact
(reference src/randstrobes.cpp:207, src/nam.cpp:771,955)."""
import threading

import numpy as np
import pytest

import oracle
import seed_util as U
from rabbitsalign_b200 import seed as S, workload as W

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(oracle.seed_reference_lib() is None, reason="needs oracle/_ref/libseed_ref.so")]


def _gpu_seed(idx, buf, off, rescue_level=2):
    gi = S.SeedIndexGpu(S.make_config(idx.params(), rescue_level=rescue_level), idx.randstrobes, idx.starts)
    sd = S.Seeder(gi)
    per, nams = sd.find_nams(buf, off)
    st = sd.stats()
    sd.close()
    gi.close()
    return per, nams, st


@pytest.mark.parametrize("name", sorted(U.CASES))
def test_gpu_seeding_equals_reference(name):
    idx, buf, off = U.make_case(name)
    per, nams, st = _gpu_seed(idx, buf, off)
    cnt, resc = U.assert_equals_reference(idx, buf, off, per, nams)
    assert st["kernel_launches"] >= 1 and st["reads"] == len(off) - 1 and st["nams"] == cnt.sum()
    assert st["reads_failed"] == 0 and st["reads_rescued"] == resc.sum()
    if name in ("r150_repeats", "rescue_heavy"):
        assert st["reads_retried"] > 0  # reads from repeats need the large scratch tier
    idx.close()


def test_gpu_seeding_120k_reads_with_repeats_and_rescue():
    """>= 100 k reads incl. repeats and rescue mode (VERDICT r1 task 2)."""
    contigs = W.seeding_genome(n_contigs=4, contig_len=2_000_000, seed=31, repeat_families=6, copies_per_contig=25)
    idx = oracle.build_seed_index(contigs, 150, 8)
    buf, off = W.seeding_reads(contigs, 120_000, seed=32, n_rate=0.0005, junk_frac=0.02)
    per, nams, st = _gpu_seed(idx, buf, off)
    cnt, resc = U.assert_equals_reference(idx, buf, off, per, nams)
    assert resc.sum() > 500 and cnt.max() > 100
    idx.close()


def test_rescue_level_one_on_gpu():
    idx, buf, off = U.make_case("rescue_heavy")
    per, nams, st = _gpu_seed(idx, buf, off, rescue_level=1)
    U.assert_equals_reference(idx, buf, off, per, nams, rescue_level=1)
    assert st["reads_rescued"] == 0
    idx.close()


def test_workers_share_one_index_and_results_are_deterministic():
    """Several host workers, one handle each, one index per GPU (north_star: index replicated per GPU, not per worker)."""
    idx, buf, off = U.make_case("r150_repeats")
    gi = S.SeedIndexGpu(S.make_config(idx.params()), idx.randstrobes, idx.starts)
    n = len(off) - 1
    cuts = [0, n // 3, 2 * n // 3, n]
    out = [None] * 3

    def work(k):
        sd = S.Seeder(gi)
        lo, hi = cuts[k], cuts[k + 1]
        o = np.ascontiguousarray(off[lo:hi + 1])
        for _ in range(3):
            per, nams = sd.find_nams(buf, o)  # offsets not starting at 0: the slice of a larger buffer
        out[k] = (per, nams)
        sd.close()
    ths = [threading.Thread(target=work, args=(k,)) for k in range(3)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    for k in range(3):
        lo, hi = cuts[k], cuts[k + 1]
        sub = buf[off[lo]:off[hi]]
        o = off[lo:hi + 1] - off[lo]
        U.assert_equals_reference(idx, np.ascontiguousarray(sub), np.ascontiguousarray(o), out[k][0], out[k][1])
    gi.close()
    idx.close()


def test_staged_run_matches_and_arguments_are_checked():
    idx, buf, off = U.make_case("r100")
    gi = S.SeedIndexGpu(S.make_config(idx.params()), idx.randstrobes, idx.starts)
    sd = S.Seeder(gi)
    per, nams = sd.find_nams(buf, off)
    sd.stage(buf, off)
    sd.run_staged()
    sd.run_staged()
    st = sd.stats()
    assert st["nams"] == len(nams) and st["kernel_ms"] > 0
    with pytest.raises(S.SeedError):
        sd.find_nams(buf, np.array([0, 70000], np.int64))  # a read longer than 65535
    per2, nams2 = sd.find_nams(buf, off)  # the handle stays usable
    assert per2["n_nams"].tolist() == per["n_nams"].tolist()
    sd.close()
    gi.close()
    with pytest.raises(S.SeedError):
        p = idx.params()
        p["bits"] = 5
        S.SeedIndexGpu(S.make_config(p), idx.randstrobes, idx.starts)
    idx.close()


def test_warp_per_read_tier_alone_equals_reference():
    """RSA_SEED_FORCE_LARGE=1 sends every read through the large tier (one warp per read, cooperative merge loops): same
    records.  Run in a subprocess because the switch is read once per process."""
    import os
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import seed_util as U\n"
        "from rabbitsalign_b200 import seed as S\n"
        "for name in ('r150_repeats', 'rescue_heavy', 'many_contigs', 'low_complexity', 'r50_short'):\n"
        "    idx, buf, off = U.make_case(name)\n"
        "    gi = S.SeedIndexGpu(S.make_config(idx.params()), idx.randstrobes, idx.starts)\n"
        "    sd = S.Seeder(gi)\n"
        "    per, nams = sd.find_nams(buf, off)\n"
        "    st = sd.stats()\n"
        "    U.assert_equals_reference(idx, buf, off, per, nams)\n"
        "    assert st['reads_retried'] > 0.5 * st['reads'], st\n"
        "    sd.close(); gi.close(); idx.close()\n"
        "print('ok')\n") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RSA_SEED_FORCE_LARGE="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
