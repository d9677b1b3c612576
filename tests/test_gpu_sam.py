"""GPU suite: byte-identical SAM.  The reference's UNMODIFIED host pipeline (seeding, NAMs, window
construction, gasal_fail gate, align_gpu, SAM writer -- compiled from /root/reference by
integration/build.sh) linked against the product (integration/gasal2_ssw.cpp -> librsa_ext.so) must write
the same SAM as the same pipeline linked against the reference's own GASAL2 kernels (golden md5s in
tests/golden/sam_golden.json, generated in the dev container by tests/golden/make_sam_golden.py).
The `alninfo` variants add the optional caller-loop edit (integration/patch_caller.py): gasal_fail and
Aligner::align_gpu are taken from the device's finish kernel; the SAM must still be the same bytes."""
import json
import os
import subprocess
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_sam_golden import B, CONFIGS, ROOT, md5_file  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "sam_golden.json")))


SUFFIX = {"record": "", "alninfo": "_alninfo", "big": "_big", "windows": "_win", "gpuseed": "_gpuseed", "gpuham": "_gpuham", "gpusam": "_gpusam", "two-gpus": "_gpusam"}


def _gpu_count():
    from rabbitsalign_b200 import ext
    return ext.load_library().rsa_ext_device_count()


@pytest.mark.parametrize("variant", ["record", "alninfo", "big", "windows", "gpuseed", "gpuham", "gpusam", "two-gpus", "reference-gpu"])
@pytest.mark.parametrize("name", sorted(GOLD))
def test_sam_is_byte_identical(name, variant, tmp_path):
    """record: unmodified caller, 512-pair slices.  alninfo: AlignmentInfo from the device.  big: unmodified caller
    compiled with -DSTREAM_BATCH_SIZE=1048576 (one call per chunk).  windows: the caller edit of SURVEY 8f rank 1
    (windows as offsets into the genome resident in HBM, one call per chunk).  gpuseed: the windows build plus seeding on
    the GPU (randstrobes, index lookup, NAM merge, rescue; integration/patch_seed.py).  gpuham: the gpuseed build plus the Hamming shortcut of
    extend_seed_part decided on the GPU for a whole chunk (integration/patch_hamming.py).  gpusam: the gpuham build plus the SAM text of
    every record written by the device formatter, one call per chunk (integration/patch_sam.py).  two-gpus: the
    gpusam build with its workers spread over two GPUs (genome and index replicated per GPU); skipped on a one-GPU box."""
    g = GOLD[name]
    BIN = os.path.join(B, ("rabbitsalign_fx_b200" if g.get("fx") else "rabbitsalign_b200") + SUFFIX.get(variant, ""))
    env = dict(os.environ)
    if variant == "two-gpus":
        if _gpu_count() < 2:
            pytest.skip("needs two GPUs")
        if g["threads"] < 2:
            pytest.skip("one worker cannot span two GPUs")
        env["RSA_EXT_DEVICES"] = "2"
    else:
        env["RSA_EXT_DEVICES"] = "1"
    if variant == "reference-gpu":
        # the reference as shipped (its own GASAL2 GPU path, sm_100a build): the golden md5s came from its kernels
        # compiled for the HOST; this confirms them against the real GPU run
        if g.get("fx"):
            pytest.skip("comparator is built in the plain-reader configuration only")
        BIN = os.path.join(B, "rabbitsalign_gasalgpu")
    if not os.path.exists(BIN):
        pytest.skip(f"{BIN} not built (needs /root/reference at build time)")
    d = str(tmp_path / name)
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_reads.py"), d] + g["make_reads_args"])
    files = sorted(g["inputs"])  # ref.fa, reads_1.fq[, reads_2.fq]
    for f in files:
        assert md5_file(os.path.join(d, f)) == g["inputs"][f], f"synthetic input {f} differs from the golden run"
    out = os.path.join(d, "b200.sam")
    args = [os.path.join(d, "ref.fa"), os.path.join(d, "reads_1.fq")] + ([os.path.join(d, "reads_2.fq")] if g["paired"] else [])
    r = subprocess.run([BIN, "-t", str(g["threads"]), "-o", out] + args, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    n = sum(1 for ln in open(out, "rb") if not ln.startswith(b"@"))
    assert n == g["records"]
    assert md5_file(out, skip_pg=True) == g["sam_md5_gasal_semantics"]
