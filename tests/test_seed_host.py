"""CPU suite: the seeding code the CUDA kernel runs (rabbitsalign_b200/csrc/kernels_seed.cuh), compiled for the host by
tests/seed_host_check.cu, against the reference's own seeding path compiled from /root/reference
(oracle/_ref/libseed_ref.so: randstrobes_query + find_nams + find_nams_rescue on the reference's own index).  Bit-exact:
NAM fields and order, nonrepetitive fraction, rescue decision.  The `-m gpu` twin is tests/test_gpu_seed.py."""
import shutil

import numpy as np
import pytest

import oracle
import seed_util as U
from rabbitsalign_b200 import seed as S

pytestmark = pytest.mark.skipif(oracle.seed_reference_lib() is None or shutil.which("nvcc") is None,
                                reason="needs oracle/_ref/libseed_ref.so (built from /root/reference) and nvcc")


@pytest.fixture(scope="module")
def harness():
    return U.host_harness()


@pytest.mark.parametrize("name", sorted(U.CASES))
def test_host_compiled_seeding_equals_reference(harness, name):
    idx, buf, off = U.make_case(name)
    per, nams, n_large = U.host_seed(harness, idx, buf, off)
    cnt, resc = U.assert_equals_reference(idx, buf, off, per, nams)
    assert cnt.sum() > 0
    if name == "rescue_heavy":
        assert resc.sum() > 100
    if name == "many_contigs":
        # several reference ids on one strand: the group order is the reference's hash-map order, not first touch
        multi = sum(1 for r in range(len(per)) if len(set(nams["ref_id"][per["nam_off"][r]:per["nam_off"][r] + per["n_nams"][r]])) > 2)
        assert multi > 20
    idx.close()


def test_rescue_level_one_never_rescues(harness):
    idx, buf, off = U.make_case("rescue_heavy")
    cfg = S.make_config(idx.params(), rescue_level=1)
    per, nams, _ = U.host_seed(harness, idx, buf, off, cfg)
    U.assert_equals_reference(idx, buf, off, per, nams, rescue_level=1)
    assert (per["flags"] & S.READ_RESCUED).sum() == 0
    idx.close()


def test_group_order_is_the_containers_not_first_touch():
    """The one order this library leaves to the binding: robin_hood's iteration order of the touched reference ids."""
    lib = oracle.seed_reference_lib()
    idx = oracle.SeedIndex.__new__(oracle.SeedIndex)
    idx.lib = lib
    assert list(idx.map_order([0, 1, 2])) != [0, 1, 2] or list(idx.map_order([5, 3, 9, 1])) != [5, 3, 9, 1]
    a = list(idx.map_order([7, 2, 11, 30, 4]))
    assert sorted(a) == [2, 4, 7, 11, 30]
