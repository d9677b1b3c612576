"""GPU suite, SURVEY 8f rank 4: rsa_sam_format (csrc/sam.cu) against the reference's own SAM writer (class Sam compiled
from /root/reference into oracle/_ref/libsam_ref.so) and the plain restatement, byte for byte, in the four writer
configurations (=/X or M CIGARs, read group + detail tags, unmapped records suppressed)."""
import numpy as np
import pytest

import oracle
from rabbitsalign_b200 import sam as S
from sam_format_util import VARIANTS, make_calls, records_from_calls, restate

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", range(len(VARIANTS)))
def test_device_sam_text_equals_reference_writer(variant):
    kw = VARIANTS[variant]
    ref_names, calls, text, cig = make_calls(4000, seed=500 + variant)
    want = oracle.sam_reference_replay(ref_names, calls, text, cig, **kw)
    if want is None:
        want = restate(ref_names, calls, text, cig, **kw)
    else:
        assert restate(ref_names, calls, text, cig, **kw) == want
    recs = records_from_calls(calls)
    f = S.SamFormatter(ref_names, **kw)
    got, offs = f.format(recs, text, cig, want_offsets=True)
    f.close()
    if got != want:
        gl, wl = got.split(b"\n"), want.split(b"\n")
        bad = [(i, a, b) for i, (a, b) in enumerate(zip(gl, wl)) if a != b][:3]
        raise AssertionError(f"{len(gl)} vs {len(wl)} lines; first differences: {bad}")
    assert offs[0] == 0 and offs[-1] == len(got) and (np.diff(offs) >= 0).all()
    starts = offs[:-1][np.diff(offs) > 0]
    assert all(got[int(s) - 1:int(s)] == b"\n" for s in starts[1:200])


def test_output_buffer_too_small_is_reported():
    ref_names, calls, text, cig = make_calls(50, seed=9)
    recs = records_from_calls(calls)
    f = S.SamFormatter(ref_names)
    import ctypes as C
    out = np.zeros(16, np.uint8)
    n = C.c_int64(0)
    rc = f.lib.rsa_sam_format(f.h, len(recs), recs.ctypes.data, text.ctypes.data, len(text), cig.ctypes.data, len(cig),
                              out.ctypes.data, 16, C.byref(n), None)
    assert rc == -1 and n.value > 16
    assert len(f.format(recs, text, cig)) == n.value   # the handle stays usable
    f.close()


def test_pinned_and_pageable_callers_get_the_same_text():
    """rsa_sam_format reads pinned pools in place and bounces pageable ones; the same for the output buffer."""
    import torch
    ref_names, calls, text, cig = make_calls(3000, seed=77)
    recs = records_from_calls(calls)
    f = S.SamFormatter(ref_names)
    want = f.format(recs, text, cig)                                             # all pageable

    def pinned(a):
        t = torch.empty(a.nbytes, dtype=torch.uint8).pin_memory()
        v = t.numpy().view(a.dtype).reshape(a.shape)
        v[...] = a
        return t, v
    keep = [pinned(np.ascontiguousarray(x)) for x in (recs, text, np.ascontiguousarray(cig, dtype=np.uint32))]
    tout = torch.empty(len(want) + 64, dtype=torch.uint8).pin_memory()
    got = f.format(keep[0][1], keep[1][1], keep[2][1], out=tout.numpy())         # all pinned
    assert bytes(got) == want
    assert f.format(keep[0][1], text, cig) == want                               # mixed
    f.close()
