"""oracle -- TEST INFRASTRUCTURE ONLY (see oracle/sw_oracle.c header).

ctypes front-ends for
  * liboracle.so          the C restatement of the reference's GPU extension path, and
  * _ref/libgasal_ref*.so the reference's own kernel headers compiled for the host
                          (oracle/ref_shim.cpp; built in the dev container from /root/reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
package.  The product package `rabbitsalign_b200` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


@dataclass
class PairResult:
    """Mirror of `struct gasal_tmp_res` (reference src/gasal2_ssw.h:31-38)."""

    score: int
    query_start: int
    query_end: int
    ref_start: int
    ref_end: int
    cigar_str: str

    def astuple(self):
        return (self.score, self.query_start, self.query_end, self.ref_start, self.ref_end, self.cigar_str)


def build(force: bool = False) -> None:
    """Compile the checkers (oracle/Makefile).  _ref/ targets are only (re)built when
    /root/reference is present; on the GPU box the prebuilt files that travelled are used."""
    if force or not os.path.exists(os.path.join(_HERE, "liboracle.so")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def pack_strings(seqs: Sequence[bytes]) -> Tuple[np.ndarray, np.ndarray]:
    """Concatenate byte strings -> (uint8 buffer, int64 offsets[n+1])."""
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    if len(seqs):
        off[1:] = np.cumsum([len(s) for s in seqs])
    buf = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if len(seqs) else np.zeros(0, np.uint8)
    if buf.size == 0:
        buf = np.zeros(1, np.uint8)
    return buf, off


_BATCH_ARGS = [
    C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
    C.c_void_p,
]


class _BatchLib:
    def __init__(self, path: str, symbol: str):
        self.path = path
        self.lib = C.CDLL(path)
        self.fn = getattr(self.lib, symbol)
        self.fn.argtypes = _BATCH_ARGS
        self.fn.restype = C.c_int

    def align_packed(self, qbuf, qoff, tbuf, toff, match=2, mismatch=8, gap_open=12, gap_extend=1):
        """Same scoring convention as solve_ssw_on_gpu (gasal2_ssw.cpp:52-56): gap_open is the
        strobealign value; GASAL gets gap_open-1."""
        n = len(qoff) - 1
        score = np.zeros(n, np.int32); qs = np.zeros(n, np.int32); qe = np.zeros(n, np.int32)
        rs = np.zeros(n, np.int32); re = np.zeros(n, np.int32); nops = np.zeros(n, np.int32)
        cap = int(4 * (qoff[-1] + toff[-1]) + 64 * n + 64)
        pool = np.zeros(cap, np.uint8)
        coff = np.zeros(n + 1, np.int64)
        rc = self.fn(n, qbuf.ctypes.data, qoff.ctypes.data, tbuf.ctypes.data, toff.ctypes.data,
                     match, mismatch, gap_open - 1, gap_extend, score.ctypes.data, qs.ctypes.data,
                     qe.ctypes.data, rs.ctypes.data, re.ctypes.data, nops.ctypes.data,
                     pool.ctypes.data, cap, coff.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"{self.path}: batch call failed rc={rc}")
        return score, qs, qe, rs, re, nops, pool, coff

    def align(self, queries: Sequence[bytes], targets: Sequence[bytes], **kw) -> List[PairResult]:
        qbuf, qoff = pack_strings(queries)
        tbuf, toff = pack_strings(targets)
        score, qs, qe, rs, re, nops, pool, coff = self.align_packed(qbuf, qoff, tbuf, toff, **kw)
        raw = pool.tobytes()
        return [PairResult(int(score[i]), int(qs[i]), int(qe[i]), int(rs[i]), int(re[i]),
                           raw[coff[i]:coff[i + 1]].decode()) for i in range(len(queries))]


_cache = {}


def restatement() -> _BatchLib:
    """The C restatement (sw_oracle.c)."""
    if "o" not in _cache:
        p = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(p):
            build()
        _cache["o"] = _BatchLib(p, "rsa_oracle_batch")
    return _cache["o"]


def reference(max_query_len: int = 500) -> Optional[_BatchLib]:
    """The reference's own kernels compiled for the host, or None when _ref/ was not built."""
    name = "libgasal_ref.so" if max_query_len == 500 else "libgasal_ref512.so"
    if name not in _cache:
        p = os.path.join(_HERE, "_ref", name)
        _cache[name] = _BatchLib(p, "gasal_ref_batch") if os.path.exists(p) else None
    return _cache[name]


def gasal_fail(query: bytes, target: bytes, r: PairResult) -> bool:
    """reference src/pc.cpp:466-478."""
    lib = restatement().lib
    lib.rsa_oracle_gasal_fail.argtypes = [C.c_int] * 7 + [C.c_char_p]
    lib.rsa_oracle_gasal_fail.restype = C.c_int
    return bool(lib.rsa_oracle_gasal_fail(len(query), len(target), r.score, r.query_start, r.query_end,
                                          r.ref_start, r.ref_end, r.cigar_str.encode()))


class SswReference:
    """The reference's CPU extension path (Aligner::align, reference src/aligner.cpp:114-210) compiled from
    /root/reference into oracle/_ref/libssw_ref_*.so.  Timing baseline for bench.py; None if not built."""

    def __init__(self, path: str):
        self.path = path
        self.lib = C.CDLL(path)
        vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
        self.lib.ssw_ref_align_batch.argtypes = [i64, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32,
                                                 vp, vp, vp, vp, vp, vp, vp, i32]
        self.lib.ssw_ref_align_batch.restype = i32
        self.lib.ssw_ref_align_gpu_batch.argtypes = [i64, vp, vp, vp, vp, i32, i32, i32, i32, i32,
                                                     vp, vp, vp, vp, vp, vp, vp,
                                                     vp, vp, vp, vp, vp, vp, vp, i32]
        self.lib.ssw_ref_align_gpu_batch.restype = i32

    def align_packed(self, qbuf, qoff, tbuf, toff, threads=1, match=2, mismatch=8, gap_open=12, gap_extend=1,
                     end_bonus=10, want_cigar=False):
        n = len(qoff) - 1
        out = {k: np.zeros(n, np.int32) for k in ("score", "qs", "qe", "rs", "re", "ed")}
        slot = 1024 if want_cigar else 0
        pool = np.zeros(n * slot, np.uint8) if want_cigar else None
        rc = self.lib.ssw_ref_align_batch(n, qbuf.ctypes.data, qoff.ctypes.data, tbuf.ctypes.data, toff.ctypes.data,
                                          match, mismatch, gap_open, gap_extend, end_bonus, threads,
                                          out["score"].ctypes.data, out["qs"].ctypes.data, out["qe"].ctypes.data,
                                          out["rs"].ctypes.data, out["re"].ctypes.data, out["ed"].ctypes.data,
                                          pool.ctypes.data if want_cigar else None, slot)
        if rc != 0:
            raise RuntimeError("ssw_ref_align_batch failed")
        if want_cigar:
            raw = pool.tobytes()
            out["cigar"] = [raw[i * slot:(i + 1) * slot].split(b"\0", 1)[0].decode() for i in range(n)]
        return out

    def align_gpu_packed(self, qbuf, qoff, tbuf, toff, g, match=2, mismatch=8, gap_open=12, gap_extend=1,
                         end_bonus=10):
        """Aligner::align_gpu (reference src/aligner.cpp:13-112) on GPU-path records `g` (dict with int32
        arrays score/qs/qe/rs/re and a list of CIGAR strings)."""
        n = len(qoff) - 1
        gp, go = pack_strings([c.encode() for c in g["cigar"]])
        out = {k: np.zeros(n, np.int32) for k in ("score", "qs", "qe", "rs", "re", "ed")}
        slot = 1024
        pool = np.zeros(n * slot, np.uint8)
        a = {k: np.ascontiguousarray(g[k], dtype=np.int32) for k in ("score", "qs", "qe", "rs", "re")}
        rc = self.lib.ssw_ref_align_gpu_batch(n, qbuf.ctypes.data, qoff.ctypes.data, tbuf.ctypes.data,
                                              toff.ctypes.data, match, mismatch, gap_open, gap_extend, end_bonus,
                                              a["score"].ctypes.data, a["qs"].ctypes.data, a["qe"].ctypes.data,
                                              a["rs"].ctypes.data, a["re"].ctypes.data, gp.ctypes.data, go.ctypes.data,
                                              out["score"].ctypes.data, out["qs"].ctypes.data, out["qe"].ctypes.data,
                                              out["rs"].ctypes.data, out["re"].ctypes.data, out["ed"].ctypes.data,
                                              pool.ctypes.data, slot)
        if rc != 0:
            raise RuntimeError("ssw_ref_align_gpu_batch failed")
        raw = pool.tobytes()
        out["cigar"] = [raw[i * slot:(i + 1) * slot].split(b"\0", 1)[0].decode() for i in range(n)]
        return out


def hamming_restatement(qbuf, qoff, tbuf, toff, match=2, mismatch=8, end_bonus=10):
    """The C restatement of the Hamming shortcut (sw_oracle.c: rsa_oracle_hamming; reference src/aln.cpp:391-404,
    src/aligner.cpp:219-302).  Returns dict of arrays hamming/status/score/ed/start/end and the CIGAR texts."""
    lib = restatement().lib
    lib.rsa_oracle_hamming.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_char_p, C.c_int]
    lib.rsa_oracle_hamming.restype = C.c_int
    n = len(qoff) - 1
    out = {k: np.zeros(n, np.int32) for k in ("hamming", "status", "score", "ed", "start", "end")}
    out["cigar"] = []
    buf = C.create_string_buffer(8192)
    one = [C.c_int() for _ in range(5)]
    for i in range(n):
        ql, tl = int(qoff[i + 1] - qoff[i]), int(toff[i + 1] - toff[i])
        hd = lib.rsa_oracle_hamming(qbuf.ctypes.data + int(qoff[i]), ql, tbuf.ctypes.data + int(toff[i]), tl, match, mismatch,
                                    end_bonus, *[C.addressof(v) for v in one], buf, 8192)
        out["hamming"][i] = hd
        for k, v in zip(("status", "score", "ed", "start", "end"), one):
            out[k][i] = v.value
        out["cigar"].append(buf.value.decode())
    return out


def hamming_reference(qbuf, qoff, tbuf, toff, match=2, mismatch=8, end_bonus=10):
    """The reference's own hamming_distance / hamming_align (compiled into oracle/_ref/libssw_ref_*.so), called as
    extend_seed_part calls them; None when _ref/ was not built."""
    ref = ssw_reference()
    if ref is None or not hasattr(ref.lib, "ssw_ref_hamming_batch"):
        return None
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    ref.lib.ssw_ref_hamming_batch.argtypes = [i64, vp, vp, vp, vp, i32, i32, i32] + [vp] * 9 + [i32]
    ref.lib.ssw_ref_hamming_batch.restype = i32
    n = len(qoff) - 1
    out = {k: np.zeros(n, np.int32) for k in ("hamming", "status", "score", "qs", "qe", "rs", "re", "ed")}
    slot = 2048
    pool = np.zeros(n * slot, np.uint8)
    rc = ref.lib.ssw_ref_hamming_batch(n, qbuf.ctypes.data, qoff.ctypes.data, tbuf.ctypes.data, toff.ctypes.data, match,
                                       mismatch, end_bonus, out["hamming"].ctypes.data, out["status"].ctypes.data,
                                       out["score"].ctypes.data, out["qs"].ctypes.data, out["qe"].ctypes.data,
                                       out["rs"].ctypes.data, out["re"].ctypes.data, out["ed"].ctypes.data,
                                       pool.ctypes.data, slot)
    if rc != 0:
        raise RuntimeError("ssw_ref_hamming_batch failed")
    raw = pool.tobytes()
    out["cigar"] = [raw[i * slot:(i + 1) * slot].split(b"\0", 1)[0].decode() for i in range(n)]
    return out


SAM_CALL_DTYPE = np.dtype([
    ("kind", "<i4"), ("is_primary", "<i4"), ("is_proper", "<i4"), ("mapq1", "<u4"), ("mapq2", "<u4"), ("unmapped_flags", "<u4"),
    ("a1", [("ref_id", "<i4"), ("ref_start", "<i4"), ("edit_distance", "<i4"), ("score", "<i4"), ("length", "<i4"),
            ("is_rc", "<i4"), ("is_unaligned", "<i4"), ("cigar_off", "<u4"), ("n_cigar", "<u4")]),
    ("a2", [("ref_id", "<i4"), ("ref_start", "<i4"), ("edit_distance", "<i4"), ("score", "<i4"), ("length", "<i4"),
            ("is_rc", "<i4"), ("is_unaligned", "<i4"), ("cigar_off", "<u4"), ("n_cigar", "<u4")]),
    ("r1", np.dtype([("name_off", "<u8"), ("seq_off", "<u8"), ("qual_off", "<u8"), ("name_len", "<u4"), ("seq_len", "<u4"),
                     ("qual_len", "<u4")], align=True)),
    ("r2", np.dtype([("name_off", "<u8"), ("seq_off", "<u8"), ("qual_off", "<u8"), ("name_len", "<u4"), ("seq_len", "<u4"),
                     ("qual_len", "<u4")], align=True)),
    ("details1", "<u4", (5,)), ("details2", "<u4", (5,))], align=True)
assert SAM_CALL_DTYPE.itemsize == 216


def sam_reference_replay(ref_names, calls, text_pool, cigar_pool, cigar_m=False, read_group=b"", output_unmapped=True,
                         show_details=False):
    """The reference's own SAM writer (class Sam, src/sam.cpp, compiled into oracle/_ref/libsam_ref.so) replaying a list of
    Sam::add / add_pair / add_unmapped / add_unmapped_pair calls (SAM_CALL_DTYPE); returns the text, or None when
    _ref/ was not built."""
    p = os.path.join(_HERE, "_ref", "libsam_ref.so")
    if not os.path.exists(p):
        return None
    if "sam" not in _cache:
        lib = C.CDLL(p)
        vp, i64, i32 = C.c_void_p, C.c_longlong, C.c_int
        lib.sam_ref_replay.argtypes = [i32, vp, vp, i32, C.c_char_p, i32, i32, i64, vp, vp, vp, vp, i64]
        lib.sam_ref_replay.restype = i64
        _cache["sam"] = lib
    lib = _cache["sam"]
    buf = b"".join(ref_names)
    off = np.zeros(len(ref_names) + 1, np.int64)
    off[1:] = np.cumsum([len(x) for x in ref_names])
    nb = np.frombuffer(buf, np.uint8).copy() if buf else np.zeros(1, np.uint8)
    calls = np.ascontiguousarray(calls, dtype=SAM_CALL_DTYPE)
    cigar_pool = np.ascontiguousarray(cigar_pool, dtype=np.uint32)
    cap = len(text_pool) * 2 + 400 * len(calls) + 16 * len(cigar_pool) + 1024
    out = np.zeros(cap, np.uint8)
    n = lib.sam_ref_replay(len(ref_names), nb.ctypes.data, off.ctypes.data, int(cigar_m), read_group, int(output_unmapped),
                           int(show_details), len(calls), calls.ctypes.data, text_pool.ctypes.data,
                           cigar_pool.ctypes.data if len(cigar_pool) else None, out.ctypes.data, cap)
    if n < 0:
        raise RuntimeError("sam_ref_replay: output buffer too small")
    return out[:n].tobytes()


def ssw_reference() -> Optional[SswReference]:
    if "ssw" not in _cache:
        flags = ""
        try:
            flags = open("/proc/cpuinfo").read()
        except OSError:
            pass
        v = "v3" if (" avx2" in flags and " bmi2" in flags and " fma" in flags) else "v2"
        p = os.path.join(_HERE, "_ref", f"libssw_ref_{v}.so")
        _cache["ssw"] = SswReference(p) if os.path.exists(p) else None
    return _cache["ssw"]


class GasalGpuReference:
    """The reference's own GPU path (GASAL2 + solve_ssw_on_gpu, compiled for sm_100a into _ref/libgasal_gpu.so):
    the strongest checker there is -- the unmodified reference CUDA kernels on the same GPU -- and the GPU
    comparator.  Needs a GPU; like the reference it exit()s the process on CUDA errors and on queries > 500."""

    def __init__(self, path: str):
        self.lib = C.CDLL(path)
        self.lib.gasal_gpu_batch.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_int]
        self.lib.gasal_gpu_batch.restype = C.c_int
        self.slice_size = int(self.lib.gasal_gpu_slice_size())

    def batch(self, qbuf, qoff, tbuf, toff, thread_id: int = 0, cigar_stride: int = 0):
        """-> (int32 array n x 5: score, query_start, query_end, ref_start, ref_end; list of CIGAR strings or None)"""
        n = len(qoff) - 1
        out5 = np.zeros((n, 5), dtype=np.int32)
        cig = np.zeros(n * cigar_stride, dtype=np.uint8) if cigar_stride else None
        qb = np.ascontiguousarray(qbuf, dtype=np.uint8)
        tb = np.ascontiguousarray(tbuf, dtype=np.uint8)
        qo = np.ascontiguousarray(qoff, dtype=np.int64)
        to = np.ascontiguousarray(toff, dtype=np.int64)
        self.lib.gasal_gpu_batch(thread_id, n, qb.ctypes.data, qo.ctypes.data, tb.ctypes.data, to.ctypes.data,
                                 out5.ctypes.data, cig.ctypes.data if cig is not None else None, cigar_stride)
        texts = None
        if cig is not None:
            raw = cig.reshape(n, cigar_stride)
            texts = [bytes(r).split(b"\0", 1)[0].decode() for r in raw]
        return out5, texts


def reference_gpu() -> Optional[GasalGpuReference]:
    if "gasal_gpu" not in _cache:
        p = os.path.join(_HERE, "_ref", "libgasal_gpu.so")
        _cache["gasal_gpu"] = GasalGpuReference(p) if os.path.exists(p) else None
    return _cache["gasal_gpu"]


NAM_DTYPE = np.dtype([("query_start", "<i4"), ("query_end", "<i4"), ("query_prev_hit_startpos", "<i4"),
                      ("ref_start", "<i4"), ("ref_end", "<i4"), ("ref_prev_hit_startpos", "<i4"),
                      ("n_hits", "<i4"), ("ref_id", "<i4"), ("score", "<f4"), ("is_rc", "<i4")])


class SeedIndex:
    """An index built by the reference's own StrobemerIndex::populate (src/index.cpp:141), with views of its arrays."""

    def __init__(self, lib, handle, keep):
        self.lib, self.h, self._keep = lib, handle, keep
        rs, n, st, ns, q = C.c_void_p(), C.c_int64(), C.c_void_p(), C.c_int64(), C.c_uint64()
        ints = (C.c_int32 * 8)()
        lib.seedref_export(handle, C.byref(rs), C.byref(n), C.byref(st), C.byref(ns), ints, C.byref(q))
        self.randstrobes = np.ctypeslib.as_array(C.cast(rs, C.POINTER(C.c_uint8)), shape=(n.value * 16,)) if n.value else np.zeros(0, np.uint8)
        self.n_randstrobes = n.value
        self.starts = np.ctypeslib.as_array(C.cast(st, C.POINTER(C.c_uint64)), shape=(ns.value,))
        (self.bits, self.filter_cutoff, self.k, self.s, self.t_syncmer, self.w_min, self.w_max, self.max_dist) = [int(x) for x in ints]
        self.q = int(q.value)

    def params(self) -> dict:
        return dict(bits=self.bits, filter_cutoff=self.filter_cutoff, k=self.k, s=self.s, t_syncmer=self.t_syncmer,
                    w_min=self.w_min, w_max=self.w_max, max_dist=self.max_dist, q=self.q)

    def rescue_cutoff(self, rescue_level: int = 2) -> int:
        return rescue_level * self.filter_cutoff if rescue_level < 100 else 1000  # src/main.cpp:415

    def find_nams(self, reads: np.ndarray, roff: np.ndarray, rescue_level: int = 2):
        """-> (nam_count[n], fraction[n], rescued[n], nams[total] as NAM_DTYPE) in the reference's pre-sort order."""
        n = len(roff) - 1
        cnt = np.zeros(n, np.int32); frac = np.zeros(n, np.float32); resc = np.zeros(n, np.uint8)
        rc = self.rescue_cutoff(rescue_level)
        total = self.lib.seedref_find_nams(self.h, reads.ctypes.data, roff.ctypes.data, n, rescue_level, rc,
                                           cnt.ctypes.data, frac.ctypes.data, resc.ctypes.data, None, 0)
        nams = np.zeros(max(1, total), NAM_DTYPE)
        got = self.lib.seedref_find_nams(self.h, reads.ctypes.data, roff.ctypes.data, n, rescue_level, rc,
                                         cnt.ctypes.data, frac.ctypes.data, resc.ctypes.data, nams.ctypes.data, total)
        assert got == total
        return cnt, frac, resc, nams[:total]

    def time_find_nams(self, reads, roff, threads: int, rescue_level: int = 2) -> int:
        return int(self.lib.seedref_time(self.h, reads.ctypes.data, roff.ctypes.data, len(roff) - 1, rescue_level,
                                         self.rescue_cutoff(rescue_level), threads))

    def map_order(self, keys) -> np.ndarray:
        k = np.ascontiguousarray(keys, dtype=np.uint32)
        out = np.zeros(len(k), np.uint32)
        self.lib.seedref_map_order(k.ctypes.data, len(k), out.ctypes.data)
        return out

    def close(self):
        if self.h:
            self.lib.seedref_free(self.h)
            self.h = None


def seed_reference_lib():
    """The reference's seeding path (oracle/_ref/libseed_ref.so) or None when it was not built."""
    if "seed" not in _cache:
        p = os.path.join(_HERE, "_ref", "libseed_ref.so")
        lib = None
        if os.path.exists(p):
            lib = C.CDLL(p)
            lib.seedref_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
            lib.seedref_build.restype = C.c_void_p
            lib.seedref_free.argtypes = [C.c_void_p]
            lib.seedref_export.argtypes = [C.c_void_p] + [C.c_void_p] * 6
            lib.seedref_find_nams.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
            lib.seedref_find_nams.restype = C.c_int64
            lib.seedref_time.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int]
            lib.seedref_time.restype = C.c_int64
            lib.seedref_map_order.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _cache["seed"] = lib
    return _cache["seed"]


def build_seed_index(contigs: Sequence[np.ndarray], read_len: int = 150, threads: int = 8) -> Optional[SeedIndex]:
    lib = seed_reference_lib()
    if lib is None:
        return None
    concat = np.ascontiguousarray(np.concatenate(contigs), dtype=np.uint8)
    off = np.zeros(len(contigs) + 1, np.int64)
    off[1:] = np.cumsum([len(c) for c in contigs])
    h = lib.seedref_build(concat.ctypes.data, off.ctypes.data, len(contigs), read_len, threads)
    return SeedIndex(lib, h, (concat, off))
