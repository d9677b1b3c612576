"""ctypes binding of librsa_ext.so (include/rsa_ext.h) and the host-side mirror of the reference's
batch interface for the extension step.

`ExtensionEngine.solve_ssw_on_gpu(queries, refs, match, mismatch, gap_open, gap_extend)` has the argument
meaning and the result fields of the reference's

    void solve_ssw_on_gpu(int thread_id, std::vector<gasal_tmp_res>&, std::vector<std::string>& querys,
                          std::vector<std::string>& refs, int match, int mismatch, int gap_open,
                          int gap_extend)                               (reference src/gasal2_ssw.h:31-47)

There is no CPU fallback: if the shared library is missing or no CUDA device is usable, constructing an
engine raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# RSA_EXT_LIB: an A/B build of the same library (kernel experiments); the product is the in-tree librsa_ext.so
LIB_PATH = os.environ.get("RSA_EXT_LIB") or os.path.join(_HERE, "librsa_ext.so")

RLE_INLINE = 40
FLAG_EXACT_ONLY = 1
FLAG_SERIALIZE = 2
FLAG_HOST_PLAN = 4
FLAG_ASCII_WINDOWS = 8

RESULT_DTYPE = np.dtype([
    ("score", "<i4"), ("query_start", "<i4"), ("query_end", "<i4"), ("ref_start", "<i4"),
    ("ref_end", "<i4"), ("n_ops", "<i2"), ("status", "<i2"), ("rle", "u1", (RLE_INLINE,)),
])
assert RESULT_DTYPE.itemsize == 64

CIGAR_INLINE = 25
ALNINFO_DTYPE = np.dtype([
    ("sw_score", "<i4"), ("edit_distance", "<i4"), ("ref_start", "<i4"), ("ref_end", "<i4"),
    ("query_start", "<i4"), ("query_end", "<i4"), ("n_cigar", "<i2"), ("status", "<i2"),
    ("cigar", "<u4", (CIGAR_INLINE,)),
])
assert ALNINFO_DTYPE.itemsize == 128


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_query_len", C.c_int32), ("max_target_len", C.c_int32),
                ("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open", C.c_int32),
                ("gap_extend", C.c_int32), ("flags", C.c_int32), ("scratch_bytes", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("pairs_fast", C.c_int64), ("pairs_exact", C.c_int64),
                ("pairs_failed", C.c_int64), ("cells", C.c_int64), ("h2d_bytes", C.c_int64),
                ("d2h_bytes", C.c_int64), ("dp_ms", C.c_double), ("tb_ms", C.c_double), ("pairs_redo", C.c_int64), ("host_plan_ms", C.c_double)]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/rsa_ext.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "rsa_ext_create", "rsa_ext_destroy", "rsa_ext_last_error", "rsa_ext_submit", "rsa_ext_submit_ptrs",
    "rsa_ext_poll", "rsa_ext_wait", "rsa_ext_rle_overflow", "rsa_ext_rle_to_text",
    "rsa_ext_stage_resident", "rsa_ext_run_resident", "rsa_ext_fetch_resident", "rsa_ext_stream",
    "rsa_ext_get_stats", "rsa_ext_version", "rsa_ext_device_count", "rsa_ext_request_alninfo", "rsa_ext_plan_debug",
    "rsa_ext_reserve", "rsa_ext_set_reference", "rsa_ext_submit_ref_windows", "rsa_ext_share_reference", "rsa_ext_scan_debug",
    "rsa_ext_hamming_align", "rsa_ext_hamming_ref_windows", "rsa_ext_packed_reference",
]

_lib = None


def load_library() -> C.CDLL:
    """Load librsa_ext.so; raises if it was not built (run `make` or __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build the CUDA extension first (make / "
                           "__graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    lib.rsa_ext_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.rsa_ext_create.restype = C.c_int
    lib.rsa_ext_destroy.argtypes = [vp]
    lib.rsa_ext_destroy.restype = None
    lib.rsa_ext_last_error.argtypes = [vp]
    lib.rsa_ext_last_error.restype = C.c_char_p
    lib.rsa_ext_submit.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    lib.rsa_ext_submit.restype = C.c_int
    lib.rsa_ext_submit_ptrs.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    lib.rsa_ext_submit_ptrs.restype = C.c_int
    lib.rsa_ext_poll.argtypes = [vp]
    lib.rsa_ext_poll.restype = C.c_int
    lib.rsa_ext_wait.argtypes = [vp]
    lib.rsa_ext_wait.restype = C.c_int
    lib.rsa_ext_rle_overflow.argtypes = [vp, i64, vp, i32]
    lib.rsa_ext_rle_overflow.restype = C.c_int
    lib.rsa_ext_rle_to_text.argtypes = [vp, i32, vp, i32]
    lib.rsa_ext_rle_to_text.restype = C.c_int
    lib.rsa_ext_stage_resident.argtypes = [vp, i64, vp, vp, vp, vp]
    lib.rsa_ext_stage_resident.restype = C.c_int
    lib.rsa_ext_run_resident.argtypes = [vp]
    lib.rsa_ext_run_resident.restype = C.c_int
    lib.rsa_ext_fetch_resident.argtypes = [vp, vp]
    lib.rsa_ext_fetch_resident.restype = C.c_int
    lib.rsa_ext_stream.argtypes = [vp]
    lib.rsa_ext_stream.restype = vp
    lib.rsa_ext_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.rsa_ext_get_stats.restype = C.c_int
    lib.rsa_ext_version.argtypes = []
    lib.rsa_ext_version.restype = C.c_int
    lib.rsa_ext_request_alninfo.argtypes = [vp, vp, i32]
    lib.rsa_ext_request_alninfo.restype = C.c_int
    lib.rsa_ext_plan_debug.argtypes = [i64, vp, vp, i64, C.c_int, vp]
    lib.rsa_ext_plan_debug.restype = C.c_int
    lib.rsa_ext_scan_debug.argtypes = [i64, vp, vp, i64, vp]
    lib.rsa_ext_scan_debug.restype = C.c_int
    lib.rsa_ext_reserve.argtypes = [vp, i64, i32, i32]
    lib.rsa_ext_reserve.restype = C.c_int
    lib.rsa_ext_set_reference.argtypes = [vp, vp, i64]
    lib.rsa_ext_set_reference.restype = C.c_int
    lib.rsa_ext_share_reference.argtypes = [vp, vp]
    lib.rsa_ext_share_reference.restype = C.c_int
    lib.rsa_ext_submit_ref_windows.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    lib.rsa_ext_submit_ref_windows.restype = C.c_int
    lib.rsa_ext_packed_reference.argtypes = [vp, vp, vp, i64]
    lib.rsa_ext_packed_reference.restype = C.c_int
    lib.rsa_ext_hamming_align.argtypes = [vp, i64, vp, vp, vp, vp, i32, vp, vp]
    lib.rsa_ext_hamming_align.restype = C.c_int
    lib.rsa_ext_hamming_ref_windows.argtypes = [vp, i64, vp, vp, vp, i32, vp, vp]
    lib.rsa_ext_hamming_ref_windows.restype = C.c_int
    _lib = lib
    return lib


@dataclass
class GasalTmpRes:
    """Field-for-field mirror of `struct gasal_tmp_res` (reference src/gasal2_ssw.h:31-38)."""

    score: int
    query_start: int
    query_end: int
    ref_start: int
    ref_end: int
    cigar_str: str

    def astuple(self):
        return (self.score, self.query_start, self.query_end, self.ref_start, self.ref_end, self.cigar_str)


class ExtensionError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"rsa_ext status {status}: {msg}")
        self.status = status


def rle_to_text(rle: np.ndarray, n_ops: int) -> str:
    """reference src/gasal2_ssw.cpp:184-243 (host CIGAR text), done by the library."""
    lib = load_library()
    rle = np.ascontiguousarray(rle, dtype=np.uint8)
    out = C.create_string_buffer(16 * max(1, int(n_ops)) + 16)
    w = lib.rsa_ext_rle_to_text(rle.ctypes.data, int(n_ops), C.cast(out, C.c_void_p), len(out))
    if w < 0:
        raise ExtensionError(w, "rle_to_text buffer too small")
    return out.raw[:w].decode()


class ExtensionEngine:
    """One handle = one host worker's GPU context (the reference's per-thread_id GASAL storage,
    src/gasal2_ssw.cpp:29,92-102)."""

    def __init__(self, device: int = 0, match: int = 2, mismatch: int = 8, gap_open: int = 12,
                 gap_extend: int = 1, max_query_len: int = 500, max_target_len: int = 2000,
                 exact_only: bool = False, scratch_bytes: int = 0, serialize: bool = False, host_plan: bool = False,
                 ascii_windows: bool = False):
        self.lib = load_library()
        self.cfg = Config(device, max_query_len, max_target_len, match, mismatch, gap_open, gap_extend,
                          (FLAG_EXACT_ONLY if exact_only else 0) | (FLAG_SERIALIZE if serialize else 0) |
                          (FLAG_HOST_PLAN if host_plan else 0) | (FLAG_ASCII_WINDOWS if ascii_windows else 0), scratch_bytes)
        h = C.c_void_p()
        rc = self.lib.rsa_ext_create(C.byref(self.cfg), C.byref(h))
        if rc != 0:
            raise ExtensionError(rc, self.lib.rsa_ext_last_error(None).decode())
        self.h = h
        self._keep = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.rsa_ext_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise ExtensionError(rc, self.lib.rsa_ext_last_error(self.h).decode())

    # ---- asynchronous C-ABI legs on packed buffers --------------------------------------------------
    def submit(self, qbuf: np.ndarray, qoff: np.ndarray, tbuf: np.ndarray, toff: np.ndarray,
               results: Optional[np.ndarray] = None) -> np.ndarray:
        n = len(qoff) - 1
        if results is None:
            results = np.zeros(n, dtype=RESULT_DTYPE)
        assert qbuf.dtype == np.uint8 and tbuf.dtype == np.uint8
        assert qoff.dtype == np.int64 and toff.dtype == np.int64
        self._keep = (qbuf, qoff, tbuf, toff, results)
        self._check(self.lib.rsa_ext_submit(self.h, n, qbuf.ctypes.data, qoff.ctypes.data, tbuf.ctypes.data,
                                            toff.ctypes.data, results.ctypes.data))
        return results

    def submit_raw(self, n, qbuf_ptr, qoff_ptr, tbuf_ptr, toff_ptr, res_ptr):
        self._check(self.lib.rsa_ext_submit(self.h, n, qbuf_ptr, qoff_ptr, tbuf_ptr, toff_ptr, res_ptr))

    def request_alninfo(self, out: Optional[np.ndarray], end_bonus: int = 10):
        """Also produce `AlignmentInfo` records (ALNINFO_DTYPE) on the device for the following submits."""
        self._aln_keep = out
        self._check(self.lib.rsa_ext_request_alninfo(self.h, out.ctypes.data if out is not None else None, end_bonus))

    def set_reference(self, seq: np.ndarray):
        """Upload the (concatenated) reference once; windows are then named by offset and length."""
        assert seq.dtype == np.uint8
        self._ref_keep = np.ascontiguousarray(seq)
        self._check(self.lib.rsa_ext_set_reference(self.h, self._ref_keep.ctypes.data, len(self._ref_keep)))

    def packed_reference(self, n_bases: int):
        """The packed planes of the resident reference as the DP kernel's staging reads them: (codes uint32[], 16 bases
        of 2 bits per word; flags uint32[], 32 "not ACGT" bits per word), whole 64-base units."""
        units = (n_bases + 63) // 64
        codes = np.zeros(units * 4, np.uint32)
        flags = np.zeros(units * 2, np.uint32)
        self._check(self.lib.rsa_ext_packed_reference(self.h, codes.ctypes.data, flags.ctypes.data, units))
        return codes, flags

    def share_reference(self, donor: "ExtensionEngine"):
        """Use the resident reference another engine on the same device uploaded (one copy in HBM)."""
        self._ref_keep = donor._ref_keep
        self._check(self.lib.rsa_ext_share_reference(self.h, donor.h))

    def align_ref_windows(self, qbuf, qoff, win_off, win_len) -> np.ndarray:
        """Blocking: queries against windows [win_off[i], win_off[i]+win_len[i]) of the resident reference."""
        n = len(qoff) - 1
        results = np.zeros(n, dtype=RESULT_DTYPE)
        wo = np.ascontiguousarray(win_off, dtype=np.int64)
        wl = np.ascontiguousarray(win_len, dtype=np.int32)
        self._keep = (qbuf, qoff, wo, wl, results)
        self._check(self.lib.rsa_ext_submit_ref_windows(self.h, n, qbuf.ctypes.data, qoff.ctypes.data, wo.ctypes.data,
                                                        wl.ctypes.data, results.ctypes.data))
        self.wait()
        return results

    def hamming_align(self, qbuf, qoff, tbuf, toff, end_bonus: int = 10):
        """The Hamming shortcut of extend_seed_part (reference src/aln.cpp:391-404, src/aligner.cpp:219-302) for n
        (read, equally long window) pairs: returns (hamming int32[n], ALNINFO records; status 0 = shortcut applies)."""
        n = len(qoff) - 1
        ham = np.zeros(n, np.int32)
        out = np.zeros(n, dtype=ALNINFO_DTYPE)
        self._check(self.lib.rsa_ext_hamming_align(self.h, n, qbuf.ctypes.data, qoff.ctypes.data, tbuf.ctypes.data,
                                                   toff.ctypes.data, end_bonus, ham.ctypes.data, out.ctypes.data))
        return ham, out

    def hamming_ref_windows(self, qbuf, qoff, win_off, end_bonus: int = 10):
        """hamming_align() with window i = |query i| bases of the resident reference starting at win_off[i]."""
        n = len(qoff) - 1
        ham = np.zeros(n, np.int32)
        out = np.zeros(n, dtype=ALNINFO_DTYPE)
        wo = np.ascontiguousarray(win_off, dtype=np.int64)
        self._check(self.lib.rsa_ext_hamming_ref_windows(self.h, n, qbuf.ctypes.data, qoff.ctypes.data, wo.ctypes.data,
                                                         end_bonus, ham.ctypes.data, out.ctypes.data))
        return ham, out

    def align_ptrs(self, queries: Sequence[bytes], targets: Sequence[bytes]) -> np.ndarray:
        """Blocking rsa_ext_submit_ptrs: n separate strings, the shape of std::vector<std::string> (the veneer's path)."""
        n = len(queries)
        results = np.zeros(n, dtype=RESULT_DTYPE)
        qb = [C.create_string_buffer(bytes(q), max(1, len(q))) for q in queries]
        tb = [C.create_string_buffer(bytes(t), max(1, len(t))) for t in targets]
        qp = (C.c_void_p * n)(*[C.addressof(x) for x in qb])
        tp = (C.c_void_p * n)(*[C.addressof(x) for x in tb])
        ql = np.array([len(q) for q in queries], dtype=np.int32)
        tl = np.array([len(t) for t in targets], dtype=np.int32)
        self._keep = (qb, tb, qp, tp, ql, tl, results)
        self._check(self.lib.rsa_ext_submit_ptrs(self.h, n, C.cast(qp, C.c_void_p), ql.ctypes.data,
                                                 C.cast(tp, C.c_void_p), tl.ctypes.data, results.ctypes.data))
        self.wait()
        return results

    def reserve(self, n: int, qlen: int, tlen: int):
        """Pre-allocate for batches of up to n pairs of (qlen x tlen)."""
        self._check(self.lib.rsa_ext_reserve(self.h, n, qlen, tlen))

    def poll(self) -> int:
        return self.lib.rsa_ext_poll(self.h)

    def wait(self):
        self._check(self.lib.rsa_ext_wait(self.h))

    def align_packed(self, qbuf, qoff, tbuf, toff) -> np.ndarray:
        res = self.submit(qbuf, qoff, tbuf, toff)
        self.wait()
        return res

    def full_rle(self, results: np.ndarray, i: int) -> np.ndarray:
        n_ops = int(results["n_ops"][i])
        if n_ops <= RLE_INLINE:
            return results["rle"][i][:n_ops]
        out = np.zeros(n_ops, np.uint8)
        w = self.lib.rsa_ext_rle_overflow(self.h, i, out.ctypes.data, n_ops)
        if w != n_ops:
            raise ExtensionError(w, self.lib.rsa_ext_last_error(self.h).decode())
        return out

    def cigar(self, results: np.ndarray, i: int) -> str:
        n_ops = int(results["n_ops"][i])
        if n_ops <= 0:
            return ""
        return rle_to_text(self.full_rle(results, i), n_ops)

    # ---- reference-shaped call --------------------------------------------------------------------
    def solve_ssw_on_gpu(self, querys: Sequence[bytes], refs: Sequence[bytes]) -> List[GasalTmpRes]:
        """Blocking batch call with the reference's semantics; scores were fixed at construction like
        the reference fixes them on the first call per thread_id (gasal2_ssw.cpp:49-57)."""
        if len(querys) != len(refs):
            raise ValueError("querys and refs differ in length")  # reference: assert (gasal2_ssw.cpp:31)
        from .workload import from_lists
        b = from_lists([bytes(q) for q in querys], [bytes(t) for t in refs])
        qbuf = b.qbuf if b.qbuf.size else np.zeros(1, np.uint8)
        tbuf = b.tbuf if b.tbuf.size else np.zeros(1, np.uint8)
        res = self.align_packed(qbuf, b.qoff, tbuf, b.toff)
        return [GasalTmpRes(int(res["score"][i]), int(res["query_start"][i]), int(res["query_end"][i]),
                            int(res["ref_start"][i]), int(res["ref_end"][i]), self.cigar(res, i))
                for i in range(len(querys))]

    # ---- device-resident legs -----------------------------------------------------------------------
    def stage_resident(self, qbuf, qoff, tbuf, toff):
        self._keep = (qbuf, qoff, tbuf, toff)
        self._check(self.lib.rsa_ext_stage_resident(self.h, len(qoff) - 1, qbuf.ctypes.data, qoff.ctypes.data,
                                                    tbuf.ctypes.data, toff.ctypes.data))

    def run_resident(self):
        self._check(self.lib.rsa_ext_run_resident(self.h))

    def fetch_resident(self, n: int) -> np.ndarray:
        res = np.zeros(n, dtype=RESULT_DTYPE)
        self._check(self.lib.rsa_ext_fetch_resident(self.h, res.ctypes.data))
        return res

    @property
    def stream(self) -> int:
        return int(self.lib.rsa_ext_stream(self.h) or 0)

    def stats(self) -> dict:
        s = Stats()
        self._check(self.lib.rsa_ext_get_stats(self.h, C.byref(s)))
        return s.asdict()


def alninfo_cigar_string(rec) -> str:
    """CIGAR text of one ALNINFO record, as Cigar::to_string prints it (reference src/cigar.cpp:47-53)."""
    return "".join(f"{int(op) >> 4}{'MIDNSHP=X'[int(op) & 0xF]}" for op in rec["cigar"][:int(rec["n_cigar"])])


def plan_debug(qoff: np.ndarray, toff: np.ndarray, scratch_cap: int = 1 << 32, exact_only: bool = False,
               time_reps: int = 0) -> dict:
    """Host-only planning probe (no CUDA call): how the first chunk of a batch would be routed."""
    lib = load_library()
    out = np.zeros(8, np.int64)
    out[7] = time_reps
    rc = lib.rsa_ext_plan_debug(len(qoff) - 1, qoff.ctypes.data, toff.ctypes.data, scratch_cap, int(exact_only),
                                out.ctypes.data)
    if rc != 0:
        raise ExtensionError(rc, "plan_debug failed")
    keys = ["pairs", "fast_pairs", "exact_pairs", "failed", "groups", "scratch_bytes", "fast_classes", "plan_ns"]
    return dict(zip(keys, out.tolist()))


def scan_debug(qoff: np.ndarray, toff: np.ndarray, scratch_cap: int = 1 << 32, time_reps: int = 0) -> dict:
    """Host pass of the device planner (no CUDA call): chunk cut, routing counts, launch geometry, scratch bound."""
    lib = load_library()
    out = np.zeros(8, np.int64)
    out[7] = time_reps
    rc = lib.rsa_ext_scan_debug(len(qoff) - 1, qoff.ctypes.data, toff.ctypes.data, scratch_cap, out.ctypes.data)
    if rc != 0:
        raise ExtensionError(rc, "scan_debug failed")
    keys = ["pairs", "fast_pairs", "exact_pairs", "failed", "group_slots", "scratch_bound", "fast_classes", "scan_ns"]
    return dict(zip(keys, out.tolist()))
