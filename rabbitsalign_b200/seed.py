"""ctypes binding of the seeding C ABI (include/rsa_seed.h, part of librsa_ext.so): randstrobe seeding, index lookup and
NAM merging for a batch of reads on the GPU, mirroring what the reference does per read in align_SE_read_part /
align_PE_read_part (reference src/aln.cpp:1937-1958: randstrobes_query, find_nams, find_nams_rescue).

No CPU fallback: constructing a `SeedIndexGpu` without the library or without a CUDA device raises."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from .ext import load_library

NAM_DTYPE = np.dtype([("query_start", "<i4"), ("query_end", "<i4"), ("query_prev_hit_startpos", "<i4"),
                      ("ref_start", "<i4"), ("ref_end", "<i4"), ("ref_prev_hit_startpos", "<i4"),
                      ("n_hits", "<i4"), ("ref_id", "<i4"), ("score", "<f4"), ("flags", "<u4")])
READ_DTYPE = np.dtype([("nam_off", "<u4"), ("n_nams", "<i4"), ("nonrepetitive_fraction", "<f4"), ("flags", "<u4")])
assert NAM_DTYPE.itemsize == 40 and READ_DTYPE.itemsize == 16
READ_RESCUED, READ_FAILED = 1, 2

SEED_ABI_SYMBOLS = ["rsa_seed_index_upload", "rsa_seed_index_free", "rsa_seed_create", "rsa_seed_destroy",
                    "rsa_seed_last_error", "rsa_seed_find_nams", "rsa_seed_get_stats", "rsa_seed_stage",
                    "rsa_seed_run_staged", "rsa_seed_stream"]


class SeedConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("k", C.c_int32), ("s", C.c_int32), ("t_syncmer", C.c_int32),
                ("w_min", C.c_int32), ("w_max", C.c_int32), ("max_dist", C.c_int32), ("q", C.c_uint64),
                ("bits", C.c_int32), ("filter_cutoff", C.c_uint32), ("rescue_level", C.c_int32),
                ("rescue_cutoff", C.c_uint32)]


class SeedStats(C.Structure):
    _fields_ = [("reads", C.c_int64), ("nams", C.c_int64), ("reads_rescued", C.c_int64), ("reads_retried", C.c_int64),
                ("reads_failed", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("kernel_ms", C.c_double),
                ("kernel_launches", C.c_int64), ("kernel_ms_large", C.c_double)]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def make_config(params: dict, device: int = 0, rescue_level: int = 2) -> SeedConfig:
    """params: bits, filter_cutoff, k, s, t_syncmer, w_min, w_max, max_dist, q (the reference's IndexParameters +
    StrobemerIndex fields); rescue_cutoff follows reference src/main.cpp:415."""
    rc = rescue_level * params["filter_cutoff"] if rescue_level < 100 else 1000
    return SeedConfig(device, params["k"], params["s"], params["t_syncmer"], params["w_min"], params["w_max"],
                      params["max_dist"], params["q"], params["bits"], params["filter_cutoff"], rescue_level, rc)


class SeedError(RuntimeError):
    pass


def _lib():
    lib = load_library()
    if not getattr(lib, "_seed_typed", False):
        vp, i64 = C.c_void_p, C.c_int64
        lib.rsa_seed_index_upload.argtypes = [C.POINTER(SeedConfig), vp, i64, vp, i64, C.POINTER(vp)]
        lib.rsa_seed_index_upload.restype = C.c_int
        lib.rsa_seed_index_free.argtypes = [vp]
        lib.rsa_seed_index_free.restype = None
        lib.rsa_seed_create.argtypes = [vp, C.POINTER(vp)]
        lib.rsa_seed_create.restype = C.c_int
        lib.rsa_seed_destroy.argtypes = [vp]
        lib.rsa_seed_destroy.restype = None
        lib.rsa_seed_last_error.argtypes = [vp]
        lib.rsa_seed_last_error.restype = C.c_char_p
        lib.rsa_seed_find_nams.argtypes = [vp, i64, vp, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64)]
        lib.rsa_seed_find_nams.restype = C.c_int
        lib.rsa_seed_get_stats.argtypes = [vp, C.POINTER(SeedStats)]
        lib.rsa_seed_get_stats.restype = C.c_int
        lib.rsa_seed_stage.argtypes = [vp, i64, vp, vp]
        lib.rsa_seed_stage.restype = C.c_int
        lib.rsa_seed_run_staged.argtypes = [vp]
        lib.rsa_seed_run_staged.restype = C.c_int
        lib.rsa_seed_stream.argtypes = [vp]
        lib.rsa_seed_stream.restype = vp
        lib._seed_typed = True
    return lib


class SeedIndexGpu:
    """The reference's index arrays resident on one GPU (rsa_seed_index_upload)."""

    def __init__(self, cfg: SeedConfig, randstrobes: np.ndarray, starts: np.ndarray):
        self.lib = _lib()
        self.cfg = cfg
        rs = np.ascontiguousarray(randstrobes).view(np.uint8)
        st = np.ascontiguousarray(starts, dtype=np.uint64)
        h = C.c_void_p()
        rc = self.lib.rsa_seed_index_upload(C.byref(cfg), rs.ctypes.data, len(rs) // 16, st.ctypes.data, len(st), C.byref(h))
        if rc != 0:
            raise SeedError(f"rsa_seed_index_upload status {rc}: {self.lib.rsa_seed_last_error(None).decode()}")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.rsa_seed_index_free(self.h)
            self.h = None


class Seeder:
    """One host worker's seeding handle (stream + buffers) on an uploaded index."""

    def __init__(self, index: SeedIndexGpu):
        self.lib = _lib()
        self.index = index
        h = C.c_void_p()
        rc = self.lib.rsa_seed_create(index.h, C.byref(h))
        if rc != 0:
            raise SeedError(f"rsa_seed_create status {rc}: {self.lib.rsa_seed_last_error(None).decode()}")
        self.h = h

    def _check(self, rc):
        if rc != 0:
            raise SeedError(f"rsa_seed status {rc}: {self.lib.rsa_seed_last_error(self.h).decode()}")

    def find_nams(self, reads: np.ndarray, roff: np.ndarray, copy: bool = True):
        """-> (per_read[n] as READ_DTYPE, nams[total] as NAM_DTYPE).  copy=False returns views of the handle's pinned result
        buffers (valid until the next call on this handle), which is what a C caller gets."""
        assert reads.dtype == np.uint8 and roff.dtype == np.int64
        n = len(roff) - 1
        pr, nm, tot = C.c_void_p(), C.c_void_p(), C.c_int64()
        self._check(self.lib.rsa_seed_find_nams(self.h, n, reads.ctypes.data, roff.ctypes.data, C.byref(pr), C.byref(nm), C.byref(tot)))
        per = np.frombuffer((C.c_uint8 * (16 * n)).from_address(pr.value), dtype=READ_DTYPE) if n else np.zeros(0, READ_DTYPE)
        nams = (np.frombuffer((C.c_uint8 * (40 * tot.value)).from_address(nm.value), dtype=NAM_DTYPE)
                if tot.value else np.zeros(0, NAM_DTYPE))
        return (per.copy(), nams.copy()) if copy else (per, nams)

    def stage(self, reads, roff):
        self._keep = (reads, roff)
        self._check(self.lib.rsa_seed_stage(self.h, len(roff) - 1, reads.ctypes.data, roff.ctypes.data))

    def run_staged(self):
        self._check(self.lib.rsa_seed_run_staged(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.rsa_seed_stream(self.h) or 0)

    def stats(self) -> dict:
        s = SeedStats()
        self._check(self.lib.rsa_seed_get_stats(self.h, C.byref(s)))
        return s.asdict()

    def close(self):
        if getattr(self, "h", None):
            self.lib.rsa_seed_destroy(self.h)
            self.h = None


def apply_group_order(per: np.ndarray, nams: np.ndarray, map_order) -> np.ndarray:
    """Re-order every read's NAM groups the way the reference's hash map iterates them (include/rsa_seed.h): `map_order`
    maps the reference ids of one strand in first-touch order to their iteration order (the binding gets it from the
    reference's own container).  Returns the NAM array in the reference's order (reads consecutive, in read order)."""
    out = np.zeros(len(nams), dtype=nams.dtype)
    w = 0
    strand = nams["flags"] & 1
    group = nams["flags"] >> 8
    for r in range(len(per)):
        lo, cnt = int(per["nam_off"][r]), int(per["n_nams"][r])
        if cnt == 0:
            continue
        seg = nams[lo:lo + cnt]
        sg, gg = strand[lo:lo + cnt], group[lo:lo + cnt]
        for s in (0, 1):
            idx = np.nonzero(sg == s)[0]
            if len(idx) == 0:
                continue
            ng = int(gg[idx].max()) + 1
            if ng == 1:
                out[w:w + len(idx)] = seg[idx]
                w += len(idx)
                continue
            first = {}
            for i in idx:
                first.setdefault(int(gg[i]), int(seg["ref_id"][i]))
            keys = [first.get(g, None) for g in range(ng)]
            # groups that produced no NAM cannot be seen here; a group always yields at least one NAM
            assert all(k is not None for k in keys)
            order = [int(x) for x in map_order(keys)]
            for ref in order:
                g = keys.index(ref)
                sel = idx[gg[idx] == g]
                out[w:w + len(sel)] = seg[sel]
                w += len(sel)
    assert w == len(nams)
    return out
