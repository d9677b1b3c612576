"""ctypes binding of the device-side SAM record formatter (include/rsa_sam.h, part of librsa_ext.so): the host-side
mirror of the reference's `class Sam` (reference src/sam.hpp:81-135, src/sam.cpp) for whole batches of records."""
import ctypes as C
from typing import Optional, Sequence

import numpy as np

from .ext import ExtensionError, load_library

RECORD_DTYPE = np.dtype([
    ("kind", "<u4"), ("flags", "<u4"), ("ref_id", "<i4"), ("pos", "<u4"), ("mapq", "<u4"), ("mate_ref", "<i4"),
    ("mate_pos", "<u4"), ("tlen", "<i4"), ("edit_distance", "<i4"), ("score", "<i4"), ("cigar_off", "<u4"),
    ("n_cigar", "<u4"), ("name_len", "<u4"), ("seq_len", "<u4"), ("qual_len", "<u4"), ("details", "<u4", (5,)),
    ("name_off", "<u8"), ("seq_off", "<u8"), ("qual_off", "<u8")], align=True)
assert RECORD_DTYPE.itemsize == 104
ALIGNMENT_DTYPE = np.dtype([("ref_id", "<i4"), ("ref_start", "<i4"), ("edit_distance", "<i4"), ("score", "<i4"),
                            ("length", "<i4"), ("is_rc", "<i4"), ("is_unaligned", "<i4"), ("cigar_off", "<u4"),
                            ("n_cigar", "<u4")], align=True)
assert ALIGNMENT_DTYPE.itemsize == 36
READ_DTYPE = np.dtype([("name_off", "<u8"), ("seq_off", "<u8"), ("qual_off", "<u8"), ("name_len", "<u4"),
                       ("seq_len", "<u4"), ("qual_len", "<u4")], align=True)
assert READ_DTYPE.itemsize == 40

SAM_ABI_SYMBOLS = ["rsa_sam_create", "rsa_sam_destroy", "rsa_sam_last_error", "rsa_sam_format", "rsa_sam_kernel_ms", "rsa_sam_single",
                   "rsa_sam_pair", "rsa_sam_unmapped"]

KIND_ALIGNED, KIND_UNMAPPED, KIND_UNMAPPED_MATE = 0, 1, 2

_bound = False


def _lib():
    global _bound
    lib = load_library()
    if not _bound:
        vp, i64, i32, u32 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32
        lib.rsa_sam_create.argtypes = [i32, i32, vp, vp, i32, C.c_char_p, i32, i32, C.POINTER(vp)]
        lib.rsa_sam_create.restype = C.c_int
        lib.rsa_sam_destroy.argtypes = [vp]
        lib.rsa_sam_destroy.restype = None
        lib.rsa_sam_last_error.argtypes = [vp]
        lib.rsa_sam_last_error.restype = C.c_char_p
        lib.rsa_sam_format.argtypes = [vp, i64, vp, vp, i64, vp, i64, vp, i64, vp, vp]
        lib.rsa_sam_format.restype = C.c_int
        lib.rsa_sam_kernel_ms.argtypes = [vp]
        lib.rsa_sam_kernel_ms.restype = C.c_double
        lib.rsa_sam_single.argtypes = [vp, vp, u32, i32, vp, vp]
        lib.rsa_sam_single.restype = None
        lib.rsa_sam_pair.argtypes = [vp, vp, vp, vp, u32, u32, i32, i32, vp, vp, vp]
        lib.rsa_sam_pair.restype = None
        lib.rsa_sam_unmapped.argtypes = [vp, u32, vp]
        lib.rsa_sam_unmapped.restype = None
        _bound = True
    return lib


class SamFormatter:
    """One formatter per output stream: reference names, CIGAR style (=/X or M), read group, unmapped/detail switches
    -- the constructor arguments of the reference's `Sam` (src/sam.hpp:84-102)."""

    def __init__(self, ref_names: Sequence[bytes], cigar_m: bool = False, read_group: bytes = b"", output_unmapped: bool = True,
                 show_details: bool = False, device: int = 0):
        self.lib = _lib()
        buf = b"".join(ref_names)
        off = np.zeros(len(ref_names) + 1, np.int64)
        off[1:] = np.cumsum([len(x) for x in ref_names])
        nb = np.frombuffer(buf, np.uint8).copy() if buf else np.zeros(1, np.uint8)
        h = C.c_void_p()
        rc = self.lib.rsa_sam_create(device, len(ref_names), nb.ctypes.data, off.ctypes.data, int(cigar_m), read_group,
                                     int(output_unmapped), int(show_details), C.byref(h))
        if rc != 0:
            raise ExtensionError(rc, self.lib.rsa_sam_last_error(None).decode())
        self.h = h

    def kernel_ms(self) -> float:
        return float(self.lib.rsa_sam_kernel_ms(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.rsa_sam_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def format(self, records: np.ndarray, text_pool: np.ndarray, cigar_pool: np.ndarray, want_offsets: bool = False,
               out: Optional[np.ndarray] = None):
        """Returns the SAM text of the records (bytes), optionally with the n + 1 line offsets.  `out`: a caller-owned
        uint8 buffer (e.g. pinned) to receive the text; the return value is then a view of it instead of a copy."""
        n = len(records)
        records = np.ascontiguousarray(records, dtype=RECORD_DTYPE)
        cigar_pool = np.ascontiguousarray(cigar_pool, dtype=np.uint32)
        if out is None:
            cap = int(records["name_len"].sum() + records["seq_len"].sum() + records["qual_len"].sum()) + 160 * n + 12 * len(cigar_pool) + 64
            buf = np.empty(cap, np.uint8)
        else:
            buf, cap = out, len(out)
        out_len = C.c_int64(0)
        offs = np.zeros(n + 1, np.int64) if want_offsets else None
        rc = self.lib.rsa_sam_format(self.h, n, records.ctypes.data, text_pool.ctypes.data, len(text_pool),
                                     cigar_pool.ctypes.data if len(cigar_pool) else None, len(cigar_pool), buf.ctypes.data, cap,
                                     C.byref(out_len), offs.ctypes.data if want_offsets else None)
        if rc != 0:
            raise ExtensionError(rc, self.lib.rsa_sam_last_error(self.h).decode())
        text = buf[:out_len.value] if out is not None else buf[:out_len.value].tobytes()
        return (text, offs) if want_offsets else text


def single_record(a: np.ndarray, read: np.ndarray, mapq: int, is_primary: bool, details=None) -> np.ndarray:
    """Sam::add (reference src/sam.cpp:117-139) -> one record descriptor."""
    out = np.zeros(1, RECORD_DTYPE)
    d = np.ascontiguousarray(details if details is not None else np.zeros(5), dtype=np.uint32)
    _lib().rsa_sam_single(a.ctypes.data, read.ctypes.data, mapq, int(is_primary), d.ctypes.data, out.ctypes.data)
    return out


def pair_records(a1, a2, r1, r2, mapq1: int, mapq2: int, is_proper: bool, is_primary: bool, details1=None, details2=None) -> np.ndarray:
    """Sam::add_pair (reference src/sam.cpp:208-318) -> two record descriptors."""
    out = np.zeros(2, RECORD_DTYPE)
    d1 = np.ascontiguousarray(details1 if details1 is not None else np.zeros(5), dtype=np.uint32)
    d2 = np.ascontiguousarray(details2 if details2 is not None else np.zeros(5), dtype=np.uint32)
    _lib().rsa_sam_pair(a1.ctypes.data, a2.ctypes.data, r1.ctypes.data, r2.ctypes.data, mapq1, mapq2, int(is_proper), int(is_primary),
                        d1.ctypes.data, d2.ctypes.data, out.ctypes.data)
    return out


def unmapped_record(read: np.ndarray, flags: int) -> np.ndarray:
    """Sam::add_unmapped (reference src/sam.cpp:73-86)."""
    out = np.zeros(1, RECORD_DTYPE)
    _lib().rsa_sam_unmapped(read.ctypes.data, flags, out.ctypes.data)
    return out
