"""Inputs and a pure-Python restatement for the SAM formatter tests.

`make_calls` builds a random list of the calls the reference's workers make on `class Sam` (reference src/sam.hpp:104-108:
add, add_pair, add_unmapped, add_unmapped_pair) with everything that changes the text: /1 /2 name suffixes, reverse
strand, secondary records, unaligned mates, reads on different references, empty qualities, soft-masked / N / odd bases,
multi-digit and zero positions, negative template lengths, =/X/I/D/S CIGARs.  `restate` formats them in plain Python,
following reference src/sam.cpp line by line; the CPU suite pins it to the reference's own writer
(oracle/_ref/libsam_ref.so), the GPU suite compares rsa_sam_format with both."""
import numpy as np

import oracle
from rabbitsalign_b200 import sam as S

PAIRED, PROPER_PAIR, UNMAP, MUNMAP, REVERSE, MREVERSE, READ1, READ2, SECONDARY = 1, 2, 4, 8, 0x10, 0x20, 0x40, 0x80, 0x100
REVCOMP = {ord(a): ord(b) for a, b in zip("ACGTUacgtu", "TGCAATGCAA")}


def make_calls(n: int, seed: int, n_refs: int = 5):
    rng = np.random.default_rng(seed)
    ref_names = [b"chr%d" % (i + 1) if i % 2 == 0 else b"contig_%d.scaffold" % i for i in range(n_refs)]
    text = bytearray()
    cig = []

    def put(b: bytes):
        off = len(text)
        text.extend(b)
        return off, len(b)

    def read(i, mate):
        name = [b"read%d" % i, b"read%d/%d" % (i, mate), b"r/1", b"/2", b"x", b"sim.%d:%d/3" % (i, mate)][int(rng.integers(0, 6))]
        L = int(rng.choice([0, 1, 36, 100, 150, 151, 250]))
        alpha = np.frombuffer(b"ACGTACGTACGTNacgtnRYU", np.uint8)
        seq = alpha[rng.integers(0, len(alpha), L)].tobytes()
        qual = b"" if rng.random() < 0.1 else bytes(rng.integers(33, 74, L).astype(np.uint8))
        r = np.zeros(1, S.READ_DTYPE)
        r["name_off"], r["name_len"] = put(name)
        r["seq_off"], r["seq_len"] = put(seq)
        r["qual_off"], r["qual_len"] = put(qual)
        return r

    def alignment(unaligned=False):
        a = np.zeros(1, S.ALIGNMENT_DTYPE)
        a["ref_id"] = rng.integers(0, n_refs)
        a["ref_start"] = int(rng.choice([0, 9, 99, 12345, 99999999, int(rng.integers(0, 2_000_000_000))]))
        a["edit_distance"] = rng.integers(0, 40)
        a["score"] = int(rng.choice([0, 7, 300, 500, int(rng.integers(0, 600))]))
        a["length"] = rng.integers(1, 400)
        a["is_rc"] = rng.integers(0, 2)
        a["is_unaligned"] = int(unaligned)
        k = int(rng.choice([0, 1, 2, 5, 12, 30])) if not unaligned else 0
        ops = []
        for _ in range(k):
            op = int(rng.choice([7, 8, 1, 2, 4, 0, 7, 8]))
            ops.append((int(rng.integers(1, 200)) << 4) | op)
        a["cigar_off"], a["n_cigar"] = len(cig), len(ops)
        cig.extend(ops)
        return a

    calls = np.zeros(n, oracle.SAM_CALL_DTYPE)
    for i in range(n):
        c = calls[i]
        kind = int(rng.choice([0, 1, 1, 1, 2, 3]))
        c["kind"] = kind
        c["is_primary"] = int(rng.random() < 0.85)
        c["is_proper"] = int(rng.random() < 0.6)
        c["mapq1"], c["mapq2"] = int(rng.choice([0, 3, 60, 255])), int(rng.integers(0, 61))
        c["details1"] = rng.integers(0, 2000, 5); c["details2"] = rng.integers(0, 2000, 5)
        c["details1"][1] = rng.integers(0, 2); c["details2"][1] = rng.integers(0, 2)   # nam_rescue is a bool
        c["r1"] = read(i, 1)[0]
        c["r2"] = read(i, 2)[0]
        if kind == 0:
            c["a1"] = alignment()[0]
        elif kind == 1:
            u = int(rng.choice([0, 0, 0, 1, 2]))   # 1: read 1 unaligned, 2: read 2 unaligned
            a1, a2 = alignment(u == 1), alignment(u == 2)
            if rng.random() < 0.6:
                a2["ref_id"] = a1["ref_id"]
                a2["ref_start"] = max(0, int(a1["ref_start"][0]) + int(rng.integers(-600, 600)))
            c["a1"], c["a2"] = a1[0], a2[0]
        elif kind == 2:
            c["unmapped_flags"] = int(rng.choice([UNMAP, UNMAP | PAIRED | READ1, UNMAP | PAIRED | MUNMAP | READ2]))
    return ref_names, calls, np.frombuffer(bytes(text), np.uint8).copy(), np.array(cig, np.uint32)


def records_from_calls(calls):
    """The product's record descriptors for the same calls, through the ABI's host helpers (rsa_sam_single / _pair /
    _unmapped mirror Sam::add / add_pair / add_unmapped)."""
    recs = []
    for c in calls:
        a1 = np.array([c["a1"]], S.ALIGNMENT_DTYPE); a2 = np.array([c["a2"]], S.ALIGNMENT_DTYPE)
        r1 = np.array([c["r1"]], S.READ_DTYPE); r2 = np.array([c["r2"]], S.READ_DTYPE)
        if c["kind"] == 0:
            recs.append(S.single_record(a1, r1, int(c["mapq1"]), bool(c["is_primary"]), c["details1"]))
        elif c["kind"] == 1:
            recs.append(S.pair_records(a1, a2, r1, r2, int(c["mapq1"]), int(c["mapq2"]), bool(c["is_proper"]), bool(c["is_primary"]),
                                       c["details1"], c["details2"]))
        elif c["kind"] == 2:
            recs.append(S.unmapped_record(r1, int(c["unmapped_flags"])))
        else:
            recs.append(S.unmapped_record(r1, PAIRED | UNMAP | MUNMAP | READ1))
            recs.append(S.unmapped_record(r2, PAIRED | UNMAP | MUNMAP | READ2))
    return np.concatenate(recs)


# ---- plain restatement of reference src/sam.cpp ---------------------------------------------------------------------
def _strip_suffix(name: bytes) -> bytes:                      # sam.cpp:31-43
    if len(name) >= 2 and name[-2:-1] == b"/" and name[-1:] in (b"1", b"2"):
        return name[:-2]
    return name


def _cigar_text(ops, cigar_m: bool) -> bytes:                 # sam.cpp:61-71, cigar.cpp:6-18,47-53
    if len(ops) == 0:
        return b"*"
    if cigar_m:
        merged = []
        for v in ops:
            op, ln = int(v) & 0xF, int(v) >> 4
            if op in (7, 8):
                op = 0
            if merged and merged[-1][0] == op:
                merged[-1][1] += ln
            else:
                merged.append([op, ln])
        return b"".join(b"%d%c" % (ln, b"MIDNSHP=X"[op]) for op, ln in merged)
    return b"".join(b"%d%c" % (int(v) >> 4, b"MIDNSHP=X"[int(v) & 0xF]) for v in ops)


def restate(ref_names, calls, text, cigars, cigar_m=False, read_group=b"", output_unmapped=True, show_details=False) -> bytes:
    tail = (b"\tRG:Z:" + read_group + b"\n") if read_group else b"\n"
    out = []
    raw = text.tobytes()
    u32 = lambda v: int(v) & 0xFFFFFFFF

    def fld(r, k):
        return raw[int(r[k + "_off"]):int(r[k + "_off"]) + int(r[k + "_len"])]

    def add_unmapped(r, flags):                               # sam.cpp:73-86
        if not output_unmapped:
            return
        out.append(_strip_suffix(fld(r, "name")) + b"\t%d\t*\t0\t0\t*\t*\t0\t0\t" % flags + (fld(r, "seq") or b"*") + b"\t" + (fld(r, "qual") or b"*") + tail)

    def add_unmapped_mate(r, flags, mate_ref_name, mate_pos):  # sam.cpp:88-110
        p = u32(mate_pos + 1)
        out.append(_strip_suffix(fld(r, "name")) + b"\t%d\t" % flags + mate_ref_name + b"\t%d\t0\t*\t=\t%d\t0\t" % (p, p) +
                   (fld(r, "seq") or b"*") + b"\t" + (fld(r, "qual") or b"*") + tail)

    def add_record(r, flags, ref_name, pos, mapq, a, mate_name, mate_pos, tlen, details):   # sam.cpp:141-206
        seq, qual = fld(r, "seq"), fld(r, "qual")
        ops = cigars[int(a["cigar_off"]):int(a["cigar_off"]) + int(a["n_cigar"])]
        line = _strip_suffix(fld(r, "name")) + b"\t%d\t" % flags + ref_name + b"\t%d\t%d\t" % (u32(pos + 1), mapq & 0xFF) + \
            _cigar_text(ops, cigar_m) + b"\t" + mate_name + b"\t%d\t%d\t" % (u32(mate_pos + 1), tlen)
        if flags & SECONDARY:
            line += b"*"
        elif flags & REVERSE:
            line += bytes(REVCOMP.get(ch, ord("N")) for ch in reversed(seq)) or b"*"
        else:
            line += seq or b"*"
        if not (flags & UNMAP):
            if flags & SECONDARY:
                line += b"\t*"
            elif flags & REVERSE:
                line += b"\t" + (qual[::-1] or b"*")
            else:
                line += b"\t" + (qual or b"*")
            line += b"\tNM:i:%d\tAS:i:%d" % (int(a["edit_distance"]), int(a["score"]))
        else:
            line += b"\t" + (qual or b"*")
        if show_details:
            line += b"\tna:i:%d\tnr:i:%d\tal:i:%d\tga:i:%d" % (details[0], 1 if details[1] else 0, details[2], details[3])
            if flags & PAIRED:
                line += b"\tmr:i:%d" % details[4]
        out.append(line + tail)

    for c in calls:
        k = int(c["kind"])
        if k == 0:                                            # Sam::add, sam.cpp:117-139
            a = c["a1"]
            flags, mapq = 0, int(c["mapq1"])
            if not a["is_unaligned"] and a["is_rc"]:
                flags |= REVERSE
            if not c["is_primary"]:
                flags |= SECONDARY
                mapq = 255
            add_record(c["r1"], flags, ref_names[int(a["ref_id"])], int(a["ref_start"]), mapq, a, b"*", -1, 0, c["details1"])
        elif k == 1:                                          # Sam::add_pair, sam.cpp:208-318
            a1, a2 = c["a1"], c["a2"]
            f1, f2 = PAIRED | READ1, PAIRED | READ2
            if not c["is_primary"]:
                f1 |= SECONDARY; f2 |= SECONDARY
            tl1 = 0
            both = not a1["is_unaligned"] and not a2["is_unaligned"]
            if both and a1["ref_id"] == a2["ref_id"]:
                dist = int(a2["ref_start"]) - int(a1["ref_start"])
                tl1 = dist + int(a2["length"]) if dist > 0 else dist - int(a1["length"])
            if c["is_proper"]:
                f1 |= PROPER_PAIR; f2 |= PROPER_PAIR
            pos1, pos2 = int(a1["ref_start"]), int(a2["ref_start"])
            if a1["is_unaligned"]:
                f1 |= UNMAP; f2 |= MUNMAP; pos1 = -1; n1 = b"*"
            else:
                if a1["is_rc"]:
                    f1 |= REVERSE; f2 |= MREVERSE
                n1 = ref_names[int(a1["ref_id"])]
            if a2["is_unaligned"]:
                f2 |= UNMAP; f1 |= MUNMAP; pos2 = -1; n2 = b"*"
            else:
                if a2["is_rc"]:
                    f1 |= MREVERSE; f2 |= REVERSE
                n2 = ref_names[int(a2["ref_id"])]
            m1, m2 = n1, n2
            if (both and a1["ref_id"] == a2["ref_id"]) or (bool(a1["is_unaligned"]) != bool(a2["is_unaligned"])):
                m1 = m2 = b"="
            if bool(a1["is_unaligned"]) != bool(a2["is_unaligned"]):
                if a1["is_unaligned"]:
                    pos1 = pos2
                else:
                    pos2 = pos1
            if a1["is_unaligned"]:
                add_unmapped_mate(c["r1"], f1, n2, pos2)
            else:
                add_record(c["r1"], f1, n1, int(a1["ref_start"]), int(c["mapq1"]), a1, m2, pos2, tl1, c["details1"])
            if a2["is_unaligned"]:
                add_unmapped_mate(c["r2"], f2, n1, pos1)
            else:
                add_record(c["r2"], f2, n2, int(a2["ref_start"]), int(c["mapq2"]), a2, m1, pos1, -tl1, c["details2"])
        elif k == 2:
            add_unmapped(c["r1"], int(c["unmapped_flags"]))
        else:                                                 # add_unmapped_pair, sam.cpp:112-115
            add_unmapped(c["r1"], PAIRED | UNMAP | MUNMAP | READ1)
            add_unmapped(c["r2"], PAIRED | UNMAP | MUNMAP | READ2)
    return b"".join(out)


VARIANTS = [dict(), dict(cigar_m=True), dict(read_group=b"grp1", show_details=True), dict(output_unmapped=False, cigar_m=True)]
