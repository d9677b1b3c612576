"""Static issue accounting of a kernel's loop from its SASS: instruction count and the sum of the stall counts in the control
words (bits 41-44 of the upper 64-bit word of each 128-bit instruction), i.e. the fewest cycles ONE warp needs per iteration.

    python tools/sass_stalls.py                     # fast_dp_kernel<4,38,false> steady row loop vs the microbenchmark's bare loop

Used for DESIGN.md 4.1 ("where the distance to the bare recipe goes").  CPU only (cuobjdump)."""
import re, subprocess, sys
def dump(binary, fun):
    out = subprocess.run(["cuobjdump","-sass","-fun",fun,binary],capture_output=True,text=True).stdout
    ins=[]
    lines=out.splitlines()
    i=0
    while i < len(lines):
        m=re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/",lines[i])
        if m and i+1 < len(lines):
            m2=re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/",lines[i+1])
            if m2:
                ins.append((int(m.group(1),16), m.group(2).strip(), int(m.group(3),16), int(m2.group(1),16)))
                i+=2; continue
        i+=1
    return ins
def analyze(ins, lo, hi, label):
    body=[x for x in ins if lo<=x[0]<=hi]
    tot=0; n=0
    from collections import Counter
    hist=Counter(); byop=Counter(); cnt=Counter()
    for a,txt,w0,w1 in body:
        stall=(w1>>41)&0xF; y=(w1>>45)&1
        op=txt.split()[0] if not txt.startswith('@') else txt.split()[1]
        op=op.split('.')[0]
        tot+=stall; n+=1; hist[stall]+=1; byop[op]+=stall; cnt[op]+=1
    print(label, 'instr',n,'sum stall',tot,'avg %.2f'%(tot/max(n,1)))
    print('  hist',sorted(hist.items()))
    print('  by op', [(o,cnt[o],round(byop[o]/cnt[o],2)) for o,_ in cnt.most_common(8)])
    return body
if __name__ == "__main__":
    import os
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    k = dump(os.path.join(ROOT, "rabbitsalign_b200", "librsa_ext.so"),
             "_ZN3rsa14fast_dp_kernelILi4ELi38ELb0EEEvPKhS2_PKNS_8PairMetaEPKNS_9FastGroupEiPhPNS_5DpEndEPNS_10RedoHeaderEPjNS_10FastConstsEiPK5uint4PK5uint2")
    # row loops = backward branches whose body holds two SHFL.UP and 7 DPX add+max per column (tools/sass_report.py); the
    # steady one (no N in the reads, all columns) is the shortest
    loops = []
    for a, txt, w0, w1 in k:
        m = re.match(r"(?:@!?U?P\d+\s+)?BRA\s+0x([0-9a-f]+)", txt)
        if m and int(m.group(1), 16) < a:
            lo = int(m.group(1), 16)
            body = [x for x in k if lo <= x[0] <= a]
            n_dpx = sum(1 for x in body if "VIADDMNMX" in x[1])
            if sum(1 for x in body if x[1].startswith("SHFL.UP")) >= 2 and 7 * 38 <= n_dpx < 8 * 38:
                loops.append((len(body), lo, a))
    n, lo, hi = min(loops)
    analyze(k, lo, hi, "fast_dp_kernel<4,38,false> steady row loop [%#x, %#x]" % (lo, hi))
    b = dump(os.path.join(ROOT, "tools", "dpx_microbench"), "_Z12bench_cell_nILi38EEvPjN3rsa10FastConstsEPKji")
    for a, txt, w0, w1 in b:
        m = re.search(r"BRA(?:\.U)? (?:!?U?P\d, )?0x([0-9a-f]+)", txt)
        if m and int(m.group(1), 16) < a:
            analyze(b, int(m.group(1), 16), a, "bare recipe loop, 38 columns")
