"""Quick probe of the seeding kernels (builder tool, not the bench): reads/s of the device-resident kernels and of the
blocking C-ABI call, next to the reference's seeding path on all host cores.  Uses the oracle only as index builder /
CPU comparator."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import oracle
from rabbitsalign_b200 import seed as S, workload as W

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
contig_len = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
t0 = time.time()
n_fam = int(sys.argv[3]) if len(sys.argv) > 3 else 8
contigs = W.seeding_genome(n_contigs=4, contig_len=contig_len, seed=41, repeat_families=n_fam, copies_per_contig=20)
idx = oracle.build_seed_index(contigs, 150, os.cpu_count())
t_index = time.time() - t0
t0 = time.time()
small, soff = W.seeding_reads(contigs, 50_000, seed=42)
reps = (n_reads + 49_999) // 50_000
buf = np.tile(small, reps)
off = np.concatenate([soff[:-1] + k * int(soff[-1]) for k in range(reps)] + [np.array([reps * int(soff[-1])])]).astype(np.int64)
n = len(off) - 1
t_reads = time.time() - t0
gi = S.SeedIndexGpu(S.make_config(idx.params()), idx.randstrobes, idx.starts)
sd = S.Seeder(gi)
sd.stage(buf, off)
for _ in range(2):
    sd.run_staged()
ms, ms_large = [], []
for _ in range(5):
    sd.run_staged()
    ms.append(sd.stats()["kernel_ms"])
    ms_large.append(sd.stats()["kernel_ms_large"])
st = sd.stats()
t0 = time.time()
sd.find_nams(buf, off, copy=False)
t0 = time.time()
per, nams = sd.find_nams(buf, off, copy=False)
t_call = time.time() - t0
st2 = sd.stats()
cores = os.cpu_count()
sub = 200_000
t0 = time.time()
found = idx.time_find_nams(buf, np.ascontiguousarray(off[:sub + 1]), cores)
t_cpu = time.time() - t0
print(json.dumps({"reads": n, "genome": 4 * contig_len, "index_entries": idx.n_randstrobes, "index_s": round(t_index, 1),
                  "kernel_ms": ms, "kernel_ms_large": ms_large, "repeat_families": n_fam, "reads_per_s_resident": n / (min(ms) * 1e-3), "nams": st["nams"], "retried": st["reads_retried"],
                  "rescued": st["reads_rescued"], "failed": st["reads_failed"], "call_s": t_call, "reads_per_s_call": n / t_call,
                  "h2d": st2["h2d_bytes"], "d2h": st2["d2h_bytes"],
                  "cpu_reference": {"cores": cores, "reads": sub, "s": t_cpu, "reads_per_s": sub / t_cpu}}))
