import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from kmer_mapper_b200 import _lib
from kmer_mapper_b200.device import DeviceIndex, Mapper
w = bench.workload("config1", 1.0)
tindex, bases, offsets = bench.generate(w, 0, torch.device("cuda", 0))
di = DeviceIndex.from_index(tindex, device=0)
n_counts = tindex.max_node_id() + 1
m = Mapper(di, n_counts)
hb = torch.empty(bases.shape[0], dtype=torch.uint8, pin_memory=True); hb.copy_(bases)
ho = torch.empty(offsets.shape[0], dtype=torch.int64, pin_memory=True); ho.copy_(offsets)
hc = torch.empty(n_counts, dtype=torch.int32, pin_memory=True)
torch.cuda.synchronize()
hb_np, ho_np, hc_np = hb.numpy(), ho.numpy(), hc.numpy().view(np.uint32)
for hp in (-1, 0, 1):
    _lib.set_option("host_pack", hp)
    rows = []
    for i in range(10):
        t0 = time.perf_counter(); m.reset(); t1 = time.perf_counter(); m.map_reads(hb_np, ho_np, w["k"]); t2 = time.perf_counter(); m.counts(out=hc_np); t3 = time.perf_counter()
        rows.append((round((t1 - t0) * 1e3, 3), round((t2 - t1) * 1e3, 3), round((t3 - t2) * 1e3, 3)))
    print("host_pack", hp, "reset/map/counts ms per step:", rows, flush=True)
