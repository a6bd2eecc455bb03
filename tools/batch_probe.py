#!/usr/bin/env python3
"""Where does the pipelined batch entry lose time?  Wall clock vs per-block device time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dark_b200 import saca, synth
n = 1 << 28
h_text = torch.empty(n, dtype=torch.uint8).pin_memory()
synth.generate("dna", 1, n, out=h_text.numpy())
o1 = torch.empty(n, dtype=torch.uint8).pin_memory()
o2 = torch.empty(n, dtype=torch.uint8).pin_memory()
con = saca.Constructor(n)
con.bwt_into(h_text.data_ptr(), n, o1.data_ptr())
for cnt in (1, 2, 4, 8):
    t0 = time.perf_counter()
    origins, stats = con.bwt_batch_into([h_text.data_ptr()] * cnt, [n] * cnt, [o1.data_ptr(), o2.data_ptr()] * (cnt // 2) + [o1.data_ptr()] * (cnt % 2), want_stats=True)
    dt = time.perf_counter() - t0
    print(f"batch {cnt}: wall {dt*1e3:.1f} ms = {dt*1e3/cnt:.1f} per block; device_ms per block:", [round(s['device_ms'], 1) for s in stats])
# plain copies
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(h_text, non_blocking=True); torch.cuda.synchronize(); h2d = time.perf_counter() - t0
    t0 = time.perf_counter(); o1.copy_(d, non_blocking=True); torch.cuda.synchronize(); d2h = time.perf_counter() - t0
print(f"H2D {h2d*1e3:.2f} ms ({n/h2d/1e9:.1f} GB/s)  D2H {d2h*1e3:.2f} ms ({n/d2h/1e9:.1f} GB/s)")
