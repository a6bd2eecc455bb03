#!/usr/bin/env python3
"""Per-tile phase timing of the TMA-staged radix pass (onesweep_tma.cuh; clock64 stamps by thread 0 of each CTA).
Needs a tuning build with -DDARK_TUNE_TRACE (DARK_BWT_LIB=...).   python tools/pass_trace2.py [log2_m]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from dark_b200 import saca, _ffi  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 27
m = 1 << lg
g = torch.Generator(device="cuda").manual_seed(1)
keys = torch.randint(0, 1 << 62, (m,), dtype=torch.int64, device="cuda", generator=g)
vals = torch.arange(m, dtype=torch.int32, device="cuda")
k2, v2 = torch.empty_like(keys), torch.empty_like(vals)
con = saca.Constructor(m, flags=_ffi.F_DEVICE_ONLY)
L = _ffi.lib()
L.dark_bwt_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
ntiles = m // 4096 + 1
trace = torch.zeros(ntiles * 12, dtype=torch.int64, device="cuda")
con.sort_pairs_device(keys.data_ptr(), vals.data_ptr(), k2.data_ptr(), v2.data_ptr(), m, 0, 8)   # warm
L.dark_bwt_debug_trace(con._ctx, trace.data_ptr())
_, ms = con.sort_pairs_device(keys.data_ptr(), vals.data_ptr(), k2.data_ptr(), v2.data_ptr(), m, 0, 8)  # one pass
torch.cuda.synchronize()
t = trace.cpu().numpy().reshape(-1, 12)
t = t[t[:, 10] != 0]
names = ["wait stage + regs", "barrier + claim/TMA issue", "rank", "barrier", "column scan + publish", "fold + wait prefix",
         "barrier", "scatter prev", "barrier", "re-order"]
d = np.diff(t[:, :11], axis=1).astype(np.float64)
mid = d[len(d) // 4: 3 * len(d) // 4]
print(f"{len(t)} tiles, one pass {ms:.3f} ms (with histogram); cycles per phase, median / mean / p90 over the middle half:")
for i, nme in enumerate(names):
    print(f"  {nme:28s} {np.median(mid[:, i]):9.0f} {mid[:, i].mean():9.0f} {np.percentile(mid[:, i], 90):9.0f}")
tot = (t[:, 10] - t[:, 0])[len(t) // 4: 3 * len(t) // 4]
print(f"  {'iteration total':28s} {np.median(tot):9.0f} {tot.mean():9.0f}")
sm = t[:, 11].astype(np.int64)
print(f"  SMs used {len(np.unique(sm))}; tiles per SM min/max {np.bincount(sm).min()}/{np.bincount(sm).max()}")
