#!/usr/bin/env python3
"""Per-tile phase timing of the round-0 re-rank (clock64 stamps by thread 0 of each CTA, a look-back warp).
    python tools/rerank_trace.py [c2|c5|c3]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from dark_b200 import saca, synth, _ffi  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
kind, seed, n = synth.CONFIGS[name]
text = synth.generate(kind, seed, n)
con = saca.Constructor(n, flags=_ffi.F_DEVICE_ONLY)
dt = torch.from_numpy(text).cuda()
db = torch.empty(n, dtype=torch.uint8, device="cuda")
con.bwt_device(dt.data_ptr(), n, db.data_ptr())  # warm
L = _ffi.lib()
L.dark_bwt_debug_trace_rerank.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
tiles = (n + 4095) // 4096
trace = torch.zeros(tiles * 8, dtype=torch.int64, device="cuda")
L.dark_bwt_debug_trace_rerank(con._ctx, trace.data_ptr())
con.bwt_device(dt.data_ptr(), n, db.data_ptr())
st = con.stats.as_dict()
L.dark_bwt_debug_trace_rerank(con._ctx, None)
t = trace.cpu().numpy().reshape(-1, 8)
t = t[t[:, 7] != 0]
d = np.diff(t, axis=1).astype(np.float64)
names = ["claim tile", "load + head flags", "thread/warp scan + sync", "publish + look-back (warp 0)", "barrier after look-back",
         "apply (SA/BWT stores, staging)", "sync + survivor write-out"]
mid = d[len(d) // 4: 3 * len(d) // 4]
print(f"{name}: {len(t)} tiles; rerank phase {st['rerank_ms']:.3f} ms (all rounds); cycles per phase, median / mean over the middle half:")
for i, nm in enumerate(names):
    print(f"  {nm:34s} {np.median(mid[:, i]):9.0f} {mid[:, i].mean():9.0f}")
tot = (t[:, 7] - t[:, 0])[len(t) // 4: 3 * len(t) // 4]
print(f"  {'tile total':34s} {np.median(tot):9.0f} {tot.mean():9.0f}")
con.close()
