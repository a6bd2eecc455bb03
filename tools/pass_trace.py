#!/usr/bin/env python3
"""Per-tile phase timing of the radix pass (clock64 stamps by thread 0 of each CTA).
    python tools/pass_trace.py [log2_m] [variant]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from dark_b200 import saca, _ffi  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 27
variant = sys.argv[2] if len(sys.argv) > 2 else "1"
os.environ["DARK_BWT_SORT_VARIANT"] = variant
m = 1 << lg
g = torch.Generator(device="cuda").manual_seed(1)
keys = torch.randint(0, 1 << 62, (m,), dtype=torch.int64, device="cuda", generator=g)
vals = torch.arange(m, dtype=torch.int32, device="cuda")
k2, v2 = torch.empty_like(keys), torch.empty_like(vals)
con = saca.Constructor(m, flags=_ffi.F_DEVICE_ONLY)
L = _ffi.lib()
L.dark_bwt_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
ntiles_max = m // 2048 + 1
trace = torch.zeros(ntiles_max * 12, dtype=torch.int64, device="cuda")
con.sort_pairs_device(keys.data_ptr(), vals.data_ptr(), k2.data_ptr(), v2.data_ptr(), m, 0, 8)   # warm
L.dark_bwt_debug_trace(con._ctx, trace.data_ptr())
_, ms = con.sort_pairs_device(keys.data_ptr(), vals.data_ptr(), k2.data_ptr(), v2.data_ptr(), m, 0, 8)  # one pass
torch.cuda.synchronize()
full = trace.cpu().numpy().reshape(-1, 12)
full = full[full[:, 7] != 0]
t = full[:, :8]
names = ["claim+clear", "load+rank", "sync1", "colpass+scan+fold", "reorder", "lookback", "scatter issue"]
d = np.diff(t, axis=1).astype(np.float64)
mid = d[len(d) // 4: 3 * len(d) // 4]
print(f"variant {variant}: {len(t)} tiles, one pass {ms:.3f} ms (with histogram); cycles per phase, median / mean over the middle half:")
for i, nme in enumerate(names):
    print(f"  {nme:20s} {np.median(mid[:, i]):9.0f} {mid[:, i].mean():9.0f}")
tot = (t[:, 7] - t[:, 0])[len(t) // 4: 3 * len(t) // 4]
print(f"  {'tile total':20s} {np.median(tot):9.0f} {tot.mean():9.0f}")

# SM clock: cycles per nanosecond over each tile; CTA residency: tiles per SM x tile time / kernel time
ns = (full[:, 10] - full[:, 8]).astype(np.float64)
cyc = (full[:, 7] - full[:, 1]).astype(np.float64)
ok = ns > 0
print(f"  SM clock from globaltimer: median {np.median(cyc[ok] / ns[ok]) * 1000:.0f} MHz")
span_ns = full[:, 10].max() - full[:, 8].min()
sm = full[:, 9].astype(np.int64)
busy = np.zeros(sm.max() + 1)
np.add.at(busy, sm, ns)
print(f"  kernel span {span_ns / 1e6:.3f} ms; SMs used {len(np.unique(sm))}; mean resident tiles per SM {busy.mean() / span_ns:.2f}")
