#!/usr/bin/env python3
"""Timing of the GPU inverse BWT on the config shapes (device-resident)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dark_b200 import saca, synth, _ffi
for name in sys.argv[1:] or ["c2", "c3", "c5"]:
    kind, seed, n = synth.CONFIGS.get(name, ("mixed", 1000, 1 << 28))
    t = synth.generate(kind, seed, n)
    con = saca.Constructor(n, flags=_ffi.F_DEVICE_ONLY)
    dt = torch.from_numpy(t).cuda()
    db = torch.empty(n, dtype=torch.uint8, device="cuda")
    origin = con.bwt_device(dt.data_ptr(), n, db.data_ptr())
    dback = torch.zeros(n, dtype=torch.uint8, device="cuda")
    best = min(con.inverse_device(db.data_ptr(), n, origin, dback.data_ptr()) for _ in range(3))
    print(json.dumps({"workload": name, "n": n, "inverse_ms": best, "inverse_GBps": n / best / 1e6, "ok": bool(torch.equal(dback, dt)),
                      "forward_ms": con.stats.device_ms}))
    con.close()
