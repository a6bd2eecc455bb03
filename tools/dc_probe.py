#!/usr/bin/env python3
"""Device time of the GPU distance coding + MTF stage on the BWT of a named workload (device-resident).
    python tools/dc_probe.py <c1|c2|c3|c5> [reps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dark_b200 import saca, synth, _ffi  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c5"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
kind, seed, n = synth.CONFIGS.get(name, ("mixed", 1000, 1 << 28))
text = synth.generate(kind, seed, n)
con = saca.Constructor(n, flags=_ffi.F_DEVICE_ONLY)
dt = torch.from_numpy(text).cuda()
db = torch.empty(n, dtype=torch.uint8, device="cuda")
origin = con.bwt_device(dt.data_ptr(), n, db.data_ptr())
fwd_ms = con.stats.device_ms
for _ in range(reps):
    info = con.dc_encode_device(db.data_ptr(), n)
    print(json.dumps({"workload": name, "n": n, "forward_ms": fwd_ms, "dc_ms": info.device_ms, "dc_GBps": n / 1e6 / info.device_ms,
                      "runs": int(info.num_items), "runs_per_byte": int(info.num_items) / n, "num_unique": int(info.num_unique)}))
con.close()
