#!/usr/bin/env python3
"""One isolated host-to-host call of dark_bwt_forward on C2 from pinned and from pageable (malloc'ed) buffers.
    DARK_BWT_HOST_THREADS=<t> DARK_BWT_HOST_CHUNK_MB=<c> python tools/pageable_probe.py [reps]
Prints one JSON line: best-of-reps milliseconds for both, and their ratio (bench.py reports the same two numbers)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from dark_b200 import saca, synth  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
kind, seed, n = synth.CONFIGS["c2"]
con = saca.Constructor(n)
h_text = torch.empty(n, dtype=torch.uint8).pin_memory()
h_bwt = torch.empty(n, dtype=torch.uint8).pin_memory()
synth.generate(kind, seed, n, out=h_text.numpy())
p_text = np.array(h_text.numpy(), copy=True)
p_bwt = np.zeros(n, dtype=np.uint8)


def best(tp, bp):
    out = []
    for _ in range(reps + 1):
        t0 = time.perf_counter()
        origin = con.bwt_into(tp, n, bp)
        out.append(time.perf_counter() - t0)
    return min(out[1:]) * 1e3, origin


pin_ms, o1 = best(h_text.data_ptr(), h_bwt.data_ptr())
page_ms, o2 = best(p_text.ctypes.data, p_bwt.ctypes.data)
assert o1 == o2 and np.array_equal(p_bwt, h_bwt.numpy())
print(json.dumps({"host_threads": os.environ.get("DARK_BWT_HOST_THREADS", "default"),
                  "chunk_mb": os.environ.get("DARK_BWT_HOST_CHUNK_MB", "default"),
                  "pinned_ms": round(pin_ms, 2), "pageable_ms": round(page_ms, 2), "ratio": round(page_ms / pin_ms, 3),
                  "device_ms": round(con.stats.device_ms, 2), "nproc": os.cpu_count()}))
con.close()
