#!/usr/bin/env python3
"""Microbenchmark of the radix-pass variants (DARK_BWT_SORT_VARIANT) through dark_bwt_sort_pairs_device.
    python tools/sort_bench.py [log2_m] [variants...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dark_b200 import saca, _ffi  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 27
variants = sys.argv[2:] or ["tma", "r1"]   # "tma" = onesweep_tma.cuh (default), "r1" = round-1 kernel, <int> = its tuning variants
m = 1 << lg
g = torch.Generator(device="cuda").manual_seed(1)
keys0 = torch.randint(0, 1 << 62, (m,), dtype=torch.int64, device="cuda", generator=g)
vals0 = torch.arange(m, dtype=torch.int32, device="cuda")
ref_vals = None
if m <= (1 << 27):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rk, ri = torch.sort(keys0, stable=True)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"torch_sort_ms": e0.elapsed_time(e1)}))
    ref_vals = ri.to(torch.int32)
    del rk, ri
for v in variants:
    os.environ.pop("DARK_BWT_SORT_VARIANT", None)
    os.environ.pop("DARK_BWT_PASS_IMPL", None)
    if v == "r1":
        os.environ["DARK_BWT_PASS_IMPL"] = "0"
    elif v != "tma":
        os.environ["DARK_BWT_SORT_VARIANT"] = str(int(v))
    con = saca.Constructor(m, flags=_ffi.F_DEVICE_ONLY)   # the knobs are read when the context is created
    best = None
    ok = True
    for rep in range(3):
        k, vv = keys0.clone(), vals0.clone()
        k2, v2 = torch.empty_like(k), torch.empty_like(vv)
        in_alt, ms = con.sort_pairs_device(k.data_ptr(), vv.data_ptr(), k2.data_ptr(), v2.data_ptr(), m, 0, 64)
        torch.cuda.synchronize()
        best = ms if best is None else min(best, ms)
        if rep == 0:
            rk, rv = (k2, v2) if in_alt else (k, vv)
            ok = bool((rk[1:] >= rk[:-1]).all().item())
            if ref_vals is not None:
                ok = ok and bool(torch.equal(rv, ref_vals))
    # 8 passes of 24 B/pair + 8 B/pair histogram read
    print(json.dumps({"variant": v, "m": m, "ok": ok, "sort_ms": best, "ms_per_pass": best / 8,
                      "pass_GBs_incl_hist": (8 * 24 + 8) * m / best / 1e6}), flush=True)
    con.close()
