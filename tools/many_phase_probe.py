import sys, os, json
sys.path.insert(0, "/root/repo")
import torch
from dark_b200 import saca, synth
kind, seed, n = synth.CONFIGS["c1"]
cnt = 64
con = saca.Constructor(cnt * n)
ins = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(cnt)]
for k in range(cnt):
    synth.generate(kind, seed + 100 * k, n, out=ins[k].numpy())
outs = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(cnt)]
for _ in range(2):
    con.bwt_many_into([t.data_ptr() for t in ins], [n] * cnt, [o.data_ptr() for o in outs])
st = con.stats.as_dict()
print(json.dumps({k: st[k] for k in ("device_ms", "init_ms", "sort_ms", "pass_ms", "keybuild_ms", "rerank_ms", "emit_ms", "h2d_ms", "d2h_ms", "rounds", "sort_passes", "kernel_launches", "active", "passes", "sigma", "symbols_per_key")}))
