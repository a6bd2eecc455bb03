#!/usr/bin/env python3
"""Small blocks (C1-sized, 768,771 bytes of text): one block is launch-latency bound (33 launches, ~0.55 ms of GPU
time, most of it idle), so throughput comes from running several contexts side by side — one context and one
host thread each, as INTEGRATION.md §5 prescribes.  Host (pinned) buffers in, host buffers out, per-thread
wall clock; every output is compared with the first thread's.

    python tools/small_blocks_probe.py [blocks_per_thread]
"""
import json
import os
import sys
import threading
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dark_b200 import saca, synth  # noqa: E402

kind, seed, n = synth.CONFIGS["c1"]
per_thread = int(sys.argv[1]) if len(sys.argv) > 1 else 64
text = torch.empty(n, dtype=torch.uint8).pin_memory()
synth.generate(kind, seed, n, out=text.numpy())
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_golden.json")))["%s:%d:%d" % (kind, seed, n)]


def worker(con, out, res, k, barrier):
    barrier.wait()
    t0 = time.perf_counter()
    origin = 0
    for _ in range(per_thread):
        origin = con.bwt_into(text.data_ptr(), n, out.data_ptr())
    res[k] = (time.perf_counter() - t0, origin, "%08x" % (zlib.crc32(out.numpy().tobytes()) & 0xFFFFFFFF))


for T in (1, 2, 4, 8, 16):
    cons = [saca.Constructor(n) for _ in range(T)]
    outs = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(T)]
    for c, o in zip(cons, outs):
        c.bwt_into(text.data_ptr(), n, o.data_ptr())  # warm-up
    res = [None] * T
    barrier = threading.Barrier(T)
    th = [threading.Thread(target=worker, args=(cons[k], outs[k], res, k, barrier)) for k in range(T)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    wall = time.perf_counter() - t0
    ok = all(r[1] == gold["origin"] and r[2] == gold["bwt_crc32"] for r in res)
    print(json.dumps({"threads": T, "blocks": T * per_thread, "block_bytes": n, "wall_ms": wall * 1e3,
                      "ms_per_block_per_thread": max(r[0] for r in res) * 1e3 / per_thread,
                      "aggregate_MBps": T * per_thread * n / 1e6 / wall, "parity": "ok" if ok else "MISMATCH"}), flush=True)
    for c in cons:
        c.close()

# ---- DISTINCT blocks text(seed=3+100k) through ONE context and ONE sort (dark_bwt_forward_many); block 0 is the C1 fixture.
# (Identical blocks are the worst case of this mode: every suffix ties with its copies down to the block end.)
for cnt in (1, 4, 16, 64, 256):
    con = saca.Constructor(cnt * n)
    ins = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(cnt)]
    for k in range(cnt):
        synth.generate(kind, seed + 100 * k, n, out=ins[k].numpy())  # seeds 100 apart: disjoint word streams (App. D: 7919 words per seed step)
    outs = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(cnt)]
    tp, bp = [t.data_ptr() for t in ins], [o.data_ptr() for o in outs]
    con.bwt_many_into(tp, [n] * cnt, bp)  # warm-up
    reps = max(1, 64 // cnt)
    t0 = time.perf_counter()
    dev = 0.0
    for _ in range(reps):
        origins = con.bwt_many_into(tp, [n] * cnt, bp)
        dev += con.stats.device_ms
    wall = time.perf_counter() - t0
    ok = origins[0] == gold["origin"] and "%08x" % (zlib.crc32(outs[0].numpy().tobytes()) & 0xFFFFFFFF) == gold["bwt_crc32"]
    print(json.dumps({"many": cnt, "calls": reps, "block_bytes": n, "wall_ms_per_call": wall * 1e3 / reps,
                      "device_ms_per_call": dev / reps, "aggregate_MBps": reps * cnt * n / 1e6 / wall,
                      "device_MBps": reps * cnt * n / 1e6 / (dev / 1e3), "rounds": con.stats.rounds,
                      "parity_block0": "ok" if ok else "MISMATCH"}), flush=True)
    con.close()
