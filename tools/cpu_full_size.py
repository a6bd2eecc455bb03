#!/usr/bin/env python3
"""One-off CPU baseline at FULL size on this machine's host cores: the oracle (C restatement
of the reference's saca.rs + TransformIterator; the Rust original cannot be built in this image) timed on
  C2: one thread, the whole 256 MiB dna(1) block      C3: one thread, the whole 64 MiB rep17(2) block
  C5: one wave of nproc threads, each on its own 256 MiB mixed(1000+b) block
Writes gpurun_out/cpu_full_size.json (committed as profiles/r2_cpu_full_size.json).
    python tools/cpu_full_size.py [c2] [c3] [c5]"""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

which = sys.argv[1:] or ["c3", "c2", "c5"]
cores = os.cpu_count() or 1
out = {"machine": {"nproc": cores}, "command": "python tools/cpu_full_size.py " + " ".join(which),
       "what": "oracle (C restatement of saca.rs + emission), gcc -O3, full-size blocks; seconds are wall clock"}
oracle.lib()
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_golden.json")))


def one(kind, seed, n):
    t = oracle.gen(kind, seed, n)
    a = oracle.Arena(n)
    t0 = time.perf_counter()
    _, origin = a.bwt_forward(t)
    return time.perf_counter() - t0, origin


if "c3" in which:
    s, o = one("rep17", 2, 1 << 26)
    out["c3_one_thread"] = {"block_bytes": 1 << 26, "seconds": s, "MBps": (1 << 26) / 1e6 / s, "origin": o, "threads": 1,
                            "origin_matches_fixture": o == gold["rep17:2:67108864"]["origin"]}
    print(json.dumps(out["c3_one_thread"]), flush=True)
if "c2" in which:
    s, o = one("dna", 1, 1 << 28)
    out["c2_one_thread"] = {"block_bytes": 1 << 28, "seconds": s, "MBps": (1 << 28) / 1e6 / s, "origin": o, "threads": 1,
                            "origin_matches_fixture": o == gold["dna:1:268435456"]["origin"]}
    print(json.dumps(out["c2_one_thread"]), flush=True)
if "c5" in which:
    n = 1 << 28
    texts = [oracle.gen("mixed", 1000 + b, n) for b in range(cores)]
    arenas = [oracle.Arena(n) for _ in range(cores)]
    res = [None] * cores

    def work(i):
        t0 = time.perf_counter()
        _, o = arenas[i].bwt_forward(texts[i])
        res[i] = (time.perf_counter() - t0, o)

    ths = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    wall = time.perf_counter() - t0
    out["c5_one_wave"] = {"block_bytes": n, "threads": cores, "blocks": cores, "seconds": wall, "MBps": cores * n / 1e6 / wall,
                          "per_block_seconds_min_max": [min(r[0] for r in res), max(r[0] for r in res)],
                          "origins_match_fixtures": res[0][1] == gold["mixed:1000:268435456"]["origin"] and (cores < 2 or res[1][1] == gold["mixed:1001:268435456"]["origin"]),
                          "corpus_extrapolation": "128 blocks = %d waves -> %.0f s for the 32 GiB corpus on this machine" % (-(-128 // cores), -(-128 // cores) * wall)}
    print(json.dumps(out["c5_one_wave"]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "cpu_full_size.json"), "w"), indent=1)
