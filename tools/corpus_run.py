#!/usr/bin/env python3
"""BASELINE.json configs[4]: a corpus of independent 256 MiB blocks (block b = mixed(seed=1000+b)), sharded
b mod G over the G GPUs of one box — one process and one context per GPU, no data-path collective.

    python tools/corpus_run.py [--blocks 128] [--block-bytes 268435456] [--batch 4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
        tools/corpus_run.py --blocks 128

Every rank generates its blocks on the host (pinned), pushes them through the pipelined batch entry
(dark_bwt_forward_batch) `--batch` at a time, and records per block: origin, CRC-32 of the BWT bytes, device time.
Rank 0 prints one JSON line: aggregate device-timed MB/s (sum of bytes / max over ranks of the summed device
times), aggregate end-to-end MB/s (wall clock of the batch calls, max over ranks) and a digest over all
(block, origin, crc) triples, which must be identical for every G.
"""
import argparse
import json
import os
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from dark_b200 import blocks as blk, saca, synth  # noqa: E402


def fixture_check(flat, n):
    """Blocks that have a committed oracle fixture (tests/golden/oracle_golden.json): origin and BWT CRC-32 must match."""
    try:
        gold = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_golden.json")))
    except Exception:
        return None
    out = {}
    for b, origin, crc in flat:
        g = gold.get("mixed:%d:%d" % (1000 + b, n))
        if g:
            out[str(b)] = "ok" if (g["origin"] == origin and g["bwt_crc32"] == "%08x" % crc) else "MISMATCH"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=128)
    ap.add_argument("--block-bytes", type=int, default=1 << 28)
    ap.add_argument("--batch", type=int, default=4)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.block_bytes
    mine = blk.shard(args.blocks, rank, world)
    con = saca.Constructor(n, device=local)
    ins = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(args.batch)]
    outs = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(args.batch)]
    records = []
    dev_ms = 0.0
    wall_s = 0.0
    if world > 1:
        dist.barrier()
    for g0 in range(0, len(mine), args.batch):
        group = mine[g0:g0 + args.batch]
        for k, b in enumerate(group):
            kind, seed, _ = synth.c5_block(b, n)
            synth.generate(kind, seed, n, out=ins[k].numpy())
        t0 = time.perf_counter()
        origins, stats = con.bwt_batch_into([ins[k].data_ptr() for k in range(len(group))], [n] * len(group),
                                            [outs[k].data_ptr() for k in range(len(group))], want_stats=True)
        wall_s += time.perf_counter() - t0
        for k, b in enumerate(group):
            dev_ms += stats[k]["device_ms"]
            records.append((b, origins[k], zlib.crc32(outs[k].numpy().tobytes()) & 0xFFFFFFFF))
    # gather the per-block records and the timings on rank 0
    all_records = [records]
    t = torch.tensor([dev_ms, wall_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, records)
        all_records = gathered
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        flat = sorted(r for part in all_records for r in part)
        assert [r[0] for r in flat] == list(range(args.blocks))
        digest = zlib.crc32(json.dumps(flat).encode()) & 0xFFFFFFFF
        total = args.blocks * n
        print(json.dumps({"workload": "C5 corpus: %d blocks of %d bytes, mixed(seed=1000+b), sharded b mod G" % (args.blocks, n),
                          "n_gpus": world, "blocks": args.blocks, "total_bytes": total,
                          "device_MBps": total / 1e6 / (float(t[0]) / 1e3), "e2e_MBps": total / 1e6 / (float(t[1]) / 1e3),
                          "max_rank_device_ms": float(t[0]), "max_rank_wall_ms": float(t[1]),
                          "digest": "%08x" % digest, "first_blocks": flat[:8], "fixture_check": fixture_check(flat, n)}), flush=True)
    con.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
