#!/usr/bin/env python3
"""bench.py — BWT-forward throughput of the B200-native path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload c2|c3|c5|c1|c4]

A "step" is one forward BWT (suffix sort + BWT emission + origin) of one synthetic block per GPU.
Default workload = BASELINE.json configs[1]: a 256 MiB block of DNA-like 4-symbol text (C2).
N > 1 (launched by torchrun, one rank per GPU): every rank transforms its own block of the same
shape — independent blocks, no data-path collective — weak scaling.

JSON line (rank 0): value = total input MB / max-over-ranks device time; `e2e` = the same metric
through the host-buffer C-ABI call (pinned host -> H2D -> transform -> D2H) timed by wall clock;
`roofline` = the radix-pass kernel (dominant) against the measured HBM copy bandwidth;
`cpu_baseline` = the oracle (C restatement of the reference's saca.rs + emission) on one host core.

`--impl reference` times that CPU restatement on all host cores (the reference is Rust and cannot
be compiled in this image), same metric/config/unit.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "bwt_forward_throughput"
UNIT = "MB/s"

WORKLOADS = {
    # name: (kind, base_seed, n, description, B_alg per input byte (SURVEY §8d / tests/golden), sum_m/N)
    "c1": ("text", 3, 768771, "C1: 768 KB English-like text block text(seed=3)", 366.437),
    "c2": ("dna", 1, 1 << 28, "C2: 256 MiB synthetic DNA-like 4-symbol block dna(seed=1+rank)", 497.5),
    "c3": ("rep17", 2, 1 << 26, "C3: 64 MiB period-17 block with sparse mutations rep17(seed=2+rank)", 2342.0),
    "c4": ("mixed", 4, 1 << 31, "C4: 2 GiB mixed binary/text block mixed(seed=4+rank)", 461.4),
    "c5": ("mixed", 1000, 1 << 28, "C5 block: 256 MiB mixed binary/text block mixed(seed=1000+rank)", 445.8),
}


def measured_pass_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one full-size radix-pass launch (ncu --set full capture
    committed under profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_ncu_pass_traffic.json")) as f:
            t = json.load(f)
        return {"bytes_per_launch": t["dram__bytes_read.sum"] + t["dram__bytes_write.sum"], "pairs_per_launch": t["pairs_per_launch"],
                "algorithmic_bytes_per_launch": t["algorithmic_bytes_per_launch"]}
    except Exception:
        return None


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(len(r) >= 7 and r[3 + k].lower() == "active" for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_oracle_rate(kind, seed, sample_n, threads, steps, warmup):
    """MB/s of the oracle (saca.rs restatement + emission) with `threads` host threads, each on
    its own block of sample_n bytes per step.  ctypes drops the GIL during the C call."""
    import numpy as np
    import oracle
    oracle.lib()
    texts = [oracle.gen(kind, seed + t, sample_n) for t in range(threads)]
    arenas = [oracle.Arena(sample_n) for _ in range(threads)]
    per_step = []

    def work(t):
        arenas[t].bwt_forward(texts[t])

    for it in range(warmup + steps):
        ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        dt = time.perf_counter() - t0
        if it >= warmup:
            per_step.append(dt)
    total = sum(per_step)
    return threads * sample_n * len(per_step) / 1e6 / total, total / len(per_step)


def run_reference(args, rank):
    """Reference arm: the reference's CPU algorithm (oracle restatement; the Rust original cannot be
    built here) on all host cores.  Rank 0 only."""
    if rank != 0:
        return
    kind, seed, n, desc, _ = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    total_steps = args.steps + args.warmup
    sample_n = min(n, 1 << 24 if total_steps <= 16 else 1 << 22)
    rate, step_s = cpu_oracle_rate(kind, seed, sample_n, cores, args.steps, args.warmup)
    sample = f"{cores} threads x one {sample_n}-byte {kind} block per step (prefix-shaped sample of the {n}-byte block)"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": desc, "block_bytes": n, "sample_bytes": sample_n,
                   "note": "CPU restatement of the reference's saca.rs + TransformIterator (oracle port; Rust reference "
                           "not buildable in this image), all host cores"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_native(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist
    from dark_b200 import saca, synth, blocks

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the forward BWT has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    kind, base_seed, n, desc, balg_per_byte = WORKLOADS[args.workload]
    seed = base_seed + rank
    # pinned host input/output (the caller's buffers of the host entry point)
    h_text = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_bwt = torch.empty(n, dtype=torch.uint8).pin_memory()
    synth.generate(kind, seed, n, out=h_text.numpy())
    con = saca.Constructor(n, device=local_rank)
    d_text = h_text.to(dev, non_blocking=False)
    d_bwt = torch.empty(n, dtype=torch.uint8, device=dev)
    stream = torch.cuda.ExternalStream(con.stream, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident steps
    for _ in range(args.warmup):
        origin = con.bwt_device(d_text.data_ptr(), n, d_bwt.data_ptr())
    # B_alg of THIS block, measured on the GPU from the suffix array (dark_bwt_lcp_profile_device), outside the timed
    # region; it must agree with the oracle profiler's committed figure for the workload
    balg_live = None
    if rank == 0 and n <= (1 << 29):
        d_sa = torch.empty(n, dtype=torch.int32, device=dev)
        con.bwt_device(d_text.data_ptr(), n, d_bwt.data_ptr(), d_sa.data_ptr())
        prof = con.lcp_profile_device(d_text.data_ptr(), n, d_sa.data_ptr())
        balg_live = prof["b_alg"] / n
        del d_sa
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    agg = {"pass_ms": 0.0, "sorted": 0, "passes": 0, "device_ms": 0.0, "init_ms": 0.0, "sort_ms": 0.0, "keybuild_ms": 0.0,
           "rerank_ms": 0.0, "emit_ms": 0.0}
    ev0.record(stream)
    for _ in range(args.steps):
        origin = con.bwt_device(d_text.data_ptr(), n, d_bwt.data_ptr())
        s = con.stats
        launches += s.kernel_launches
        agg["pass_ms"] += s.pass_ms
        agg["sorted"] += s.sorted_elements
        agg["passes"] += s.sort_passes
        for k in ("device_ms", "init_ms", "sort_ms", "keybuild_ms", "rerank_ms", "emit_ms"):
            agg[k] += getattr(s, k)
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = ev0.elapsed_time(ev1)
    stats = con.stats.as_dict()

    # ---- end to end through the host-buffer C-ABI: pinned host blocks in, BWT bytes + origin out, every
    # copy inside the timed region.  Blocks go through the pipelined batch entry (dark_bwt_forward_batch:
    # copy-in of block k+1 / copy-out of block k-1 overlap the transform of block k), which is how a
    # corpus of independent blocks is fed; the latency of one isolated call is reported beside it.
    e2e_steps = max(4, min(args.steps, 32))  # one pipelined batch of K blocks (the first copy-in and the last copy-out are exposed)
    h_bwt2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    outs = [h_bwt.data_ptr() if i % 2 == 0 else h_bwt2.data_ptr() for i in range(e2e_steps)]
    con.bwt_into(h_text.data_ptr(), n, h_bwt.data_ptr())   # warm
    barrier()
    t0 = time.perf_counter()
    origin_h = con.bwt_into(h_text.data_ptr(), n, h_bwt.data_ptr())
    single_call_s = time.perf_counter() - t0
    assert origin_h == origin
    assert torch.equal(h_bwt[: 1 << 20], d_bwt[: 1 << 20].cpu())
    barrier()
    t0 = time.perf_counter()
    origins = con.bwt_batch_into([h_text.data_ptr()] * e2e_steps, [n] * e2e_steps, outs)
    e2e_s = time.perf_counter() - t0
    assert all(o == origin for o in origins)
    assert torch.equal(h_bwt2[-(1 << 20):], d_bwt[-(1 << 20):].cpu())

    # max over ranks of the time, sum over ranks of the work (dark_b200.blocks.aggregate; gloo-tested)
    ms_total, total_bytes = blocks.aggregate(ms_total, n * args.steps)
    e2e_ms, e2e_bytes = blocks.aggregate(e2e_s * 1e3, n * e2e_steps)
    _, launches = blocks.aggregate(0.0, launches)

    if rank == 0:
        ms_per_step = ms_total / args.steps
        value = total_bytes / 1e6 / (ms_total / 1e3)
        e2e_value = e2e_bytes / 1e6 / (e2e_ms / 1e3)
        peak, peak_src = measured_peak_gbs()
        pass_gbs = 24.0 * agg["sorted"] / (agg["pass_ms"] / 1e3) / 1e9 if agg["pass_ms"] > 0 else 0.0
        path_gbs = balg_per_byte * n / (ms_per_step / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": desc, "block_bytes": n, "blocks_per_step": world, "sharding": "independent blocks, no collective",
                       "l2": "inputs larger than L2 (text %d MiB, working set ~%.1f GiB): no flush between steps" % (n >> 20, 46.5 * n / 2**30),
                       "origin": origin, "sigma": stats["sigma"], "symbols_per_key": stats["symbols_per_key"],
                       "rounds": stats["rounds"], "active_per_round": stats["active"], "passes_per_round": stats["passes"]},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n, "d2h_bytes_per_step": n + 8,
                    "steps": e2e_steps, "timer": "host wall clock around dark_bwt_forward_batch (pipelined copies) on pinned host buffers",
                    "single_call_ms": single_call_s * 1e3, "single_call_value": n / 1e6 / single_call_s},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_onesweep_pass (radix pass: 12 B read + 12 B written per pair)",
                         "achieved": pass_gbs, "peak": peak, "unit": "GB/s", "frac": pass_gbs / peak, "peak_source": peak_src,
                         "launches_per_step": agg["passes"] / args.steps,
                         "avg_launch_ms": agg["pass_ms"] / max(1, agg["passes"]),
                         "traffic": (measured_pass_traffic() or {}).get("bytes_per_launch"),
                         "traffic_detail": measured_pass_traffic(),
                         "path_b_alg_per_byte": balg_per_byte, "path_b_alg_per_byte_measured_on_gpu": balg_live, "path_achieved": path_gbs, "path_frac": path_gbs / peak,
                         "path_frac_of_8TBs": path_gbs / 8000.0},
            "phases_ms_per_step": {k: agg[k] / args.steps for k in ("device_ms", "init_ms", "sort_ms", "keybuild_ms",
                                                                    "rerank_ms", "emit_ms")},
        }
        if world == 1 and not args.no_cpu_baseline:
            sample_n = min(n, 1 << 25)
            rate, step_s = cpu_oracle_rate(kind, base_seed, sample_n, 1, 1, 0)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"one {sample_n}-byte {kind}(seed={base_seed}) block, 1 thread, {step_s:.1f} s "
                                              f"(oracle = C restatement of saca.rs + emission)"}
        print(json.dumps(line), flush=True)
    con.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py: --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    run_native(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
