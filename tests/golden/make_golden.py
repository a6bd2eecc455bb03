#!/usr/bin/env python3
"""Generate the committed golden fixtures with the CPU oracle (oracle/oracle_cli).

The reference (Rust) cannot run in this image, so the fixtures are outputs of
the oracle — the C restatement of /root/reference/src/saca.rs + TransformIterator,
itself pinned by the reference's known-answer test (saca.rs:409-413) — on the
SURVEY.md App. D generators.  Each record holds n, origin, CRC-32 of the text,
of the BWT bytes and of the little-endian SA, the SA-IS recursion trace and the
§8(d) LCP profile (m_r, B_alg).

    python tests/golden/make_golden.py [small|full|c4|c5more]

`small` (seconds) and `full` (C2/C3/C5-block shapes, minutes of CPU) write
tests/golden/oracle_golden.json; `c4` (2 GiB block, ~15 min, ~25 GB RAM) adds
the C4 record.
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CLI = os.path.join(ROOT, "oracle", "build", "oracle_cli")
OUT = os.path.join(HERE, "oracle_golden.json")

SMALL = [
    ("text", 3, 768771),       # C1
    ("dna", 1, 1 << 16),
    ("dna", 1, 1 << 20),
    ("dna", 1, 1 << 24),
    ("dna", 7, (1 << 20) + 12345),
    ("rep17", 2, 1 << 16),
    ("rep17", 2, 1 << 20),
    ("rep17", 2, 1 << 22),
    ("mixed", 4, 1 << 20),
    ("mixed", 4, 1 << 24),
    ("mixed", 1000, 1 << 22),
    ("text", 5, 100003),
]
FULL = [
    ("dna", 1, 1 << 28),       # C2
    ("rep17", 2, 1 << 26),     # C3
    ("mixed", 4, 1 << 28),
    ("mixed", 1000, 1 << 28),  # C5 block 0
    ("mixed", 1001, 1 << 28),  # C5 block 1
]
C4 = [("mixed", 4, 1 << 31)]
C5MORE = [("mixed", 1000 + b, 1 << 28) for b in range(2, 8)]   # C5 blocks 2..7: what bench.py --gpus 2 transforms besides 0 and 1


def key(kind, seed, n):
    return f"{kind}:{seed}:{n}"


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    cases = {"small": SMALL, "full": FULL, "c4": C4, "c5more": C5MORE}[which]
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    gold = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            gold = json.load(f)
    for kind, seed, n in cases:
        out = subprocess.check_output([CLI, kind, str(seed), str(n), "--profile"], text=True)
        rec = json.loads(out)
        for k in ("saca_s", "emit_s", "mb_per_s"):   # timings are not golden
            rec.pop(k, None)
        gold[key(kind, seed, n)] = rec
        print(key(kind, seed, n), "origin", rec["origin"], "bwt", rec["bwt_crc32"], flush=True)
        with open(OUT, "w") as f:
            json.dump(gold, f, indent=1, sort_keys=True)
            f.write("\n")


if __name__ == "__main__":
    main()
