#!/usr/bin/env python3
"""Small forward transforms covering every kernel and forced code path, for compute-sanitizer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import oracle
from dark_b200 import saca, synth
cases = [("mixed", 4, 70001), ("dna", 1, 40003), ("rep17", 2, 50000), ("text", 3, 30011), ("dna", 9, 2), ("text", 1, 9)]
for kind, seed, n in cases:
    t = synth.generate(kind, seed, n)
    with saca.Constructor(n) as c:
        b, o, s = c.bwt_and_sa(t)
    bo, oo, so = oracle.bwt_forward(t, want_sa=True)
    assert o == oo and np.array_equal(b, bo) and np.array_equal(s, so), (kind, seed, n)
with saca.Constructor(1 << 16) as c:
    res = c.bwt_blocks([synth.generate("mixed", 5, 50001), synth.generate("dna", 6, 65536), synth.generate("text", 7, 777)])
print("sanitize probe ok", len(res))
