"""dark_b200 — B200-native forward Burrows–Wheeler transform for the `dark` compressor.

One hot path only: the replacement of kvark/dark's `saca::Constructor::compute` +
`compress::bwt::TransformIterator` (reference: src/saca.rs:344-384, src/block/dc.rs:45-50,
src/block/raw.rs:39-44) by hand-written sm_100a CUDA kernels behind a C ABI
(include/dark_bwt.h, built into dark_b200/lib/libdark_bwt.so).

    from dark_b200 import saca
    con = saca.Constructor(len(block))      # Constructor::new
    bwt, origin = con.bwt(block)            # compute + TransformIterator at the two call sites
    sa = con.compute(block)                 # the suffix array itself, as the reference returns it

There is no CPU fallback: importing works anywhere, but creating a Constructor without the
built library or without a CUDA device raises.
"""
from . import saca, synth, blocks  # noqa: F401
from ._ffi import Stats, DarkBwtError, lib_path, build_native  # noqa: F401

__all__ = ["saca", "synth", "blocks", "Stats", "DarkBwtError", "lib_path", "build_native"]
