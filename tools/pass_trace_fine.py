#!/usr/bin/env python3
"""Where does a tile's wait for its prefix come from?  (tuning build -DDARK_TUNE_TRACE -DDARK_TUNE_TRACE_FINE)
publish(t) = globaltimer when tile t's counts were written, got(t) = when its prefix arrived (both by thread 255)."""
import sys
sys.argv = [sys.argv[0], '27']
exec(open('tools/pass_trace2.py').read().split("names = [")[0])
import numpy as np
full = trace.cpu().numpy().reshape(-1, 12)
full = full[:m // 4096]
pub = full[:, 1].astype(np.float64)
got = full[:, 2].astype(np.float64)
ok = (pub > 0) & (got > 0)
print("tiles with both stamps:", ok.sum(), "of", len(full), " pass ms", ms)
t0 = pub[ok].min()
pub -= t0; got -= t0
runmax = np.maximum.accumulate(np.where(pub > 0, pub, 0))
mid = slice(len(full) // 4, 3 * len(full) // 4)
sel = ok[mid]
def st(a): return "median %.0f mean %.0f p90 %.0f p99 %.0f ns" % (np.median(a), a.mean(), np.percentile(a, 90), np.percentile(a, 99))
print("own wait got(t)-pub(t):          ", st((got - pub)[mid][sel]))
print("in-order skew runmax(t)-pub(t):  ", st((runmax - pub)[mid][sel]))
print("scanner lag got(t)-runmax(t):    ", st((got - runmax)[mid][sel]))
d = np.diff(pub[mid][sel])
print("publish interval per tile:       ", st(d), " (negative = out of order: %.1f%%)" % (100.0 * (d < 0).mean()))
