import sys; sys.path.insert(0, '.')
import numpy as np
from realtime_yukarin_b200.engine import default_engine
np.set_printoptions(linewidth=200, precision=17, suppress=False)
eng = default_engine()
rng = np.random.default_rng(0)
nb = 513
sid = eng.synth_create(24000, 5.0, 1024, 1024)
f0 = np.zeros(60, np.float32)
a = 0
while a < 60:
    L = int(rng.integers(5, 30))
    if rng.random() < 0.6:
        f0[a:a + L] = (rng.uniform(150, 450) + np.cumsum(rng.standard_normal(min(L, 60 - a)))).astype(np.float32)
    a += L
sp = np.exp(-9 + rng.standard_normal((60, nb))).astype(np.float32)
ap = rng.uniform(0.01, 0.9, (60, nb)).astype(np.float32)
eng.synth_decode(sid, f0.astype(np.float64), sp, ap)
ns = 7080
g_if0, g_vuv, g_tp = eng.debug_synth_timebase(sid, ns)
# numpy restatement of the oracle's first-call time base (hf = 0)
fp, fs = 5.0 / 1000.0, 24000
f0d = f0.astype(np.float64)
ct = np.arange(60, dtype=np.float64) * fp
ta = np.arange(ns, dtype=np.float64) / float(fs)
k = np.clip(np.searchsorted(ct, ta, side='right'), 1, 59)
s = (ta - ct[k - 1]) / (ct[k] - ct[k - 1])
cv = (f0d != 0).astype(np.float64)
fi = f0d[k - 1] + s * (f0d[k] - f0d[k - 1])
vi = cv[k - 1] + s * (cv[k] - cv[k - 1])
vi = (vi > 0.5).astype(np.float64)
fi = np.where(vi == 0, 500.0, fi)
print('f0 frames', f0)
d = np.where(g_vuv != vi)[0]
print('vuv mismatches at samples', d[:20], 's there', s[d[:10]])
d2 = np.where(g_if0 != fi)[0]
print('if0 mismatches', len(d2), d2[:20])
if len(d2): print(' gpu', g_if0[d2[:5]], '\n ref', fi[d2[:5]], '\n s', s[d2[:5]], 'k', k[d2[:5]])
