import sys; sys.path.insert(0, '.')
import numpy as np
from oracle import world as W
from realtime_yukarin_b200.engine import default_engine
np.set_printoptions(linewidth=200, precision=6, suppress=True)
eng = default_engine()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
nb = 513
ref = W.RealtimeSynthesizer(24000, 5.0, 1024, 1024)
sid = eng.synth_create(24000, 5.0, 1024, 1024)
seen = 0
for k in range(40):
    # random voiced / unvoiced segments with float32 f0
    f0 = np.zeros(60, np.float32)
    a = 0
    while a < 60:
        L = int(rng.integers(5, 30))
        if rng.random() < 0.6:
            f0[a:a + L] = (rng.uniform(150, 450) + np.cumsum(rng.standard_normal(min(L, 60 - a)))).astype(np.float32)
        a += L
    sp = np.exp(-9 + rng.standard_normal((60, nb))).astype(np.float32)
    ap = rng.uniform(0.01, 0.9, (60, nb)).astype(np.float32)
    yr = ref.decode(f0.astype(np.float64), sp, ap)
    yg = eng.synth_decode(sid, f0.astype(np.float64), sp, ap)
    ir, tr, vr = ref.pulses()
    ig, tg, vg, st = eng.debug_synth_pulses(sid, 0)
    n = min(len(ir), len(ig))
    same = np.array_equal(ir[:n], ig[:n]) and len(ir) == len(ig)
    msg = f'chunk {k}: len {len(yg)} vs {len(yr)} pulses {len(ig)} vs {len(ir)} same={same}'
    if len(yr) == len(yg) and len(yr): msg += f' rmse {np.sqrt(np.mean((yr - yg) ** 2)):.2e}'
    print(msg)
    if not same:
        d = np.where(ir[:n] != ig[:n])[0]
        print('  first diffs at', d[:10], 'oracle', ir[d[:10]], 'gpu', ig[d[:10]], 'vuv', vr[d[:10]], 'times', tr[d[:5]], tg[d[:5]])
        break
