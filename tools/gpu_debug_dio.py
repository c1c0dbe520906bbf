import sys; sys.path.insert(0, '.')
import numpy as np
from oracle import world as W
from realtime_yukarin_b200 import synthetic
from realtime_yukarin_b200.engine import default_engine
eng = default_engine()
np.set_printoptions(linewidth=200, precision=4, suppress=True)
for stream, seconds in ((0, 0.3), (1, 1.0), (2, 0.1)):
    x = synthetic.synthetic_speech(seconds, stream)
    f0o, t, cando, scoreo = W.dio(x.astype(np.float64), 24000, 5.0, 71.0, 800.0, debug=True)
    f0ro = W.stonemask(x.astype(np.float64), 24000, t, f0o)
    f0g, _ = eng.world_f0(x, 24000, 5.0, 71.0, 800.0)
    f0raw, cand, score, counts = eng.debug_dio(len(x), 24000, 5.0, 71.0, 800.0)
    print('stream', stream, 'frames', len(f0o), 'counts', counts.tolist())
    print(' raw dio max abs diff', np.abs(f0raw - f0o).max(), 'stonemask diff', np.abs(f0g - f0ro).max())
    dc = np.abs(cand - cando); print(' cand diff max', dc.max(), 'where', np.argwhere(dc > 1e-6)[:10].tolist())
    ds = np.abs(score - scoreo) / (np.abs(scoreo) + 1e-9); print(' score rel diff max', ds.max())
    bad = np.where(np.abs(f0raw - f0o) > 1e-6)[0]
    if len(bad):
        print(' bad frames', bad[:20], '\n gpu', f0raw[bad[:10]], '\n orc', f0o[bad[:10]])
        for b in range(cand.shape[0]):
            print('  band', b, 'gpu', cand[b, :10], '\n         orc', cando[b, :10])
    bad = np.where(np.abs(f0g - f0ro) > 1e-6)[0]
    print(' stonemask bad frames', bad[:20], f0g[bad[:10]], f0ro[bad[:10]], 'dio there', f0raw[bad[:10]])
