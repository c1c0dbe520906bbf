"""Diagnostics (needs a libryk.so built with RYK_NVCC_EXTRA=-DRYK_TC_TIMELINE): per-CTA phase timeline of k_conv_tc."""
import os, sys; sys.path.insert(0, '.')
import numpy as np
from realtime_yukarin_b200.engine import default_engine
eng = default_engine()
rng = np.random.default_rng(0)
LAYERS = {'c1': (0, 384, 512, 64, 0, 128), 'd6': (1, 192, 256, 128, 128, 64), 'd4': (1, 48, 64, 512, 512, 256), 'c3': (0, 96, 128, 256, 0, 512),
          'c4': (0, 48, 64, 512, 0, 512), 'c5': (0, 24, 32, 512, 0, 512), 'c6': (0, 12, 16, 512, 0, 512), 'c7': (0, 6, 8, 512, 0, 512),
          'd0': (1, 3, 4, 512, 0, 512), 'd1': (1, 6, 8, 512, 512, 512), 'd2': (1, 12, 16, 512, 512, 512), 'd3': (1, 24, 32, 512, 512, 512)}
for name in sys.argv[1:]:
    tr, H, W, C0, C1, Cout = LAYERS[name]
    in0 = rng.standard_normal((1, H, W, C0)).astype(np.float32)
    in1 = rng.standard_normal((1, H, W, C1)).astype(np.float32) if C1 else None
    Cin = C0 + C1
    Wt = (rng.standard_normal((Cin, Cout, 4, 4) if tr else (Cout, Cin, 4, 4)) / np.sqrt(Cin * 4)).astype(np.float32)
    path = f'gpurun_out/tl_{name}.txt'
    os.environ['RYK_TC_TIMELINE_FILE'] = path
    out, ms = eng.test_conv_layer(in0, in1, Wt, np.ones(Cout, np.float32), np.zeros(Cout, np.float32), tr, 4, 2, 1, 1, use_tc=1, repeat=2)
    rows = np.loadtxt(path, dtype=np.float64, comments='#')
    t = rows[:, :7]; sm = rows[:, 9].astype(int)
    t0 = t[:, 0].min()
    ph = ['setup(alloc+sync)', 'first full', 'mma loop issue', 'accum ready', 'epilogue', 'dealloc']
    d = np.diff(t, axis=1)
    hdr = open(path).readline().strip()
    print(f'== {name}: {hdr} {len(rows)} CTAs, kernel span {(t[:, 6].max() - t0) / 1e3:.1f} us (timed {ms * 1e3:.1f} us incl. sync)')
    for i, p_ in enumerate(ph):
        print(f'   {p_:20s} mean {d[:, i].mean() / 1e3:7.2f} us  p10 {np.percentile(d[:, i], 10) / 1e3:7.2f}  p90 {np.percentile(d[:, i], 90) / 1e3:7.2f}')
    life = (t[:, 6] - t[:, 0]) / 1e3
    print(f'   CTA lifetime mean {life.mean():.2f} us; start times: first wave {np.sort(t[:, 0] - t0)[:3] / 1e3}, CTAs per SM max {np.bincount(sm).max()} min {np.bincount(sm, minlength=148).min()}')
    # concurrency on the busiest SM
    b = np.bincount(sm).argmax(); idx = np.where(sm == b)[0]; o = np.argsort(t[idx, 0])
    print('   busiest SM timeline (start, end us):', [(round((t[i, 0] - t0) / 1e3, 1), round((t[i, 6] - t0) / 1e3, 1)) for i in idx[o]][:12])
