"""Diagnostics (needs RYK_LIB=<libryk built with RYK_NVCC_EXTRA=-DRYK_TC_TIMELINE>): per-CTA phase timeline of the pair kernel k_conv_tc2."""
import os, sys; sys.path.insert(0, '.')
import numpy as np
from realtime_yukarin_b200.engine import default_engine
eng = default_engine()
rng = np.random.default_rng(0)
LAYERS = {'c1': (0, 384, 512, 64, 0, 128), 'c2': (0, 192, 256, 128, 0, 256), 'd6': (1, 192, 256, 128, 128, 64), 'd5': (1, 96, 128, 256, 256, 128),
          'd4': (1, 48, 64, 512, 512, 256)}
for name in sys.argv[1:]:
    tr, H, W, C0, C1, Cout = LAYERS[name]
    in0 = rng.standard_normal((1, H, W, C0)).astype(np.float32)
    in1 = rng.standard_normal((1, H, W, C1)).astype(np.float32) if C1 else None
    Cin = C0 + C1
    Wt = (rng.standard_normal((Cin, Cout, 4, 4) if tr else (Cout, Cin, 4, 4)) / np.sqrt(Cin * 4)).astype(np.float32)
    path = f'gpurun_out/tl2_{name}.txt'
    os.environ['RYK_TC_TIMELINE_FILE'] = path
    out, ms = eng.test_conv_layer(in0, in1, Wt, np.ones(Cout, np.float32), np.zeros(Cout, np.float32), tr, 4, 2, 1, 1, use_tc=1, repeat=2)
    rows = np.loadtxt(path, dtype=np.float64, comments='#')
    hdr = open(path).readline().strip()
    t = rows[:, :8]; rank = rows[:, 8].astype(int); sm = rows[:, 9].astype(int)
    t0 = t[:, 0].min()
    L = rank == 0
    print(f'== {name}: {hdr}; {len(rows)} CTAs, kernel span {(t[:, 6].max() - t0) / 1e3:.1f} us (graph-timed {ms * 1e3:.1f} us)')
    def stat(label, d):
        print(f'   {label:34s} mean {d.mean() / 1e3:7.2f} us  p10 {np.percentile(d, 10) / 1e3:7.2f}  p90 {np.percentile(d, 90) / 1e3:7.2f}')
    stat('start -> cluster barrier passed', t[:, 7] - t[:, 0])
    stat('grid dependency wait', t[:, 1] - t[:, 7])
    stat('leader: ready -> first A full', t[L, 2] - t[L, 1])
    stat('leader: MMA loop (issue)', t[L, 3] - t[L, 2])
    stat('ready -> accumulators complete', t[:, 4] - t[:, 1])
    stat('epilogue (ld + stage + TMA store)', t[:, 5] - t[:, 4])
    stat('final cluster barrier + dealloc', t[:, 6] - t[:, 5])
    stat('CTA lifetime', t[:, 6] - t[:, 0])
    starts = np.sort(t[:, 0] - t0) / 1e3
    print(f'   CTA start times us: p50 {np.percentile(starts, 50):.1f} p90 {np.percentile(starts, 90):.1f} max {starts.max():.1f}; CTAs per SM max {np.bincount(sm).max()} min {np.bincount(sm, minlength=148).min()}')
    b = np.bincount(sm).argmax(); idx = np.where(sm == b)[0]; o = np.argsort(t[idx, 0])
    print('   busiest SM (start, end us):', [(round((t[i, 0] - t0) / 1e3, 1), round((t[i, 6] - t0) / 1e3, 1)) for i in idx[o]][:10])
