"""Diagnostics: per-CTA event timeline of k_conv_halo (needs the -DRYK_TC_TIMELINE build: RYK_LIB=realtime_yukarin_b200/csrc/libryk_tl.so).
usage: RYK_LIB=... python tools/gpu_tc3_timeline.py c2 d4 ...   -> prints, per layer, the gaps between events of CTA 0 (clocks of its SM)."""
import os, sys; sys.path.insert(0, '.')
import numpy as np
from realtime_yukarin_b200.engine import default_engine
eng = default_engine()
rng = np.random.default_rng(0)
LAYERS = {'c1': (0, 384, 512, 64, 0, 128), 'c2': (0, 192, 256, 128, 0, 256), 'c3': (0, 96, 128, 256, 0, 512), 'd3': (1, 24, 32, 512, 512, 512),
          'd4': (1, 48, 64, 512, 512, 256), 'd5': (1, 96, 128, 256, 256, 128), 'd6': (1, 192, 256, 128, 128, 64)}
NAMES = {0: 'start', 1: 'pre-pdl', 2: 'post-pdl', 10: 'A', 11: 'B', 20: 'wA', 21: 'gotA', 22: 'gotB', 23: 'iss', 29: 'tile', 30: 'wT', 31: 'gotT', 32: 'wS', 33: 'gotS', 34: 'st', 39: 'end'}
for name in sys.argv[1:]:
    tr, H, W, C0, C1, Cout = LAYERS[name]
    in0 = rng.standard_normal((1, H, W, C0)).astype(np.float32)
    in1 = rng.standard_normal((1, H, W, C1)).astype(np.float32) if C1 else None
    Cin = C0 + C1
    Wt = (rng.standard_normal((Cin, Cout, 4, 4) if tr else (Cout, Cin, 4, 4)) / np.sqrt(Cin * 4)).astype(np.float32)
    path = f'gpurun_out/tl3_{name}.txt'
    os.environ['RYK_TC_TIMELINE_FILE'] = path
    eng.test_conv_layer(in0, in1, Wt, np.ones(Cout, np.float32), np.zeros(Cout, np.float32), tr, 4, 2, 1, 1, use_tc=1, repeat=0)   # warm (L2, tensormaps)
    eng.test_conv_layer(in0, in1, Wt, np.ones(Cout, np.float32), np.zeros(Cout, np.float32), tr, 4, 2, 1, 1, use_tc=1, repeat=0)
    lines = open(path).read().strip().split('\n')
    print('==', name, lines[0])
    rows = {}
    for ln in lines[1:]:
        f = ln.split()
        rows[(int(f[0]), int(f[1]))] = [(int(x.split(':')[0]), int(x.split(':')[1])) for x in f[2:]]
    for cta in (0, 77):
        if (cta, 0) not in rows:
            continue
        t0 = rows[(cta, 0)][0][1]
        for role, rn in ((0, 'producer'), (1, 'mma'), (2, 'epilogue')):
            ev = rows.get((cta, role), [])
            s = ' '.join(f'{NAMES.get(t, t)}@{(c - t0)}' for t, c in ev[:70])
            print(f'  cta {cta} {rn}: {s}')
    # summary over CTAs: total span, time in mma waits
    spans = []
    for (cta, role), ev in rows.items():
        if role == 2 and ev:
            spans.append(ev[-1][1] - rows[(cta, 0)][0][1])
    print(f'  CTA lifetime clocks: mean {np.mean(spans):.0f} min {np.min(spans)} max {np.max(spans)} ({np.mean(spans) / 1.965e3:.1f} us at 1.965 GHz)')
