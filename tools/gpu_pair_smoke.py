"""First contact with the CTA-pair kernel: one small conv through ryk_test_conv_layer with RYK_TC2=2, in a child process with a
timeout so that a hang cannot take the box down."""
import os, sys; sys.path.insert(0, '.')
import numpy as np
os.environ['RYK_TC2'] = sys.argv[1] if len(sys.argv) > 1 else '2'
from realtime_yukarin_b200.engine import default_engine
import torch, torch.nn.functional as F
eng = default_engine()
rng = np.random.default_rng(0)
for tr, B, H, W, C0, C1, Cout in ((0, 1, 32, 64, 64, 0, 128), (1, 1, 12, 16, 128, 128, 64), (1, 2, 24, 32, 64, 64, 128), (1, 1, 3, 4, 256, 0, 256)):
    in0 = rng.standard_normal((B, H, W, C0)).astype(np.float32)
    in1 = rng.standard_normal((B, H, W, C1)).astype(np.float32) if C1 else None
    Cin = C0 + C1
    Wt = (rng.standard_normal((Cin, Cout, 4, 4) if tr else (Cout, Cin, 4, 4)) / np.sqrt(Cin * 16 / (4 if tr else 1))).astype(np.float32)
    x = in0 if in1 is None else np.concatenate([in0, in1], 3)
    xt = torch.from_numpy(x).permute(0, 3, 1, 2).double()
    y = F.conv_transpose2d(xt, torch.from_numpy(Wt).double(), stride=2, padding=1) if tr else F.conv2d(xt, torch.from_numpy(Wt).double(), stride=2, padding=1)
    ref = y.permute(0, 2, 3, 1).float().numpy()
    got, ms = eng.test_conv_layer(in0, in1, Wt, np.ones(Cout, np.float32), np.zeros(Cout, np.float32), tr, 4, 2, 1, 0, use_tc=1, repeat=2)
    print('case', (tr, B, H, W, C0, C1, Cout), 'max err', float(np.abs(got - ref).max()), 'ref max', float(np.abs(ref).max()), 'ms', ms, flush=True)
