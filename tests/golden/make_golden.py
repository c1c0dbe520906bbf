"""Regenerates tests/golden/golden_small.npz from the CPU oracle (run from the repo root).
The reference ships no golden vectors for this path and cannot run here, so these
fixtures pin the oracle against itself over time (regression), not against the reference."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import pipeline as opipe  # noqa: E402
from oracle import world as W  # noqa: E402
from realtime_yukarin_b200 import synthetic  # noqa: E402

meta = dict(seconds=0.9, stream=13)
x = synthetic.synthetic_speech(meta['seconds'], stream=meta['stream'])
cfg = opipe.PathConfig()
f = opipe.extract_features(x, cfg)
s = W.RealtimeSynthesizer(24000, 5.0, 1024, 1024)
y = s.decode(f['f0'].ravel().astype(np.float64), f['sp'], f['ap'])
out = Path(__file__).resolve().parent
np.savez_compressed(out / 'golden_small.npz', wave=x, f0=f['f0'].ravel(), voiced=f['voiced'].ravel(),
                    log_sp_sub=np.log(f['sp'][:, ::16]).astype(np.float32), ap_sub=f['ap'][:, ::16], mc=f['mc'],
                    resynth=y.astype(np.float32))
(out / 'golden_meta.json').write_text(json.dumps(meta))
print('wrote', out / 'golden_small.npz', 'frames', len(f['f0']), 'samples out', len(y))
