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
# offline Synthesis() (pulse plan + waveform), output-gate power of 0.3 s chunks of that waveform, re-blocker statuses
yo, pidx, pshift, pvuv = W.synthesize(f['f0'].ravel().astype(np.float64), f['sp'], f['ap'], 24000, 5.0, return_pulses=True)
powers = np.array([W.stft_power_db_mean(yo[a:a + 7200]) for a in range(0, len(yo) - 7199, 7200)])
rb = opipe.OutputReblockOracle(7200, 30.0)
statuses = np.array([rb.push(y[a:a + 1024] if a // 1024 % 5 else 1e-6 * y[a:a + 1024])[0] for a in range(0, len(y), 1024)], dtype=np.int32)
np.savez_compressed(out / 'golden_widen.npz', offline=yo.astype(np.float32), pulse_index=pidx, pulse_shift=pshift, pulse_vuv=pvuv, gate_power=powers,
                    reblock_status=statuses)
(out / 'golden_meta.json').write_text(json.dumps(meta))
print('wrote', out / 'golden_small.npz', 'frames', len(f['f0']), 'samples out', len(y))
