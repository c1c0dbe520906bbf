"""Regenerates tests/golden/harvest_f0.npz (regression fixture of the oracle's Harvest restatement; self-generated: the reference
ships no golden vectors and pyworld is not installable here)."""
from pathlib import Path

import numpy as np

from oracle import world as ow
from realtime_yukarin_b200 import synthetic

seconds, stream = 1.5, 2
x = synthetic.synthetic_speech(seconds, stream=stream).astype(np.float64)
f0, t = ow.harvest(x, 24000)
np.savez_compressed(Path(__file__).resolve().parent / 'harvest_f0.npz', f0=f0, seconds=seconds, stream=stream)
print('voiced frames', int((f0 > 0).sum()), 'of', len(f0))
