import sys; sys.path.insert(0, '.')
import numpy as np
from realtime_yukarin_b200 import synthetic
from realtime_yukarin_b200.engine import default_engine
eng = default_engine()
x = synthetic.synthetic_speech(0.3, stream=0)
got = eng.world_analyze(x, 24000, 5.0, 71.0, 800.0, 1024, 8, 0.466)
print('ok', got['f0'][:5])
