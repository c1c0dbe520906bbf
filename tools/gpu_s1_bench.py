"""Stand-alone timing of the stage-1 forward: one cluster kernel (csrc/s1_fused.cu) vs the 16-layer launch sequence, per padded length,
plus the fused kernel's phase timeline (CTA 0, device globaltimer).  Run under gpurun; prints to stdout."""
import sys
import tempfile
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from realtime_yukarin_b200 import synthetic
from realtime_yukarin_b200.engine import default_engine
from realtime_yukarin_b200.models import AcousticConverter, F0Converter
from realtime_yukarin_b200.params import create_from_json

eng = default_engine()
d = tempfile.mkdtemp(prefix='ryk_s1bench_')
paths = synthetic.write_synthetic_models(d, seed=0)
f0c = F0Converter(paths['input_statistics_path'], paths['target_statistics_path'])
ac = AcousticConverter(create_from_json(paths['stage1_config_path']), paths['stage1_model_path'], f0_converter=f0c, engine=eng)
eng.set_precision('fp16')
print('cluster size', eng.set_stage1_fused(True))
only = [int(a) for a in sys.argv[1:]] or [128, 256, 384, 512, 640]
for Tp in only:
    fused, layered, tl = eng.stage1_bench(Tp, 50)
    print(f'Tp {Tp}: fused {fused * 1e3:.1f} us / forward, layered {layered * 1e3:.1f} us / forward (stand-alone, back-to-back)')
    names = ['layer0'] + [f'L{l + 1}' for l in range(14)]
    line = [f'start->layer0+barrier {tl[1]:.1f}']
    for l in range(14):
        line.append(f'L{l + 1}: tasks {tl[2 + 2 * l] - tl[1 + 2 * l]:.1f} barrier {tl[3 + 2 * l] - tl[2 + 2 * l]:.1f}')
    line.append(f'layer15 {tl[30] - tl[29]:.1f}; total {tl[30]:.1f} us')
    print('   ' + ' | '.join(line))
