import sys; sys.path.insert(0, '.')
import tempfile
import numpy as np
from oracle import nets as onets, pipeline as opipe
from realtime_yukarin_b200 import synthetic
from realtime_yukarin_b200.config import VocodeMode
from realtime_yukarin_b200.engine import default_engine
from realtime_yukarin_b200.models import AcousticConverter, F0Converter, SuperResolution
from realtime_yukarin_b200.params import create_from_json, create_sr_from_json
from realtime_yukarin_b200.stream import ConvertStream, DecodeStream, EncodeStream, StreamWrapper
from realtime_yukarin_b200.vocoder import RealtimeVocoder
from realtime_yukarin_b200.voice_changer import VoiceChanger
np.set_printoptions(linewidth=200, precision=5, suppress=True)
eng = default_engine()
paths = synthetic.write_synthetic_models(tempfile.mkdtemp(), seed=3, base1=16, base2=16)
f0c = F0Converter(paths['input_statistics_path'], paths['target_statistics_path'])
ac = AcousticConverter(create_from_json(paths['stage1_config_path']), paths['stage1_model_path'], f0_converter=f0c, engine=eng)
sr = SuperResolution(create_sr_from_json(paths['stage2_config_path']), paths['stage2_model_path'], engine=eng)
p1, p2 = onets.load_npz(paths['stage1_model_path']), onets.load_npz(paths['stage2_model_path'])
acp = create_from_json(paths['stage1_config_path']).dataset.acoustic_param
eng.set_precision('fp32')
T, extra = 0.3, (0.0, 0.5, 0.0)
voc = RealtimeVocoder(acoustic_param=acp, out_sampling_rate=24000, extract_f0_mode=VocodeMode.WORLD)
voc.create_synthesizer(buffer_size=1024, number_of_pointers=16)
es, cs, ds = EncodeStream(voc), ConvertStream(VoiceChanger(ac, sr, threshold=60)), DecodeStream(voc)
ws = [StreamWrapper(es, extra[0]), StreamWrapper(cs, extra[1]), StreamWrapper(ds, extra[2])]
orc = opipe.StreamOracle(opipe.PathConfig(), p1, p2, f0c.stats(), buffer_time=T, extra=extra, backend='torch')
x = synthetic.synthetic_speech(2.4, int(sys.argv[1]) if len(sys.argv) > 1 else 21)
n = round(T * 24000)
for k in range(len(x) // n):
    chunk = x[k * n:(k + 1) * n]
    es.add(start_time=extra[0] + k * T, data=chunk); f = ws[0].process_next(T)
    cs.add(start_time=extra[1] + k * T, data=f); c = ws[1].process_next(T)
    ds.add(start_time=extra[2] + k * T, data=c); y = ws[2].process_next(T)
    r = orc.push(chunk)
    e, cv = orc.last['encoded'], orc.last['converted']
    print(f'chunk {k}: enc f0 maxdiff {np.abs(f.f0 - e["f0"]).max():.3e} voiced eq {np.array_equal(f.voiced, e["voiced"])} '
          f'mc {np.abs(f.mc - e["mc"]).max():.2e} ap {np.abs(f.ap - e["ap"]).max():.2e} | conv f0 {np.abs(c.f0 - cv["f0"]).max():.3e} '
          f'logsp {np.abs(np.log(c.sp) - np.log(cv["sp"])).max():.2e} ap {np.abs(c.ap - cv["ap"]).max():.2e} | out len {len(y)} vs {len(r)} '
          f'rmse {np.sqrt(np.mean((y[:min(len(y), len(r))] - r[:min(len(y), len(r))]) ** 2)):.3e}')
    bad = np.where(np.abs(f.ap - e['ap']).max(axis=1) > 1e-3)[0]
    if len(bad): print('   ap-bad frames', bad, 'f0 there', e['f0'].ravel()[bad])
    bad = np.where(np.abs(c.f0 - cv['f0']).ravel() > 1e-3)[0]
    if len(bad): print('   conv-f0-bad frames', bad, c.f0.ravel()[bad], cv['f0'].ravel()[bad])
