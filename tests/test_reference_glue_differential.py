"""The reference's REAL glue classes on top of this package's third-party replacements (CPU, reference checkout required).

`realtime_voice_conversion/{stream,segment}/*.py`, `yukarin_wrapper/voice_changer.py` and
`yukarin_wrapper/acoustic_feature_wrapper.py` are pure Python over `yukarin` / `become_yukarin` / the vocoder; here they are
imported from the read-only checkout with `yukarin`, `become_yukarin` resolved by `dropin` and `yukarin_wrapper.vocoder` (the
module that binds pyworld / world4py) replaced by ours.  Driving the reference's EncodeStream -> ConvertStream(VoiceChanger) ->
DecodeStream chain the way its workers do must give exactly what this package's own classes give on the same engine."""
import importlib
import sys
from pathlib import Path

import numpy as np
import pytest

REF_ROOT = Path('/root/reference')
pytestmark = pytest.mark.skipif(not (REF_ROOT / 'realtime_voice_conversion').exists(), reason='reference checkout not present (GPU box)')


class _RealReferencePackage:
    """Context manager: `realtime_voice_conversion` resolves to the real checkout (except yukarin_wrapper.vocoder = ours)."""

    def __enter__(self):
        from realtime_yukarin_b200 import dropin, vocoder
        import types
        dropin.install()                                   # yukarin / become_yukarin / librosa aliases
        self.saved = {k: v for k, v in sys.modules.items() if k == 'realtime_voice_conversion' or k.startswith('realtime_voice_conversion.')}
        for k in self.saved:
            del sys.modules[k]
        pkg = types.ModuleType('realtime_voice_conversion')
        pkg.__path__ = [str(REF_ROOT / 'realtime_voice_conversion')]      # real files for every submodule ...
        sys.modules['realtime_voice_conversion'] = pkg
        yw = types.ModuleType('realtime_voice_conversion.yukarin_wrapper')
        yw.__path__ = [str(REF_ROOT / 'realtime_voice_conversion' / 'yukarin_wrapper')]
        sys.modules['realtime_voice_conversion.yukarin_wrapper'] = yw
        voc = types.ModuleType('realtime_voice_conversion.yukarin_wrapper.vocoder')    # ... except the pyworld / world4py binding
        voc.Vocoder, voc.RealtimeVocoder = vocoder.Vocoder, vocoder.RealtimeVocoder
        sys.modules['realtime_voice_conversion.yukarin_wrapper.vocoder'] = voc
        return self

    def load(self, name):
        return importlib.import_module(f'realtime_voice_conversion.{name}')

    def __exit__(self, *exc):
        for k in [k for k in sys.modules if k == 'realtime_voice_conversion' or k.startswith('realtime_voice_conversion.')]:
            del sys.modules[k]
        sys.modules.update(self.saved)
        return False


@pytest.mark.parametrize('T,extra', [(0.3, (0.0, 0.5, 0.0)), (0.1, (0.1, 0.2, 0.1))])
def test_reference_streams_and_voice_changer_over_our_replacements(small_models, T, extra):
    from realtime_yukarin_b200 import engine as eng_mod
    from realtime_yukarin_b200 import stream as our_stream
    from realtime_yukarin_b200 import synthetic
    from realtime_yukarin_b200 import voice_changer as our_vc
    from realtime_yukarin_b200.config import VocodeMode
    from realtime_yukarin_b200.models import AcousticConverter, F0Converter, SuperResolution
    from realtime_yukarin_b200.params import create_from_json, create_sr_from_json
    from realtime_yukarin_b200.vocoder import RealtimeVocoder
    from tests.fake_engine import OracleEngine
    fake = OracleEngine(small_models['stage1_model_path'], small_models['stage2_model_path'])
    eng_mod.set_default_engine(fake)
    try:
        f0c = F0Converter(small_models['input_statistics_path'], small_models['target_statistics_path'])
        ac = AcousticConverter(create_from_json(small_models['stage1_config_path']), small_models['stage1_model_path'], f0_converter=f0c, engine=fake)
        sr = SuperResolution(create_sr_from_json(small_models['stage2_config_path']), small_models['stage2_model_path'], engine=fake)
        acp = create_from_json(small_models['stage1_config_path']).dataset.acoustic_param

        def chain(EncodeStream, ConvertStream, DecodeStream, StreamWrapper, VoiceChanger):
            voc = RealtimeVocoder(acoustic_param=acp, out_sampling_rate=24000, extract_f0_mode=VocodeMode.WORLD)
            voc.create_synthesizer(buffer_size=1024, number_of_pointers=16)
            es, cs, ds = EncodeStream(vocoder=voc), ConvertStream(voice_changer=VoiceChanger(acoustic_converter=ac, super_resolution=sr, threshold=60)), DecodeStream(vocoder=voc)
            ws = [StreamWrapper(stream=es, extra_time=extra[0]), StreamWrapper(stream=cs, extra_time=extra[1]), StreamWrapper(stream=ds, extra_time=extra[2])]
            x = synthetic.synthetic_speech(1.5, 31)
            n = round(T * 24000)
            outs = []
            for k in range(len(x) // n):
                es.add(start_time=extra[0] + k * T, data=x[k * n:(k + 1) * n])
                f = ws[0].process_next(time_length=T)
                cs.add(start_time=extra[1] + k * T, data=f)
                c = ws[1].process_next(time_length=T)
                ds.add(start_time=extra[2] + k * T, data=c)
                y = ws[2].process_next(time_length=T)
                outs.append((np.asarray(f.f0).copy(), np.asarray(c.f0).copy(), np.asarray(c.sp).copy(), np.asarray(y.wave if hasattr(y, 'wave') else y).copy()))
            return outs

        with _RealReferencePackage() as ref:
            rs = ref.load('stream')
            rvc = ref.load('yukarin_wrapper.voice_changer')
            assert Path(rs.__file__).is_relative_to(REF_ROOT) and Path(rvc.__file__).is_relative_to(REF_ROOT)      # really the checkout's code
            got_ref = chain(rs.EncodeStream, rs.ConvertStream, rs.DecodeStream, rs.StreamWrapper, rvc.VoiceChanger)
        got_ours = chain(our_stream.EncodeStream, our_stream.ConvertStream, our_stream.DecodeStream, our_stream.StreamWrapper, our_vc.VoiceChanger)
        assert len(got_ref) == len(got_ours) > 0
        for k, (a, b) in enumerate(zip(got_ref, got_ours)):
            for u, v in zip(a, b):
                assert u.shape == v.shape and np.array_equal(u, v, equal_nan=True), k
    finally:
        eng_mod.set_default_engine(None)
