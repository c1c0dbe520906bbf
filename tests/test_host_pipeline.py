"""Host-side logic end to end on CPU: the product's EncodeStream/ConvertStream/DecodeStream +
VoiceChanger + RealtimeVocoder, with the GPU engine replaced by an oracle-backed stand-in, must
reproduce the oracle's independently written chunked stream (closed-form window indexing) exactly."""
import numpy as np
import pytest

from oracle import nets as onets
from oracle import pipeline as opipe
from realtime_yukarin_b200 import engine as eng_mod
from realtime_yukarin_b200 import synthetic
from realtime_yukarin_b200.config import VocodeMode
from realtime_yukarin_b200.models import AcousticConverter, F0Converter, SuperResolution
from realtime_yukarin_b200.params import create_from_json, create_sr_from_json
from realtime_yukarin_b200.stream import ConvertStream, DecodeStream, EncodeStream, StreamWrapper
from realtime_yukarin_b200.vocoder import RealtimeVocoder
from realtime_yukarin_b200.voice_changer import VoiceChanger
from tests.fake_engine import OracleEngine


@pytest.mark.parametrize('T,extra', [(0.3, (0.0, 0.5, 0.0)), (0.1, (0.1, 0.2, 0.0)), (0.3, (0.1, 0.5, 0.1)), (0.2, (0.0, 0.0, 0.0)), (0.1, (0.0, 0.5, 0.0))])
def test_stream_classes_reproduce_oracle_stream(small_models, T, extra):
    paths = small_models
    fake = OracleEngine(paths['stage1_model_path'], paths['stage2_model_path'])
    eng_mod.set_default_engine(fake)
    try:
        f0c = F0Converter(paths['input_statistics_path'], paths['target_statistics_path'])
        ac = AcousticConverter(create_from_json(paths['stage1_config_path']), paths['stage1_model_path'], f0_converter=f0c, engine=fake)
        sr = SuperResolution(create_sr_from_json(paths['stage2_config_path']), paths['stage2_model_path'], engine=fake)
        acp = create_from_json(paths['stage1_config_path']).dataset.acoustic_param
        voc = RealtimeVocoder(acoustic_param=acp, out_sampling_rate=24000, extract_f0_mode=VocodeMode.WORLD)
        voc.create_synthesizer(buffer_size=1024, number_of_pointers=16)
        es, cs, ds = EncodeStream(voc), ConvertStream(VoiceChanger(ac, sr, threshold=60)), DecodeStream(voc)
        ws = [StreamWrapper(es, extra[0]), StreamWrapper(cs, extra[1]), StreamWrapper(ds, extra[2])]
        p1, p2 = onets.load_npz(paths['stage1_model_path']), onets.load_npz(paths['stage2_model_path'])
        orc = opipe.StreamOracle(opipe.PathConfig(), p1, p2, f0c.stats(), buffer_time=T, extra=extra, backend='torch')
        x = synthetic.synthetic_speech(1.8, 21)
        n = round(T * 24000)
        for k in range(len(x) // n):
            chunk = x[k * n:(k + 1) * n]
            es.add(start_time=extra[0] + k * T, data=chunk)
            f = ws[0].process_next(T)
            cs.add(start_time=extra[1] + k * T, data=f)
            c = ws[1].process_next(T)
            ds.add(start_time=extra[2] + k * T, data=c)
            y = ws[2].process_next(T)
            r = orc.push(chunk)
            assert np.array_equal(f.f0, orc.last['encoded']['f0']), k
            assert np.array_equal(c.f0, orc.last['converted']['f0']), k
            assert np.allclose(c.sp, orc.last['converted']['sp'], rtol=1e-5), k
            assert len(y) == len(r), (k, len(y), len(r))
            assert np.allclose(y, r, atol=1e-9), k
    finally:
        eng_mod.set_default_engine(None)


def test_stage1_checkpoint_without_statistics_is_rejected(tmp_path):
    """a bare Chainer predictor.npz (encoder/decoder links only) must not silently convert un-normalised features."""
    import numpy as np
    import pytest
    from realtime_yukarin_b200 import synthetic
    from realtime_yukarin_b200.models import STATS_KEYS, load_stage1_stats
    p = synthetic.make_stage1_params(0, base=8)
    st = load_stage1_stats(p, tmp_path / 'predictor.npz', 9, 9)
    assert np.array_equal(st[0], synthetic.MC_MEAN_IN) and np.array_equal(st[3], synthetic.MC_STD_OUT)
    bare = {k: v for k, v in p.items() if k not in STATS_KEYS}
    with pytest.raises(ValueError, match='normalisation statistics'):
        load_stage1_stats(bare, tmp_path / 'predictor.npz', 9, 9)
    ident = load_stage1_stats(bare, tmp_path / 'predictor.npz', 9, 9, feature_stats='identity')
    assert not ident[0].any() and (ident[1] == 1).all()
    np.savez(tmp_path / 'stats.npz', **{k: p[k] for k in STATS_KEYS})              # side file next to the model
    side = load_stage1_stats(bare, tmp_path / 'predictor.npz', 9, 9)
    assert np.array_equal(side[2], synthetic.MC_MEAN_OUT)
    with pytest.raises(ValueError, match='lengths'):
        load_stage1_stats(bare, tmp_path / 'predictor.npz', 8, 9)
