"""GPU parity for the SURVEY 8(f) rows built on top of the hot path: offline Synthesis() (Vocoder.decode =
pyworld.synthesize), the output silence gate + re-blocking of the decode worker, and the worker pipeline
(RealtimePipeline = encode | convert | decode workers + audio-loop re-ordering), all through the C ABI.

Tolerances: pulse positions / voiced flags / gate decisions exact; fractional pulse shifts 1e-12 s; waveforms sample RMSE
<= 1e-9 * max(1, peak) (FP64 on both sides, same operation order); mean STFT power 1e-9 dB."""
import numpy as np
import pytest

from oracle import nets as onets
from oracle import pipeline as opipe
from oracle import world as oworld
from realtime_yukarin_b200 import synthetic

pytestmark = pytest.mark.gpu
CFG = opipe.PathConfig()


@pytest.mark.parametrize('seconds,stream', [(1.0, 2), (0.3, 5), (3.0, 11)])
def test_offline_synthesis_matches_oracle(engine, seconds, stream):
    x = synthetic.synthetic_speech(seconds, stream=stream)
    f = opipe.extract_features(x, CFG)
    f0 = f['f0'].ravel().astype(np.float64)
    yr, ir, sr_, vr = oworld.synthesize(f0, f['sp'], f['ap'], CFG.fs, CFG.frame_period, return_pulses=True)
    yg, ig, sg, vg = engine.world_synthesize(f0, f['sp'], f['ap'], CFG.fs, CFG.frame_period, return_pulses=True)
    assert len(yg) == len(yr) == int(len(f0) * CFG.frame_period * CFG.fs / 1000)
    assert np.array_equal(ig, ir), 'pulse positions differ'
    assert np.array_equal(vg, vr), 'voiced flags at the pulses differ'
    assert np.allclose(sg, sr_, rtol=0, atol=1e-12)
    rmse = float(np.sqrt(np.mean((yg - yr) ** 2)))
    print(f'offline synthesis {seconds} s: {len(ir)} pulses, rmse {rmse:.3e}, rms {float(np.sqrt(np.mean(yr ** 2))):.3e}')
    assert rmse <= 1e-9 * max(1.0, float(np.abs(yr).max()))


def test_offline_synthesis_edge_cases(engine):
    nb = 513
    # all-unvoiced, constant envelope; one frame; f0 below the synthesis floor fs / fft + 1 (treated as unvoiced)
    for f0 in (np.zeros(40), np.zeros(1), np.full(30, 20.0), np.concatenate([np.zeros(10), np.full(25, 220.0), np.zeros(7)])):
        sp = np.full((len(f0), nb), 1e-4, np.float32)
        ap = np.full((len(f0), nb), 0.3, np.float32)
        yr = oworld.synthesize(f0, sp, ap, 24000, 5.0)
        yg = engine.world_synthesize(f0, sp, ap, 24000, 5.0)
        assert len(yg) == len(yr)
        if len(yr):
            assert np.allclose(yg, yr, rtol=0, atol=1e-12 * max(1.0, float(np.abs(yr).max()))), len(f0)


def test_vocoder_decode_offline(engine):
    from realtime_yukarin_b200 import engine as eng_mod
    from realtime_yukarin_b200.config import VocodeMode
    from realtime_yukarin_b200.feature import AcousticFeature, Wave
    from realtime_yukarin_b200.params import AcousticParam
    from realtime_yukarin_b200.vocoder import Vocoder
    eng_mod.set_default_engine(engine)
    voc = Vocoder(acoustic_param=AcousticParam(), out_sampling_rate=24000, extract_f0_mode=VocodeMode.WORLD)
    x = synthetic.synthetic_speech(0.8, stream=4)
    feat = voc.encode(Wave(wave=x, sampling_rate=24000))
    w = voc.decode(feat)
    ref = oworld.synthesize(np.asarray(feat.f0, np.float64).ravel(), feat.sp, feat.ap, 24000, 5.0)
    assert len(w.wave) == len(ref)
    assert float(np.sqrt(np.mean((w.wave - ref) ** 2))) < 1e-9


@pytest.mark.parametrize('n,scale', [(7200, 0.1), (2400, 1e-3), (24000, 1e-6), (7200, 0.0), (4800, 3e-5)])
def test_output_gate_matches_oracle(engine, n, scale):
    rng = np.random.default_rng(n + 1)
    x = rng.standard_normal(n) * scale
    for thr in (60.0, 80.0):
        pw, keep = engine.output_gate(x, thr)
        ref = oworld.stft_power_db_mean(x)
        assert abs(pw - ref) < 1e-9, (pw, ref)
        assert keep == (not (ref < -thr))


def test_reblocker_matches_decode_worker_logic(engine):
    """Host-buffer entry point: same fragment bookkeeping and gate decisions as decode_worker.py:38-59."""
    rng = np.random.default_rng(3)
    for chunk in (2400, 7200):
        rid = engine.reblock_create(chunk, 16384, 80.0)
        ref = opipe.OutputReblockOracle(chunk, 80.0)
        for k, blocks in enumerate((2, 3, 2, 0, 3, 9, 7, 7, 1, 8, 7, 7)):
            amp = 1e-9 if k in (5, 6) else 0.05
            w = rng.standard_normal(blocks * 1024) * amp
            st, got, pw = engine.reblock_push(rid, w)
            rst, rchunk = ref.push(w)
            assert st == rst, (chunk, k, st, rst)
            if rst == 1:
                assert np.array_equal(got, rchunk)
            if rst != 0:
                assert abs(pw - ref.last_power) < 1e-9
        engine.reblock_destroy(rid)


def _load(engine, paths):
    from realtime_yukarin_b200.models import AcousticConverter, F0Converter, SuperResolution
    from realtime_yukarin_b200.params import create_from_json, create_sr_from_json
    f0c = F0Converter(paths['input_statistics_path'], paths['target_statistics_path'])
    ac = AcousticConverter(create_from_json(paths['stage1_config_path']), paths['stage1_model_path'], f0_converter=f0c, engine=engine)
    sr = SuperResolution(create_sr_from_json(paths['stage2_config_path']), paths['stage2_model_path'], engine=engine)
    return ac, sr, f0c


def _config(paths, buffer_time, out_thr):
    from realtime_yukarin_b200.config import Config, VocodeMode
    return Config(input_device_name=None, output_device_name=None, input_rate=24000, output_rate=24000, frame_period=5.0,
                  buffer_time=buffer_time, extract_f0_mode=VocodeMode.WORLD, vocoder_buffer_size=1024, input_scale=0.5, output_scale=2.0,
                  input_silent_threshold=60.0, output_silent_threshold=out_thr, encode_extra_time=0.0, convert_extra_time=0.5,
                  decode_extra_time=0.0, **{k: paths[k] for k in ('input_statistics_path', 'target_statistics_path', 'stage1_model_path',
                                                                  'stage1_config_path', 'stage2_model_path', 'stage2_config_path')})


@pytest.mark.parametrize('T', [0.3, 0.1])
def test_realtime_pipeline_matches_oracle_workers(engine, small_models, T):
    """RealtimePipeline (device session + device re-blocker/gate, 3 chunks in flight) == the oracle's chunked stream followed by
    the decode worker's re-blocking and silence gate, item by item and in index order."""
    from realtime_yukarin_b200.worker import Item, RealtimePipeline
    ac, sr, f0c = _load(engine, small_models)
    engine.set_precision('fp32')
    p1, p2 = onets.load_npz(small_models['stage1_model_path']), onets.load_npz(small_models['stage2_model_path'])
    x = synthetic.synthetic_speech(3.0, stream=41)
    x[int(1.2 * 24000):int(2.1 * 24000)] *= 1e-6                   # a muted stretch (the input gate turns it into silent frames)
    n = round(T * 24000)
    K = len(x) // n
    # the oracle's raw decode-worker input, then a threshold that gates roughly half of the chunks (stage 2 with random weights
    # does not keep muted frames quiet, so a fixed dB value would gate nothing or everything)
    orc = opipe.StreamOracle(CFG, p1, p2, f0c.stats(), buffer_time=T, extra=(0.0, 0.5, 0.0), backend='torch')
    raw = [orc.push(x[k * n:(k + 1) * n]) for k in range(K)]
    probe = opipe.OutputReblockOracle(n, 1e9)
    powers = []
    for r in raw:
        if probe.push(r)[0] != 0:
            powers.append(probe.last_power)
    ps = np.sort(np.asarray(powers))
    gi = int(np.argmax(np.diff(ps)))                               # widest gap between consecutive chunk powers: robust split point
    thr = -float(0.5 * (ps[gi] + ps[gi + 1]))
    cfg = _config(small_models, T, thr)
    pipe = RealtimePipeline(cfg, acoustic_param=ac.config.dataset.acoustic_param, engine=engine, depth=3)
    rb = opipe.OutputReblockOracle(cfg.out_audio_chunk, cfg.output_silent_threshold)
    expected = [rb.push(r) for r in raw]
    margin = min(abs(pw + thr) for pw in powers)                   # distance of the closest chunk to the threshold (dB)
    assert margin > 1e-3, 'degenerate threshold'
    got = []
    for k in range(K):
        pipe.put(Item(item=x[k * n:(k + 1) * n], index=k))
        while True:
            it = pipe.get_nowait()
            if it is None:
                break
            got.append(it)
    pipe.flush()
    while True:
        it = pipe.get_nowait()
        if it is None:
            break
        got.append(it)
    assert [it.index for it in got] == list(range(K))
    n_chunks = n_silent = 0
    for it, (st, ref) in zip(got, expected):
        assert (it.item is None) == (ref is None), (it.index, st)
        if ref is not None:
            n_chunks += 1
            rmse = float(np.sqrt(np.mean((it.item - ref) ** 2)))
            assert rmse < 1e-3, (it.index, rmse)
        n_silent += st == 2
    print(f'pipeline T={T}: {K} items, {n_chunks} chunks played, {n_silent} gated as silent')
    assert n_chunks > 0 and n_silent > 0
    pipe.close()
    engine.set_precision('fp16')


def test_audio_loop_body(engine, small_models):
    """process() = one iteration of run.py:160-199: scaled input in, exactly out_audio_chunk float32 samples out, zeros while
    nothing is ready or the chunk was silent."""
    from realtime_yukarin_b200.worker import RealtimePipeline
    ac, sr, f0c = _load(engine, small_models)
    cfg = _config(small_models, 0.3, 80.0)
    pipe = RealtimePipeline(cfg, acoustic_param=ac.config.dataset.acoustic_param, engine=engine, depth=2)
    x = synthetic.synthetic_speech(3.0, stream=8)
    n = cfg.in_audio_chunk
    outs = [pipe.process(x[k * n:(k + 1) * n]) for k in range(len(x) // n)]
    assert all(o.dtype == np.float32 and len(o) == cfg.out_audio_chunk for o in outs)
    assert not outs[0].any()                                      # nothing can be ready after the first put
    assert any(o.any() for o in outs[3:])
    pipe.close()
