"""SURVEY 8(f) ranks 1-3 on the CPU: the oracle's offline Synthesis() and output-gate restatements against independent
numpy / scipy computations and analytic properties, and the host-side worker logic (Item / OutputReblocker / the
run.py audio-loop body) with the GPU engine replaced by the oracle-backed stand-in."""
import numpy as np
import pytest
import scipy.signal as ss

from oracle import pipeline as opipe
from oracle import world as W
from realtime_yukarin_b200 import engine as eng_mod
from realtime_yukarin_b200 import synthetic
from tests.fake_engine import OracleEngine

CFG = opipe.PathConfig()


def _librosa_style_power_db_mean(x, n_fft=2048, hop=512):
    xp = np.pad(x, n_fft // 2, mode='reflect')
    win = ss.get_window('hann', n_fft, fftbins=True)
    frames = 1 + len(x) // hop
    S = np.stack([np.fft.rfft(xp[f * hop:f * hop + n_fft] * win) for f in range(frames)], 1)
    db = 10 * np.log10(np.maximum(1e-10, np.abs(S) ** 2))
    return float(np.maximum(db, db.max() - 80.0).mean())


@pytest.mark.parametrize('n,scale', [(7200, 0.1), (2400, 1e-3), (24000, 1e-7), (7200, 0.0)])
def test_output_gate_oracle_matches_numpy_stft(n, scale):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) * scale
    assert abs(W.stft_power_db_mean(x) - _librosa_style_power_db_mean(x)) < 1e-9
    if scale == 0.0:       # all-zero chunk: every bin clamps to amin -> -100 dB, below any sensible threshold
        assert W.stft_power_db_mean(x) == pytest.approx(-100.0)


def test_output_gate_scaling_property():
    """power_to_db is 20 log10 of the amplitude: scaling the chunk by 10 raises the mean by exactly 20 dB (no clipping active)."""
    rng = np.random.default_rng(5)
    x = rng.standard_normal(7200) * 0.01
    assert W.stft_power_db_mean(10 * x) - W.stft_power_db_mean(x) == pytest.approx(20.0, abs=1e-9)


def test_offline_synthesis_length_and_pulses():
    x = synthetic.synthetic_speech(1.0, stream=2)
    f = opipe.extract_features(x, CFG)
    f0 = f['f0'].ravel().astype(np.float64)
    y, idx, shift, vuv = W.synthesize(f0, f['sp'], f['ap'], 24000, 5.0, return_pulses=True)
    assert len(y) == int(len(f0) * 5.0 * 24000 / 1000)                 # pyworld.synthesize's y_length
    assert np.all(np.diff(idx) > 0) and idx[0] >= 0 and idx[-1] < len(y) - 1
    assert np.all((shift >= 0) & (shift <= 1.0 / 24000 + 1e-18))         # fractional shift lies inside one sample
    # voiced pulses follow the f0 contour: interval between consecutive voiced pulses ~ fs / f0 at that frame
    iv = np.diff(idx)
    fr = np.minimum(len(f0) - 1, (idx[:-1] / 120.0).astype(int))
    both_voiced = (vuv[:-1] == 1) & (vuv[1:] == 1) & (f0[fr] > 0) & (f0[np.minimum(len(f0) - 1, fr + 1)] > 0)
    rel = np.abs(iv[both_voiced] - 24000.0 / f0[fr[both_voiced]]) / (24000.0 / f0[fr[both_voiced]])
    assert np.median(rel) < 0.05
    # unvoiced stretches pulse at the 500 Hz default: exactly every 48 samples
    un = (vuv[:-1] == 0) & (vuv[1:] == 0)
    assert un.any() and np.all(np.abs(iv[un] - 48) <= 1)
    # analysis -> synthesis keeps the level and the coarse spectrum
    n = min(len(x), len(y))
    assert 0.7 < np.sqrt(np.mean(y[:n] ** 2)) / np.sqrt(np.mean(x[:n] ** 2)) < 1.4
    X, Y = np.abs(np.fft.rfft(x[:n]))[:3000], np.abs(np.fft.rfft(y[:n]))[:3000]
    assert np.corrcoef(X, Y)[0, 1] > 0.7


def test_offline_synthesis_unvoiced_only_is_noise_shaped_by_sp():
    """f0 = 0 everywhere: no periodic part; output power follows the spectral envelope level (linearity in sqrt(sp))."""
    nb = 513
    f0 = np.zeros(100)
    ap = np.full((100, nb), 0.5, np.float32)
    y1 = W.synthesize(f0, np.full((100, nb), 1e-4, np.float32), ap, 24000, 5.0)
    y2 = W.synthesize(f0, np.full((100, nb), 4e-4, np.float32), ap, 24000, 5.0)
    assert np.allclose(y2, 2.0 * y1, rtol=1e-9, atol=1e-15)            # same noise draws, amplitude = sqrt(sp)
    assert np.isfinite(y1).all() and np.abs(y1).max() > 0


def test_reblock_oracle_matches_reference_fragment_logic():
    """decode_worker.py:38-59 traced by hand: blocks of 1024 -> chunks of 2400 samples, at most one per step."""
    rb = opipe.OutputReblockOracle(2400, 80.0)
    rng = np.random.default_rng(0)
    stream = rng.standard_normal(20 * 1024) * 0.05
    got, fed = [], 0
    for blocks in (2, 3, 2, 0, 3, 2, 2, 3, 3):
        st, chunk = rb.push(stream[fed:fed + blocks * 1024])
        fed += blocks * 1024
        got.append((st, chunk))
    emitted = [c for s, c in got if s == 1]
    assert [s for s, _ in got] == [0, 1, 1, 0, 1, 1, 1, 1, 1]      # 2048 | 5120 -> 2720 | 4768 -> 2368 | 2368 | 5440 -> ... (hand trace)
    assert np.array_equal(np.concatenate(emitted), stream[:len(emitted) * 2400])
    assert rb.push(np.zeros(2400))[0] in (1, 2)
    rb2 = opipe.OutputReblockOracle(2400, 80.0)
    assert rb2.push(np.zeros(4096))[0] == 2                                  # digital silence is gated away


def test_worker_classes_with_oracle_engine(small_models):
    from realtime_yukarin_b200.worker import Item, OutputReblocker
    fake = OracleEngine(small_models['stage1_model_path'], small_models['stage2_model_path'])
    eng_mod.set_default_engine(fake)
    try:
        it = Item(item=np.zeros(3), index=7)
        assert it.index == 7 and len(it.item) == 3
        rb = OutputReblocker(out_audio_chunk=7200, output_silent_threshold=80.0)
        ref = opipe.OutputReblockOracle(7200, 80.0)
        rng = np.random.default_rng(1)
        for k in range(6):
            w = rng.standard_normal(7 * 1024) * (0.1 if k != 3 else 1e-9)
            a = rb.push(w)
            st, b = ref.push(w)
            assert (a is None) == (b is None)
            if a is not None:
                assert np.array_equal(a, b)
            assert rb.last_status == st
        rb.close()
    finally:
        eng_mod.set_default_engine(None)


def test_vocoder_decode_is_offline_synthesis(small_models):
    """Vocoder.decode (vocoder.py:50-62) goes through world_synthesize and returns pyworld.synthesize's length."""
    from realtime_yukarin_b200.config import VocodeMode
    from realtime_yukarin_b200.feature import AcousticFeature
    from realtime_yukarin_b200.params import create_from_json
    from realtime_yukarin_b200.vocoder import Vocoder
    fake = OracleEngine(small_models['stage1_model_path'], small_models['stage2_model_path'])
    eng_mod.set_default_engine(fake)
    try:
        acp = create_from_json(small_models['stage1_config_path']).dataset.acoustic_param
        voc = Vocoder(acoustic_param=acp, out_sampling_rate=24000, extract_f0_mode=VocodeMode.WORLD)
        x = synthetic.synthetic_speech(0.5, stream=9)
        f = opipe.extract_features(x, CFG)
        feat = AcousticFeature(f0=f['f0'], sp=f['sp'], ap=f['ap'], mc=f['mc'], voiced=f['voiced'])
        w = voc.decode(feat)
        assert w.sampling_rate == 24000 and len(w.wave) == int(len(f['f0']) * 5.0 * 24)
        assert np.array_equal(w.wave, W.synthesize(f['f0'].ravel().astype(np.float64), f['sp'], f['ap'], 24000, 5.0))
    finally:
        eng_mod.set_default_engine(None)


def test_realtime_pipeline_audio_loop_on_cpu(small_models):
    """worker.RealtimePipeline over the oracle-backed session stand-in: put / get_nowait keep index order and item
    semantics (None = no chunk yet or silent), and process() is the run.py:160-199 loop body (scaling, re-ordering, zeros)."""
    from realtime_yukarin_b200.config import Config, VocodeMode
    from realtime_yukarin_b200.worker import Item, RealtimePipeline
    fake = OracleEngine(small_models['stage1_model_path'], small_models['stage2_model_path'])
    cfg = Config(input_device_name=None, output_device_name=None, input_rate=24000, output_rate=24000, frame_period=5.0, buffer_time=0.3,
                 extract_f0_mode=VocodeMode.WORLD, vocoder_buffer_size=1024, input_scale=0.5, output_scale=2.0, input_silent_threshold=60.0,
                 output_silent_threshold=80.0, encode_extra_time=0.0, convert_extra_time=0.5, decode_extra_time=0.0,
                 **{k: small_models[k] for k in ('input_statistics_path', 'target_statistics_path', 'stage1_model_path', 'stage1_config_path',
                                                 'stage2_model_path', 'stage2_config_path')})
    x = synthetic.synthetic_speech(2.4, stream=17)
    n = cfg.in_audio_chunk
    K = len(x) // n
    # expected: the same oracle objects driven by hand
    from oracle import nets as onets
    p1, p2 = onets.load_npz(small_models['stage1_model_path']), onets.load_npz(small_models['stage2_model_path'])
    stats = (float(np.log(150.0)), 0.2, float(np.log(250.0)), 0.2)
    orc = opipe.StreamOracle(opipe.PathConfig(threshold_db=60.0), p1, p2, stats, buffer_time=0.3, extra=(0.0, 0.5, 0.0), backend='torch')
    rb = opipe.OutputReblockOracle(cfg.out_audio_chunk, 80.0)
    expected = [rb.push(orc.push((x[k * n:(k + 1) * n] * cfg.input_scale).astype(np.float32)))[1] for k in range(K)]

    def want(e):
        return np.zeros(cfg.out_audio_chunk, np.float32) if e is None else (e * cfg.output_scale)[:cfg.out_audio_chunk].astype(np.float32)

    # (a) a device that has finished by the time the loop asks (the stand-in computes inside submit; poll -> True): the chunk is played
    #     in the SAME iteration, like the reference's get_nowait() once its workers are done -- not `depth` iterations later
    pipe = RealtimePipeline(cfg, engine=fake, depth=2)
    pipe.put(Item(item=x[:n] * cfg.input_scale, index=0))
    it = pipe.get_nowait()
    assert it is not None and it.index == 0 and (it.item is None) == (expected[0] is None)
    assert pipe.get_nowait() is None                      # nothing else in flight
    pipe.close()

    pipe = RealtimePipeline(cfg, engine=fake, depth=2)
    outs = [pipe.process(x[k * n:(k + 1) * n]) for k in range(K)]
    assert all(o.dtype == np.float32 and len(o) == cfg.out_audio_chunk for o in outs)
    for k in range(K):
        assert np.array_equal(outs[k], want(expected[k])), k
    assert pipe.drain() == []                             # nothing left in the pipeline
    pipe.close()

    # (b) a device that is never done when polled: the loop plays zeros until `depth` forces a collect (iteration k plays item k - depth),
    #     and drain() returns the tail that a finite (wav file) run would otherwise lose
    class Lagging(OracleEngine):
        def session_poll(self, sid, ticket):
            return False

        def reblock_poll(self, rid, ticket):
            return False
    lag = Lagging(small_models['stage1_model_path'], small_models['stage2_model_path'])
    pipe = RealtimePipeline(cfg, engine=lag, depth=2)
    pipe.put(Item(item=x[:n] * cfg.input_scale, index=0))
    assert pipe.get_nowait() is None                      # still in flight
    it = pipe.get()                                       # blocking get collects it
    assert it.index == 0 and (it.item is None) == (expected[0] is None)
    pipe.close()
    pipe = RealtimePipeline(cfg, engine=lag, depth=2)
    outs = [pipe.process(x[k * n:(k + 1) * n]) for k in range(K)]
    assert not outs[0].any() and not outs[1].any()
    for k in range(2, K):
        assert np.array_equal(outs[k], want(expected[k - 2])), k
    tail = pipe.drain()
    assert len(tail) == sum(e is not None for e in expected[K - 2:])
    for w, e in zip(tail, [e for e in expected[K - 2:] if e is not None]):
        assert np.array_equal(w, want(e))
    pipe.close()


def _numpy_offline_synthesis(f0, sp, ap, fs, frame_period_ms, fft):
    """Second, independent writing of WORLD's offline Synthesis() in vectorised numpy (numpy.fft, numpy.interp, cumulative sums in the
    oracle's blocked order), used only to cross-check oracle/world_oracle.c: wo_synthesize."""
    fp = frame_period_ms / 1000.0
    n = len(f0)
    ny = int(n * frame_period_ms * fs / 1000.0)
    lowest = fs / fft + 1.0
    cf = np.where(f0 < lowest, 0.0, f0)
    cv = (cf != 0).astype(float)
    ct = np.arange(n + 1) * fp
    cf = np.append(cf, cf[-1] * 2 - cf[-2])
    cv = np.append(cv, cv[-1] * 2 - cv[-2])
    t = np.arange(ny) / fs
    vuv = (np.interp(t, ct, cv) > 0.5).astype(float)
    f0i = np.where(vuv == 0, 500.0, np.interp(t, ct, cf))
    inc = 2 * np.pi * f0i / fs
    tp = np.empty(ny)
    base = 0.0
    for b0 in range(0, ny, 256):                        # the oracle's fixed blocked summation order
        loc = np.cumsum(inc[b0:b0 + 256])
        tp[b0:b0 + 256] = base + loc
        base = base + loc[-1]
    wp = np.fmod(tp, 2 * np.pi)
    idx = np.nonzero(np.abs(np.diff(wp)) > np.pi)[0]
    y1, y2 = wp[idx] - 2 * np.pi, wp[idx + 1]
    shift = (-y1 / (y2 - y1)) / fs
    half = fft // 2
    i = np.arange(half)
    dcr = 0.5 - 0.5 * np.cos(2 * np.pi * (i + 1.0) / (1.0 + fft))
    dcr = np.concatenate([dcr, dcr[::-1]])
    dcr /= dcr.sum()
    k = np.arange(half + 1)

    def min_phase(log_half):                            # folded-cepstrum minimum phase, as WORLD's GetMinimumPhaseSpectrum
        full = np.concatenate([log_half, log_half[-2:0:-1]])
        cep = np.fft.fft(full).real                     # WORLD uses a forward FFT and divides by n at the end
        cep[1:half] *= 2.0
        cep[half + 1:] = 0.0
        spec = np.fft.fft(cep)[:half + 1] / fft
        return np.exp(spec.real) * np.exp(1j * spec.imag)

    y = np.zeros(ny)
    for p, q in enumerate(idx):
        nxt = idx[min(p + 1, len(idx) - 1)]
        noise_size = min(int(nxt - q), fft)
        tt = t[q]
        fl, ce = min(n - 1, int(np.floor(tt / fp))), min(n - 1, int(np.ceil(tt / fp)))
        w = tt / fp - fl
        s = np.abs(sp[fl].astype(float)) if fl == ce else (1 - w) * np.abs(sp[fl].astype(float)) + w * np.abs(sp[ce].astype(float))
        clip = lambda a: np.clip(a.astype(float), 0.001, 0.999999999999)
        a = clip(ap[fl]) ** 2 if fl == ce else ((1 - w) * clip(ap[fl]) + w * clip(ap[ce])) ** 2
        if vuv[q] == 0 or a[0] > 0.999:
            per = np.zeros(fft)
        else:
            m = min_phase(np.log(s * (1 - a) + 1e-12) / 2)
            coef = 2 * np.pi * shift[p] * fs / fft
            re2 = np.cos(coef * k)
            m = m * (re2 - 1j * np.sqrt(1 - re2 * re2))
            per = np.fft.fftshift(np.fft.irfft(m, fft) * fft)
            dc = per[half:].sum()
            per = np.concatenate([-dc * dcr[:half], per[half:] - dc * dcr[half:]])
        nz = np.zeros(fft)
        if noise_size > 0:
            r = W.randn_stream(int(q), noise_size)
            nz[:noise_size] = r - r.mean()
        m = min_phase(np.log(s * a) / 2 if vuv[q] != 0 else np.log(s) / 2)
        aper = np.fft.fftshift(np.fft.irfft(m * np.fft.rfft(nz), fft) * fft)
        resp = (per * np.sqrt(noise_size) + aper) / fft
        off = int(q) - half + 1
        lo, hi = max(0, -off), min(fft, ny - off)
        y[off + lo:off + hi] += resp[lo:hi]
    return y, idx, shift


def test_offline_synthesis_oracle_agrees_with_an_independent_numpy_writing():
    x = synthetic.synthetic_speech(0.5, stream=3)
    f = opipe.extract_features(x, CFG)
    f0 = f['f0'].ravel().astype(np.float64)
    y, idx, shift, _ = W.synthesize(f0, f['sp'], f['ap'], 24000, 5.0, return_pulses=True)
    y2, idx2, shift2 = _numpy_offline_synthesis(f0, f['sp'], f['ap'], 24000, 5.0, 1024)
    assert np.array_equal(idx, idx2)
    assert np.allclose(shift, shift2, rtol=0, atol=1e-15)
    assert len(y) == len(y2)
    assert np.abs(y - y2).max() < 1e-9 * max(1.0, np.abs(y).max())


def test_run_module_wav_mode_matches_the_audio_loop(tmp_path, small_models):
    """realtime_yukarin_b200.run (the reference's run.py:22-199 with wav files instead of PyAudio devices): config.yaml -> models ->
    RealtimePipeline -> audio loop; the written wav is the concatenation of what process() returns chunk by chunk."""
    import yaml
    from realtime_yukarin_b200 import run as run_mod
    from realtime_yukarin_b200 import wave_io
    from realtime_yukarin_b200.config import Config
    from realtime_yukarin_b200.converter import YukarinConverter
    from realtime_yukarin_b200.worker import RealtimePipeline
    fake = OracleEngine(small_models['stage1_model_path'], small_models['stage2_model_path'])
    eng_mod.set_default_engine(fake)
    try:
        cfg = dict(input_device_name=None, output_device_name=None, input_rate=24000, output_rate=24000, frame_period=5, buffer_time=0.3,
                   extract_f0_mode='world', vocoder_buffer_size=1024, input_scale=0.5, output_scale=2.0, input_silent_threshold=60,
                   output_silent_threshold=80, encode_extra_time=0.0, convert_extra_time=0.5, decode_extra_time=0.0,
                   **{k: str(small_models[k]) for k in ('input_statistics_path', 'target_statistics_path', 'stage1_model_path',
                                                        'stage1_config_path', 'stage2_model_path', 'stage2_config_path')})
        (tmp_path / 'config.yaml').write_text(yaml.safe_dump(cfg))
        x = synthetic.synthetic_speech(1.6, stream=29)
        wave_io.write_wav(tmp_path / 'in.wav', x, 24000)
        n = run_mod.run(tmp_path / 'config.yaml', wav_in=tmp_path / 'in.wav', wav_out=tmp_path / 'out.wav', engine=fake, depth=2)
        assert n == len(x) // 7200
        got, sr = wave_io.read_wav(tmp_path / 'out.wav')
        assert sr == 24000 and len(got) == n * 7200
        # the same loop by hand
        config = Config.from_yaml(tmp_path / 'config.yaml')
        conv = YukarinConverter.make_yukarin_converter(**{k: small_models[k] for k in (
            'input_statistics_path', 'target_statistics_path', 'stage1_model_path', 'stage1_config_path', 'stage2_model_path',
            'stage2_config_path')})
        pipe = RealtimePipeline(config, acoustic_param=conv.acoustic_converter.config.dataset.acoustic_param, engine=fake, depth=2)
        ref = np.concatenate([pipe.process(x[k * 7200:(k + 1) * 7200]) for k in range(n)])
        pipe.close()
        assert np.array_equal(got, ref) and np.abs(got).max() > 0
    finally:
        eng_mod.set_default_engine(None)


def test_pipeline_logs_per_item_stage_times_at_debug(small_models, caplog):
    """SURVEY section 5 "Metrics/logging": the reference's workers log `<index>: <seconds>` per item at DEBUG on the loggers
    'encode' / 'convert' / 'decode' (worker/encode_worker.py:44, convert_worker.py:59, decode_worker.py:66); init_logger sets the
    level from $LOG_LEVEL (worker/utility.py:16-29)."""
    import logging
    from realtime_yukarin_b200.config import Config, VocodeMode
    from realtime_yukarin_b200.worker import RealtimePipeline, init_logger
    fake = OracleEngine(small_models['stage1_model_path'], small_models['stage2_model_path'])
    cfg = Config(input_device_name=None, output_device_name=None, input_rate=24000, output_rate=24000, frame_period=5.0, buffer_time=0.3,
                 extract_f0_mode=VocodeMode.WORLD, vocoder_buffer_size=1024, input_scale=1.0, output_scale=1.0, input_silent_threshold=60.0,
                 output_silent_threshold=80.0, encode_extra_time=0.0, convert_extra_time=0.5, decode_extra_time=0.0,
                 **{k: small_models[k] for k in ('input_statistics_path', 'target_statistics_path', 'stage1_model_path', 'stage1_config_path',
                                                 'stage2_model_path', 'stage2_config_path')})
    lg = logging.getLogger('rykprobe')
    n_handlers = len(lg.handlers)
    init_logger(lg)                                         # console handler, level from $LOG_LEVEL (WARNING by default), no file
    assert len(lg.handlers) == n_handlers + 1 and lg.level == logging.WARNING
    x = synthetic.synthetic_speech(0.7, stream=3)
    n = cfg.in_audio_chunk
    with caplog.at_level(logging.DEBUG):
        for name in ('encode', 'convert', 'decode'):
            logging.getLogger(name).setLevel(logging.DEBUG)
        try:
            pipe = RealtimePipeline(cfg, engine=fake, depth=2)
            pipe.process(x[:n]); pipe.process(x[n:2 * n])
            pipe.close()
        finally:
            for name in ('encode', 'convert', 'decode'):
                logging.getLogger(name).setLevel(logging.NOTSET)
    for name in ('encode', 'convert', 'decode'):
        msgs = [r.getMessage() for r in caplog.records if r.name == name and r.levelno == logging.DEBUG]
        assert len(msgs) == 2 and msgs[0].startswith('0: ') and msgs[1].startswith('1: '), (name, msgs)
