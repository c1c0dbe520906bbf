"""wav I/O and resampling row: filter design and geometry pinned against scipy.signal (CPU), wav
round trips, and -- on the GPU -- ryk_resample_poly against scipy.signal.resample_poly itself."""
import struct

import numpy as np
import pytest
import scipy.signal as ss

from realtime_yukarin_b200 import wave_io


@pytest.mark.parametrize('up,down', [(80, 147), (2, 3), (3, 2), (160, 441), (1, 2)])
def test_filter_design_matches_scipy_firwin(up, down):
    h = wave_io.resample_filter(up, down)
    ref = ss.firwin(2 * 10 * max(up, down) + 1, 1.0 / max(up, down), window=('kaiser', 5.0)) * up
    assert np.abs(h - ref).max() < 1e-15 * up


@pytest.mark.parametrize('n,up,down', [(44100, 80, 147), (1000, 3, 2), (999, 2, 3), (7, 160, 441), (48000, 24000, 48000)])
def test_geometry_matches_scipy(n, up, down):
    u, d, n_out, _, _ = wave_io.resample_geometry(n, up, down)
    assert n_out == len(ss.resample_poly(np.zeros(n), up, down))
    assert u * down == up * d


def test_wav_round_trip_and_formats(tmp_path):
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(5000) * 0.2).astype(np.float32)
    p = tmp_path / 'a.wav'
    wave_io.write_wav(p, x, 24000)
    y, sr = wave_io.read_wav(p)
    assert sr == 24000 and np.array_equal(x, y)
    # PCM16 stereo written by hand: read back as the channel mean
    pcm = (np.stack([x, -0.5 * x], 1) * 32767).astype('<i2')
    body = pcm.tobytes()
    hdr = b'RIFF' + struct.pack('<I', 36 + len(body)) + b'WAVEfmt ' + struct.pack('<IHHIIHH', 16, 1, 2, 44100, 44100 * 4, 4, 16) + b'data' + struct.pack('<I', len(body))
    (tmp_path / 'b.wav').write_bytes(hdr + body)
    y, sr = wave_io.read_wav(tmp_path / 'b.wav')
    assert sr == 44100 and np.allclose(y, pcm.astype(np.float32).mean(1) / 32768.0, atol=1e-7)
    with pytest.raises(ValueError):
        (tmp_path / 'c.wav').write_bytes(b'nope' * 8)
        wave_io.read_wav(tmp_path / 'c.wav')


def test_load_wave_resamples_with_engine(tmp_path, small_models):
    """load_wave(path, sampling_rate) = librosa.load(path, sr=...) as check.py:80 uses it; engine replaced by the scipy-backed stand-in."""
    from tests.fake_engine import OracleEngine
    fake = OracleEngine(small_models['stage1_model_path'], small_models['stage2_model_path'])
    t = np.arange(44100) / 44100.0
    x = (0.3 * np.sin(2 * np.pi * 440 * t)).astype(np.float32)
    wave_io.write_wav(tmp_path / 't.wav', x, 44100)
    w = wave_io.load_wave(tmp_path / 't.wav', 24000, engine=fake)
    assert w.sampling_rate == 24000 and len(w.wave) == 24000
    ref = 0.3 * np.sin(2 * np.pi * 440 * np.arange(24000) / 24000.0)
    assert np.abs(w.wave[2000:-2000] - ref[2000:-2000]).max() < 2e-3          # a 440 Hz tone survives 44.1 -> 24 kHz
    same = wave_io.load_wave(tmp_path / 't.wav', None)
    assert same.sampling_rate == 44100 and np.array_equal(same.wave, x)


@pytest.mark.gpu
@pytest.mark.parametrize('n,rate_in,rate_out', [(44100, 44100, 24000), (5000, 16000, 24000), (24000, 48000, 24000), (333, 44100, 24000), (7200, 24000, 16000)])
def test_gpu_resampler_matches_scipy(engine, n, rate_in, rate_out):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) * 0.3).astype(np.float32)
    y = wave_io.resample(x, rate_in, rate_out, engine=engine)
    import math
    g = math.gcd(rate_in, rate_out)
    ref = ss.resample_poly(x.astype(np.float64), rate_out // g, rate_in // g)
    assert len(y) == len(ref)
    assert np.abs(y - ref).max() < 1e-6 * max(1.0, float(np.abs(ref).max()))   # float32 output of an FP64 accumulation
