"""The CPU oracle checked against independent known answers (numpy.fft, closed forms, published
constants) and against the committed golden fixtures.  The reference ships no golden vectors for
this arithmetic (PARITY UNPINNED), so these tests pin the restatement to WORLD's / SPTK's published
behaviour as far as that can be stated without the upstream sources."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import nets as onets
from oracle import pipeline as opipe
from oracle import world as W
from realtime_yukarin_b200 import synthetic

GOLDEN = Path(__file__).resolve().parent / 'golden'
FS = 24000


def harmonic(f0_track, amp=0.1, nharm=20, fs=FS, seed=0):
    rng = np.random.default_rng(seed)
    ph = 2 * np.pi * np.cumsum(f0_track) / fs
    x = sum(np.sin(h * ph) / h for h in range(1, nharm + 1)) * amp
    return x + 1e-4 * rng.standard_normal(len(x))


def test_fft_matches_numpy():
    rng = np.random.default_rng(0)
    for n in (64, 1024, 2048, 16384):
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        assert np.abs(W.fft(x) - np.fft.fft(x)).max() < 1e-10 * n
        assert np.abs(W.fft(x, +1) - np.fft.ifft(x) * n).max() < 1e-10 * n


def test_interp1_matches_numpy_inside_and_extrapolates_outside():
    rng = np.random.default_rng(1)
    x = np.sort(rng.uniform(0, 10, 20))
    y = rng.standard_normal(20)
    xi = np.linspace(x[0], x[-1], 57)
    assert np.abs(W.interp1(x, y, xi) - np.interp(xi, x, y)).max() < 1e-12
    lo = W.interp1(x, y, np.array([x[0] - 1.0]))[0]
    assert abs(lo - (y[0] + (y[1] - y[0]) / (x[1] - x[0]) * -1.0)) < 1e-12        # matlab interp1 'linear', 'extrap'


def test_randn_is_marsaglia_xorshift128():
    # first xor128 output with Marsaglia's seeds is 3701687786; WORLD sums 12 outputs >> 4
    x, y, z, w = 123456789, 362436069, 521288629, 88675123
    tot = 0
    first = None
    for _ in range(12):
        t = (x ^ (x << 11)) & 0xFFFFFFFF
        x, y, z = y, z, w
        w = ((w ^ (w >> 19)) ^ (t ^ (t >> 8))) & 0xFFFFFFFF
        first = w if first is None else first
        tot += w >> 4
    assert first == 3701687786
    r = W.randn_stream(0, 1000)
    assert abs(r[0] - (tot / 268435456.0 - 6.0)) < 1e-15
    assert abs(r.mean()) < 0.15 and abs(r.std() - 1.0) < 0.1
    assert np.array_equal(W.randn_stream(37, 5), W.randn_stream(0, 42)[37:])


def test_dio_stonemask_track_a_known_f0():
    t = np.arange(int(FS * 1.0)) / FS
    track = 180 + 40 * np.sin(2 * np.pi * 1.2 * t)
    x = harmonic(track)
    f0, tp = W.dio(x, FS)
    assert len(f0) == 201 and np.allclose(tp, np.arange(201) * 0.005)
    f0r = W.stonemask(x, FS, tp, f0)
    true = 180 + 40 * np.sin(2 * np.pi * 1.2 * tp)
    inner = slice(10, -10)
    assert (f0r[inner] > 0).all()
    assert np.abs(f0[inner] - true[inner]).max() / 180 < 0.03           # DIO: coarse
    assert np.median(np.abs(f0r[inner] - true[inner])) < 0.3            # StoneMask: refined to a fraction of a Hz
    silence = np.zeros(FS // 2)
    f0s, _ = W.dio(silence, FS)
    assert (f0s == 0).all()


def test_cheaptrick_envelope_and_d4c_aperiodicity_sanity():
    t = np.arange(FS // 2) / FS
    x = harmonic(np.full(len(t), 150.0), amp=0.2)
    f0, tp = W.dio(x, FS)
    f0 = W.stonemask(x, FS, tp, f0)
    sp = W.cheaptrick(x, FS, tp, f0)
    ap = W.d4c(x, FS, tp, f0)
    assert sp.shape == (len(f0), 513) and np.isfinite(sp).all() and (sp > 0).all()
    mid = sp[len(f0) // 2]
    # 1/h harmonic amplitudes -> envelope falls with frequency; smooth (no harmonic ripple left)
    assert mid[10] > mid[60] > mid[120]
    assert np.abs(np.diff(np.log(mid[5:130]))).max() < 0.5
    a = ap[len(f0) // 2]
    assert a[0] == pytest.approx(0.001) and a[5:100].max() < 0.2        # clean periodic source: low aperiodicity in band
    noise = np.random.default_rng(3).standard_normal(FS // 2) * 0.05
    apn = W.d4c(noise, FS, tp, np.full(len(tp), 150.0))
    assert apn[len(tp) // 2, 50:400].mean() > 0.5                        # noise analysed as if voiced: high aperiodicity
    assert np.all(W.d4c(x, FS, tp, np.zeros(len(tp))) == 1.0 - 1e-12)   # unvoiced frames


def _freqt_ref(c, order, a):
    """Independent (textbook, non-in-place) frequency transform by the all-pass chain, O(n * order)."""
    g = np.zeros(order + 1)
    for ci in c[::-1]:
        d = g.copy()
        g[0] = ci + a * d[0]
        if order >= 1:
            g[1] = (1 - a * a) * d[0] + a * d[1]
        for j in range(2, order + 1):
            g[j] = d[j - 1] + a * (d[j] - g[j - 1])
    return g


def test_sptk_conversions():
    rng = np.random.default_rng(4)
    c = rng.standard_normal(40) * 0.5 ** np.arange(40)
    assert np.abs(W.freqt(c, 8, 0.466) - _freqt_ref(c, 8, 0.466)).max() < 1e-12
    # freqt with alpha = 0 is a truncation
    assert np.allclose(W.freqt(c, 8, 0.0), c[:9])
    # sp2mc of a flat spectrum P: c0 = log(P) / 2, rest 0 ; mc2sp inverts it
    mc = W.sp2mc(np.full((3, 513), 0.01), 8, 0.466)
    assert np.allclose(mc[:, 0], np.log(0.01) / 2) and np.abs(mc[:, 1:]).max() < 1e-12
    assert np.allclose(W.mc2sp(mc, 0.466, 1024), 0.01)
    # a smooth (low-quefrency) envelope survives the order-8 round trip closely
    k = np.arange(513)
    logsp = -8 + 2.0 * np.cos(np.pi * k / 512) + 0.5 * np.cos(2 * np.pi * k / 512)
    sp = np.exp(logsp)[None]
    back = W.mc2sp(W.sp2mc(sp, 24, 0.0), 0.0, 1024)
    assert np.abs(np.log(back) - logsp).max() < 1e-9


def test_frame_mse_matches_numpy_reflect_framing():
    rng = np.random.default_rng(5)
    x = rng.standard_normal(2000).astype(np.float32)
    fl, hop = 256, 100
    xp = np.pad(x.astype(np.float64), fl // 2, mode='reflect')
    n = 1 + len(x) // hop
    ref = np.array([np.mean(xp[i * hop:i * hop + fl] ** 2) for i in range(n)])
    assert np.allclose(W.frame_mse(x, fl, hop, n), ref, rtol=1e-12)
    cfg = opipe.PathConfig()
    sil = np.zeros(24000, np.float32)
    sil[6000:12000] = 0.1 * rng.standard_normal(6000)
    m = opipe.effective_mask(sil, 200, cfg, 60.0)
    assert m[60:95].all() and not m[:40].any() and not m[110:].any()
    assert opipe.effective_mask(sil, 200, cfg, None).all()


def test_realtime_synthesizer_contract():
    fft = W.cheaptrick_fft_size(FS)
    assert fft == 1024
    nb = fft // 2 + 1
    s = W.RealtimeSynthesizer(FS, 5.0, fft, 1024)
    f0 = np.full(200, 200.0)
    sp = np.full((200, nb), 1e-4, np.float32)
    ap = np.full((200, nb), 0.01, np.float32)
    y = s.decode(f0, sp, ap)
    idx, tm, vuv = s.pulses()
    assert len(y) % 1024 == 0 and len(y) >= 199 * 120 - 2 * 1024                     # whole blocks only
    assert (vuv == 1).all() and np.all(np.abs(np.diff(idx) - FS / 200.0) <= 1)       # one pulse per period
    # chunked feeding == one-shot feeding (state hand-off between AddParameters calls)
    s2 = W.RealtimeSynthesizer(FS, 5.0, fft, 1024)
    ys = [s2.decode(f0[a:a + 60], sp[a:a + 60], ap[a:a + 60]) for a in range(0, 200, 60)]
    y2 = np.concatenate(ys)
    n = min(len(y), len(y2))
    assert n > 15000 and np.abs(y[:n] - y2[:n]).max() < 1e-9 * np.abs(y).max() + 1e-12
    # unvoiced: pulses at the 500 Hz default rate, aperiodic only; sp = 0 -> NaN (scrubbed by DecodeStream)
    s3 = W.RealtimeSynthesizer(FS, 5.0, fft, 1024)
    y3 = s3.decode(np.zeros(100), np.zeros((100, nb), np.float32), np.zeros((100, nb), np.float32))
    i3, _, v3 = s3.pulses()
    assert (v3 == 0).all() and np.all(np.abs(np.diff(i3) - 48) <= 1) and np.isnan(y3).any()


def test_analysis_synthesis_round_trip():
    x = synthetic.synthetic_speech(1.0, stream=9).astype(np.float64)
    cfg = opipe.PathConfig()
    f = opipe.extract_features(x, cfg)
    s = W.RealtimeSynthesizer(FS, 5.0, 1024, 1024)
    y = s.decode(f['f0'].ravel().astype(np.float64), f['sp'], f['ap'])
    n = min(len(x), len(y))
    X = np.abs(np.fft.rfft(x[2000:2000 + 8192] * np.hanning(8192)))
    Y = np.abs(np.fft.rfft(y[2000:2000 + 8192] * np.hanning(8192)))
    corr = np.corrcoef(np.log(X[20:2500] + 1e-9), np.log(Y[20:2500] + 1e-9))[0, 1]
    assert corr > 0.8
    assert 0.5 < np.sqrt(np.mean(y[:n] ** 2)) / np.sqrt(np.mean(x[:n] ** 2)) < 2.0


def test_unet_backends_agree_and_match_plain_torch_layers(small_models):
    p1 = onets.load_npz(small_models['stage1_model_path'])
    rng = np.random.default_rng(6)
    x = rng.standard_normal((9, 128)).astype(np.float32)
    a = onets.unet_forward(x, p1, 1, 'numpy')
    b = onets.unet_forward(x, p1, 1, 'torch')
    assert a.shape == (9, 128) and np.abs(a - b).max() < 1e-4
    mc = (synthetic.MC_MEAN_IN + synthetic.MC_STD_IN * rng.standard_normal((60, 9))).astype(np.float32)
    y = onets.stage1_convert(mc, p1, 'torch')
    assert y.shape == (60, 9) and np.isfinite(y).all()
    # 'minimum' padding: appending frames that equal the per-channel minimum changes nothing for T -> T' in the same 128-bucket
    p2 = onets.load_npz(small_models['stage2_model_path'])
    sp = np.exp(-9 + rng.standard_normal((20, 513))).astype(np.float32)
    out = onets.stage2_convert(sp, p2, 'torch')
    assert out.shape == (20, 513) and (out > 0).all()
    assert np.array_equal(out[:, 512], out[:, 511])                       # edge-padded Nyquist bin


def test_golden_fixtures():
    """Regression pin: committed vectors made by tests/golden/make_golden.py from this oracle."""
    meta = json.loads((GOLDEN / 'golden_meta.json').read_text())
    z = np.load(GOLDEN / 'golden_small.npz')
    x = synthetic.synthetic_speech(meta['seconds'], stream=meta['stream'])
    assert np.array_equal(x, z['wave'])
    cfg = opipe.PathConfig()
    f = opipe.extract_features(x, cfg)
    assert np.array_equal(f['voiced'].ravel(), z['voiced'])
    assert np.allclose(f['f0'].ravel(), z['f0'], rtol=1e-6)
    assert np.allclose(np.log(f['sp'][:, ::16]), z['log_sp_sub'], atol=1e-4)
    assert np.allclose(f['ap'][:, ::16], z['ap_sub'], atol=1e-5)
    assert np.allclose(f['mc'], z['mc'], atol=1e-4)
    s = W.RealtimeSynthesizer(FS, 5.0, 1024, 1024)
    y = s.decode(f['f0'].ravel().astype(np.float64), f['sp'], f['ap'])
    assert len(y) == len(z['resynth']) and np.abs(y - z['resynth']).max() < 1e-4 * np.abs(z['resynth']).max()


def test_golden_fixtures_widen():
    """Regression pin of the SURVEY 8(f) oracle rows (offline Synthesis, output gate, re-blocker): tests/golden/golden_widen.npz."""
    meta = json.loads((GOLDEN / 'golden_meta.json').read_text())
    z = np.load(GOLDEN / 'golden_widen.npz')
    x = synthetic.synthetic_speech(meta['seconds'], stream=meta['stream'])
    f = opipe.extract_features(x, opipe.PathConfig())
    yo, pidx, pshift, pvuv = W.synthesize(f['f0'].ravel().astype(np.float64), f['sp'], f['ap'], 24000, 5.0, return_pulses=True)
    assert np.array_equal(pidx, z['pulse_index']) and np.array_equal(pvuv, z['pulse_vuv'])
    assert np.allclose(pshift, z['pulse_shift'], rtol=0, atol=1e-15)
    assert len(yo) == len(z['offline']) and np.abs(yo - z['offline']).max() < 1e-6 * max(1.0, np.abs(z['offline']).max())
    powers = np.array([W.stft_power_db_mean(yo[a:a + 7200]) for a in range(0, len(yo) - 7199, 7200)])
    assert np.allclose(powers, z['gate_power'], rtol=0, atol=1e-6)
    s = W.RealtimeSynthesizer(FS, 5.0, 1024, 1024)
    y = s.decode(f['f0'].ravel().astype(np.float64), f['sp'], f['ap'])
    rb = opipe.OutputReblockOracle(7200, 30.0)
    statuses = [rb.push(y[a:a + 1024] if a // 1024 % 5 else 1e-6 * y[a:a + 1024])[0] for a in range(0, len(y), 1024)]
    assert statuses == z['reblock_status'].tolist()


def test_cheaptrick_and_d4c_agree_with_independent_numpy_writings():
    """oracle/world_oracle.c (C, scalar loops) vs tests/independent_world.py (numpy, written separately from the published
    algorithm): spectral envelope within 1e-8 in the log domain (running sums over 1e-7-level bins), aperiodicity within 1e-10, on voiced and unvoiced frames."""
    from tests import independent_world as iw
    x = synthetic.synthetic_speech(0.6, stream=5).astype(np.float64)
    f0, t = W.dio(x, FS)[:2]
    f0 = W.stonemask(x, FS, t, f0)
    assert (f0 > 0).sum() > 20 and (f0 == 0).sum() > 0
    sp = W.cheaptrick(x, FS, t, f0)
    assert np.abs(np.log(iw.cheaptrick_np(x, FS, t, f0)) - np.log(sp)).max() < 1e-8
    ap = W.d4c(x, FS, t, f0)
    ap = ap[0] if isinstance(ap, tuple) else ap
    assert np.abs(iw.d4c_np(x, FS, t, f0) - ap).max() < 1e-10


@pytest.mark.parametrize('stream', [5, 9, 13])
def test_dio_and_stonemask_agree_with_independent_numpy_writings(stream):
    """DIO (filter bank, four zero-crossing trains, candidate scoring, four-step contour repair) and StoneMask, C oracle vs
    tests/independent_world.py: same voiced decisions, f0 within 1e-9 Hz."""
    from tests import independent_world as iw
    x = synthetic.synthetic_speech(0.8, stream=stream).astype(np.float64)
    f0_ref, t_ref = W.dio(x, FS)[:2]
    f0, t = iw.dio_np(x, FS)
    assert np.array_equal(t, t_ref) and np.array_equal(f0 > 0, f0_ref > 0) and (f0_ref > 0).sum() > 50
    assert np.abs(f0 - f0_ref).max() < 1e-9
    assert np.abs(iw.stonemask_np(x, FS, t_ref, f0_ref) - W.stonemask(x, FS, t_ref, f0_ref)).max() < 1e-9


def test_realtime_synthesizer_agrees_with_an_independent_numpy_writing():
    """The stateful realtime synthesizer (hand-off phase / f0 between AddParameters calls, blocked phase sum, pulse detection, block
    emission rule, position-addressed noise, half-length dc-remover): C oracle vs tests/independent_world.NumpyRealtimeSynth fed the
    same 60-frame chunks -- identical pulses, identical block counts, samples within 1e-12."""
    from tests.independent_world import NumpyRealtimeSynth
    x = synthetic.synthetic_speech(1.2, stream=7)
    f = opipe.extract_features(x, opipe.PathConfig())
    ref, mine = W.RealtimeSynthesizer(FS, 5.0, 1024, 1024), NumpyRealtimeSynth(FS, 5.0, 1024, 1024)
    total = 0
    for a in range(0, len(f['f0']), 60):
        f0 = f['f0'][a:a + 60].ravel().astype(np.float64)
        yr, ym = ref.decode(f0, f['sp'][a:a + 60], f['ap'][a:a + 60]), mine.decode(f0, f['sp'][a:a + 60], f['ap'][a:a + 60])
        assert len(yr) == len(ym)
        if len(yr):
            assert np.abs(yr - ym).max() < 1e-12
        total += len(yr)
    idx, _, vuv = ref.pulses()
    assert total >= 4 * 1024 and len(idx) > 100
    assert np.array_equal(idx, [p[0] for p in mine.pulses]) and np.array_equal(vuv, [p[2] for p in mine.pulses])
