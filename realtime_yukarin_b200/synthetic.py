"""Seeded synthetic model files, statistics and audio (there is no network: no real checkpoints).

The files are written in the layouts the reference loads (Chainer `save_npz` key names, pickled
{'mean','var'} statistics, config.json with the field names the hot path reads), so the same
directory feeds this package, the oracle and -- were it installed -- the reference itself.
Weights use a variance-preserving (He) initialisation instead of pix2pix's N(0, 0.02) so that every
layer carries O(1) activations: the numerics of the FP16 tensor-core path are then exercised at
realistic magnitudes instead of on a signal that decays to zero.
"""
import json
import pickle
from pathlib import Path
from typing import Dict

import numpy

ENC_MULT = [1, 2, 4, 8, 8, 8, 8, 8]
DEC_MULT = [8, 8, 8, 8, 4, 2, 1]


def _layer_dims(base: int, in_ch: int, out_ch: int):
    dims = []
    for i in range(8):
        cin = in_ch if i == 0 else base * ENC_MULT[i - 1]
        dims.append(('encoder', i, False, cin, base * ENC_MULT[i], 3 if i == 0 else 4))
    for d in range(7):
        cin = base * ENC_MULT[7] if d == 0 else base * DEC_MULT[d - 1] + base * ENC_MULT[7 - d]
        dims.append(('decoder', d, True, cin, base * DEC_MULT[d], 4))
    dims.append(('decoder', 7, False, 2 * base, out_ch, 3))
    return dims


def make_unet_params(seed: int, ndim: int, in_ch: int, out_ch: int, base: int, first_gain: float = 1.0,
                     out_std: float = 1.0, out_bias: float = 0.0) -> Dict[str, numpy.ndarray]:
    rng = numpy.random.default_rng(seed)
    p: Dict[str, numpy.ndarray] = {}
    for part, i, transposed, cin, cout, k in _layer_dims(base, in_ch, out_ch):
        kshape = (k,) * ndim
        taps = (k ** ndim) if not transposed else (k // 2) ** ndim       # taps that reach one output
        fan_in = cin * taps
        plain = (part == 'encoder' and i == 0) or (part == 'decoder' and i == 7)
        if part == 'decoder' and i == 7:
            std = out_std / numpy.sqrt(fan_in)
        else:
            gain = numpy.sqrt(2.0 / (1 + 0.04)) if part == 'encoder' else numpy.sqrt(2.0)
            std = gain / numpy.sqrt(fan_in) * (first_gain if (part == 'encoder' and i == 0) else 1.0)
        shape = (cin, cout) + kshape if transposed else (cout, cin) + kshape
        W = (rng.standard_normal(shape) * std).astype(numpy.float32)
        b = (rng.standard_normal(cout) * 0.02).astype(numpy.float32)
        if part == 'decoder' and i == 7:
            b = b + numpy.float32(out_bias)
        if plain:
            p[f'{part}/c{i}/W'] = W
            p[f'{part}/c{i}/b'] = b
        else:
            p[f'{part}/c{i}/c/W'] = W
            p[f'{part}/c{i}/c/b'] = b
            p[f'{part}/c{i}/batchnorm/gamma'] = (1 + 0.05 * rng.standard_normal(cout)).astype(numpy.float32)
            p[f'{part}/c{i}/batchnorm/beta'] = (0.05 * rng.standard_normal(cout)).astype(numpy.float32)
            p[f'{part}/c{i}/batchnorm/avg_mean'] = (0.1 * rng.standard_normal(cout)).astype(numpy.float32)
            p[f'{part}/c{i}/batchnorm/avg_var'] = rng.uniform(0.8, 1.2, cout).astype(numpy.float32)
    return p


MC_MEAN_IN = numpy.array([-4.6, 1.6, 0.45, 0.25, 0.1, 0.05, 0.0, 0.0, 0.0], numpy.float32)
MC_STD_IN = numpy.array([1.6, 0.6, 0.4, 0.3, 0.25, 0.2, 0.2, 0.15, 0.15], numpy.float32)
MC_MEAN_OUT = numpy.array([-4.2, 1.4, 0.55, 0.2, 0.12, 0.02, 0.03, 0.0, 0.01], numpy.float32)
MC_STD_OUT = numpy.array([0.9, 0.4, 0.3, 0.22, 0.2, 0.15, 0.15, 0.12, 0.12], numpy.float32)


def make_stage1_params(seed: int = 0, base: int = 64, channels: int = 9) -> Dict[str, numpy.ndarray]:
    p = make_unet_params(seed + 101, ndim=1, in_ch=channels, out_ch=channels, base=base, out_std=1.0)
    p['stats/in_mean'], p['stats/in_std'] = MC_MEAN_IN[:channels].copy(), MC_STD_IN[:channels].copy()
    p['stats/out_mean'], p['stats/out_std'] = MC_MEAN_OUT[:channels].copy(), MC_STD_OUT[:channels].copy()
    return p


def make_stage2_params(seed: int = 0, base: int = 64) -> Dict[str, numpy.ndarray]:
    # input: log power spectrum around -9 +- 3 (silent frames: ln 1e-16 = -36.8); first_gain keeps every layer at O(1) RMS so that the
    # output log spectrum stays around -6 +- 1.5 and the synthesized audio at a realistic level (RMS ~0.05): the 1e-3 sample-RMSE
    # tolerance of north_star is an absolute figure on normalised audio
    return make_unet_params(seed + 202, ndim=2, in_ch=1, out_ch=1, base=base, first_gain=0.035, out_std=1.2, out_bias=-8.5)


def write_synthetic_models(directory, seed: int = 0, base1: int = 64, base2: int = 64) -> Dict[str, Path]:
    d = Path(directory)
    (d / 'model_stage1').mkdir(parents=True, exist_ok=True)
    (d / 'model_stage2').mkdir(parents=True, exist_ok=True)
    paths = dict(
        input_statistics_path=d / 'input_statistics.npy', target_statistics_path=d / 'target_statistics.npy',
        stage1_model_path=d / 'model_stage1' / 'predictor.npz', stage1_config_path=d / 'model_stage1' / 'config.json',
        stage2_model_path=d / 'model_stage2' / 'predictor.npz', stage2_config_path=d / 'model_stage2' / 'config.json')
    with open(paths['input_statistics_path'], 'wb') as f:
        numpy.save(f, numpy.array({'mean': float(numpy.log(150.0)), 'var': 0.04}, dtype=object), allow_pickle=True)
    with open(paths['target_statistics_path'], 'wb') as f:
        numpy.save(f, numpy.array({'mean': float(numpy.log(250.0)), 'var': 0.04}, dtype=object), allow_pickle=True)
    numpy.savez(paths['stage1_model_path'], **make_stage1_params(seed, base1))
    numpy.savez(paths['stage2_model_path'], **make_stage2_params(seed, base2))
    acoustic_param = dict(sampling_rate=24000, pad_second=0, threshold_db=None, frame_period=5, order=8, alpha=0.466,
                          f0_floor=71.0, f0_ceil=800.0, fft_length=1024, dtype='float32')
    paths['stage1_config_path'].write_text(json.dumps(dict(
        dataset=dict(acoustic_param=acoustic_param, in_features=['mc'], out_features=['mc']),
        model=dict(in_channels=9, out_channels=9, generator_base_channels=base1, generator_extensive_layers=8)), indent=1))
    paths['stage2_config_path'].write_text(json.dumps(dict(
        dataset=dict(param=dict(voice_param=dict(sample_rate=24000, top_db=None, pad_second=0.0),
                                acoustic_feature_param=dict(frame_period=5, order=8, alpha=0.466, f0_estimating_method='dio'))),
        model=dict(generator_base_channels=base2)), indent=1))
    return paths


def make_crepe_params(seed: int = 0, capacity: str = 'tiny') -> Dict[str, numpy.ndarray]:
    """Seeded random weights of the CREPE architecture (He-style scales, mildly perturbed BatchNorm statistics): the trained weights
    cannot be fetched here, so parity tests exercise the architecture and the decoders, not pitch accuracy."""
    mult = {'tiny': 4, 'small': 8, 'medium': 16, 'large': 24, 'full': 32}[capacity]
    rng = numpy.random.default_rng(9000 + seed)
    filters = [n * mult for n in (32, 4, 4, 4, 8, 16)]
    widths = [512, 64, 64, 64, 64, 64]
    p = {}
    cin = 1
    for l, (f, w) in enumerate(zip(filters, widths)):
        p[f'conv{l + 1}.W'] = (rng.standard_normal((f, cin, w)) * numpy.sqrt(2.0 / (cin * w))).astype(numpy.float32)
        p[f'conv{l + 1}.b'] = (0.05 * rng.standard_normal(f)).astype(numpy.float32)
        p[f'bn{l + 1}.gamma'] = (1.0 + 0.1 * rng.standard_normal(f)).astype(numpy.float32)
        p[f'bn{l + 1}.beta'] = (0.1 * rng.standard_normal(f)).astype(numpy.float32)
        p[f'bn{l + 1}.mean'] = (0.3 + 0.1 * rng.standard_normal(f)).astype(numpy.float32)
        p[f'bn{l + 1}.var'] = (0.5 + 0.2 * rng.random(f)).astype(numpy.float32)
        cin = f
    p['dense.W'] = (rng.standard_normal((360, 4 * filters[5])) * numpy.sqrt(0.5 / (4 * filters[5]))).astype(numpy.float32)
    p['dense.b'] = (0.5 * rng.standard_normal(360)).astype(numpy.float32)
    return p


def write_crepe_model(directory, seed: int = 0, capacity: str = 'tiny') -> Path:
    d = Path(directory)
    d.mkdir(parents=True, exist_ok=True)
    path = d / f'crepe_{capacity}.npz'
    numpy.savez(path, **make_crepe_params(seed, capacity))
    return path


def synthetic_speech(seconds: float, stream: int = 0, fs: int = 24000, silence_fraction: float = 0.2) -> numpy.ndarray:
    """Voiced harmonic source with a random-walk f0 (100-300 Hz), 20 harmonics with 1/h roll-off,
    amplitude 0.1-0.3, -40 dB white noise and ~20 % silent gaps."""
    rng = numpy.random.default_rng(1234 + stream)
    n = int(round(seconds * fs))
    hop = fs // 100
    nseg = n // hop + 2
    steps = rng.standard_normal(nseg) * 4.0
    f0_coarse = numpy.empty(nseg)
    f = rng.uniform(120, 260)
    for i in range(nseg):
        f = min(300.0, max(100.0, f + steps[i]))
        f0_coarse[i] = f
    f0 = numpy.interp(numpy.arange(n) / hop, numpy.arange(nseg), f0_coarse)
    phase = 2 * numpy.pi * numpy.cumsum(f0) / fs
    x = numpy.zeros(n)
    for h in range(1, 21):
        x += numpy.sin(h * phase + rng.uniform(0, 2 * numpy.pi)) / h
    amp_coarse = rng.uniform(0.1, 0.3, nseg)
    amp = numpy.interp(numpy.arange(n) / hop, numpy.arange(nseg), amp_coarse)
    x = x / numpy.max(numpy.abs(x)) * amp
    # silent gaps
    gate = numpy.ones(n)
    t = 0
    while t < n:
        voiced_len = int(rng.uniform(0.25, 0.9) * fs)
        silent_len = int(voiced_len * silence_fraction / (1 - silence_fraction) * rng.uniform(0.5, 1.5))
        t += voiced_len
        gate[t:t + silent_len] = 0.0
        t += silent_len
    ramp = int(0.005 * fs)
    kernel = numpy.ones(ramp) / ramp
    gate = numpy.convolve(gate, kernel, mode='same')
    x = x * gate + rng.standard_normal(n) * 10 ** (-40 / 20) * 0.1
    return x.astype(numpy.float32)
