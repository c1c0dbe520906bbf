"""CPU restatement of the CREPE f0 front-end -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Reference call site: realtime_voice_conversion/yukarin_wrapper/acoustic_feature_wrapper.py:65-80
    t, f0, confidence, _ = crepe.predict(x, fs, viterbi=True, model_capacity='full', step_size=frame_period, verbose=0)
    voiced = (crepe.predict_voicing(confidence) == 1) | (confidence > 0.1);  f0[~voiced] = 0
`crepe` (marl/crepe, un-vendored, unpinned: requirements.txt) is absent here, and so are its trained weights: PARITY UNPINNED.  This
file restates the published package (crepe/core.py, v0.0.12+: build_and_load_model, get_activation, to_local_average_cents,
to_viterbi_cents, predict, predict_voicing) and hmmlearn's Viterbi decoder (_hmmc._viterbi: log domain, first maximum wins):
  * 1024-sample frames of the 16 kHz signal (zero-padded by 512 on both sides), hop = int(16000 * step_size / 1000), each frame
    normalised to zero mean / unit standard deviation (std clipped at 1e-8);
  * six blocks Conv2D(f, (w, 1), stride, 'same', relu) -> BatchNorm (eps 1e-3) -> MaxPool(2) [-> Dropout: identity at inference],
    filters m * [32, 4, 4, 4, 8, 16] (m = 32 for 'full'), widths [512, 64, 64, 64, 64, 64], strides [4, 1, 1, 1, 1, 1];
    flatten (time-major) -> Dense(360, sigmoid);
  * pitch path = Viterbi over 360 bins (banded triangular transitions of half-width 12, emission 0.1 self + 0.9 / 360) observed
    through the per-frame argmax; cents = activation-weighted mean over +-4 bins around the path; f0 = 10 * 2 ** (cents / 1200);
  * voicing = Viterbi of a 2-state Gaussian HMM (means 0 / 1, variance 0.25, self transition 0.99) on the confidence.
DECIDE C1: the 24 kHz -> 16 kHz resampling (crepe uses resampy's kaiser_best) is the package's polyphase resampler (scipy.signal.
resample_poly), the same substitution as for librosa.load; callers hand 16 kHz audio to `predict`.
DECIDE C2: weights are a plain npz {conv<l>.W (cout, cin, k), conv<l>.b, bn<l>.gamma/beta/mean/var, dense.W (360, 64 m), dense.b};
tests use seeded synthetic weights (realtime_yukarin_b200.synthetic.write_crepe_model).
"""
from typing import Dict

import numpy as np

MODEL_SRATE = 16000
FILTERS = [32, 4, 4, 4, 8, 16]
WIDTHS = [512, 64, 64, 64, 64, 64]
STRIDES = [4, 1, 1, 1, 1, 1]
BN_EPS = 1e-3
CENTS_MAPPING = np.linspace(0, 7180, 360) + 1997.3794084376191


def same_padding(n_in: int, k: int, stride: int):
    """TensorFlow 'same': out = ceil(n / s); total padding split with the extra sample on the right."""
    n_out = -(-n_in // stride)
    total = max((n_out - 1) * stride + k - n_in, 0)
    return n_out, total // 2, total - total // 2


def frames_of(audio16k: np.ndarray, step_ms: float) -> np.ndarray:
    audio = np.pad(np.asarray(audio16k, np.float32), 512, mode='constant', constant_values=0)
    hop = int(MODEL_SRATE * step_ms / 1000)
    n_frames = 1 + int((len(audio) - 1024) / hop)
    idx = np.arange(1024)[None, :] + hop * np.arange(n_frames)[:, None]
    frames = audio[idx].copy()
    frames -= np.mean(frames, axis=1)[:, np.newaxis]
    frames /= np.clip(np.std(frames, axis=1)[:, np.newaxis], 1e-8, None)
    return frames


def get_activation(audio16k: np.ndarray, weights: Dict[str, np.ndarray], step_ms: float = 10.0) -> np.ndarray:
    import torch
    import torch.nn.functional as F
    x = torch.from_numpy(frames_of(audio16k, step_ms))[:, None, :]            # (frames, 1, 1024)
    with torch.no_grad():
        for l in range(6):
            W = torch.from_numpy(np.asarray(weights[f'conv{l + 1}.W'], np.float32))
            b = torch.from_numpy(np.asarray(weights[f'conv{l + 1}.b'], np.float32))
            _, left, right = same_padding(x.shape[2], WIDTHS[l], STRIDES[l])
            x = F.conv1d(F.pad(x, (left, right)), W, b, stride=STRIDES[l])
            x = torch.relu(x)
            g, beta = weights[f'bn{l + 1}.gamma'], weights[f'bn{l + 1}.beta']
            mean, var = weights[f'bn{l + 1}.mean'], weights[f'bn{l + 1}.var']
            a = (np.asarray(g, np.float64) / np.sqrt(np.asarray(var, np.float64) + BN_EPS)).astype(np.float32)
            c = (np.asarray(beta, np.float64) - np.asarray(mean, np.float64) * a).astype(np.float32)
            x = x * torch.from_numpy(a)[None, :, None] + torch.from_numpy(c)[None, :, None]
            x = F.max_pool1d(x, 2)
        flat = x.permute(0, 2, 1).reshape(x.shape[0], -1)                       # time-major flatten: index = t * C + c
        z = flat @ torch.from_numpy(np.asarray(weights['dense.W'], np.float32)).T + torch.from_numpy(np.asarray(weights['dense.b'], np.float32))
        return torch.sigmoid(z).numpy()


def to_local_average_cents(salience: np.ndarray, center=None) -> float:
    if center is None:
        center = int(np.argmax(salience))
    start, end = max(0, center - 4), min(len(salience), center + 5)
    s = salience[start:end]
    return float(np.sum(s * CENTS_MAPPING[start:end]) / np.sum(s))


def pitch_hmm_tables():
    """(log start [360], log transition [360][360], log emission of the observed bin: (self, other))."""
    with np.errstate(divide='ignore'):
        xx, yy = np.meshgrid(range(360), range(360))
        transition = np.maximum(12 - abs(xx - yy), 0).astype(np.float64)
        transition = transition / np.sum(transition, axis=1)[:, None]
        self_emission = 0.1
        e_self = self_emission + (1 - self_emission) / 360
        e_other = (1 - self_emission) / 360
        return np.log(np.ones(360) / 360), np.log(transition), np.log(np.array([e_self, e_other]))


def viterbi(log_start: np.ndarray, log_trans: np.ndarray, frame_logprob: np.ndarray) -> np.ndarray:
    """hmmlearn _hmmc._viterbi: lattice in the log domain, np.argmax tie-breaking (first maximum)."""
    T, n = frame_logprob.shape
    lattice = np.empty((T, n))
    lattice[0] = log_start + frame_logprob[0]
    for t in range(1, T):
        lattice[t] = np.max(lattice[t - 1][:, None] + log_trans, axis=0) + frame_logprob[t]
    path = np.empty(T, np.int64)
    path[T - 1] = where = int(np.argmax(lattice[T - 1]))
    for t in range(T - 2, -1, -1):
        path[t] = where = int(np.argmax(lattice[t] + log_trans[:, where]))
    return path


def to_viterbi_cents(salience: np.ndarray) -> np.ndarray:
    log_start, log_trans, log_emit = pitch_hmm_tables()
    obs = np.argmax(salience, axis=1)
    frame = np.full((len(obs), 360), log_emit[1])
    frame[np.arange(len(obs)), obs] = log_emit[0]
    path = viterbi(log_start, log_trans, frame)
    return np.array([to_local_average_cents(salience[i, :], int(path[i])) for i in range(len(obs))]), path


def predict_voicing(confidence: np.ndarray) -> np.ndarray:
    log_start = np.log(np.array([0.5, 0.5]))
    log_trans = np.log(np.array([[0.99, 0.01], [0.01, 0.99]]))
    c = np.asarray(confidence, np.float64)[:, None]
    means, var = np.array([0.0, 1.0])[None, :], 0.25
    frame = -0.5 * (np.log(2 * np.pi) + np.log(var) + (c - means) ** 2 / var)
    return viterbi(log_start, log_trans, frame)


def predict(audio16k: np.ndarray, weights: Dict[str, np.ndarray], step_ms: float = 5.0, viterbi_path: bool = True):
    """crepe.predict on 16 kHz audio: (time, frequency, confidence, activation)."""
    activation = get_activation(audio16k, weights, step_ms)
    confidence = activation.max(axis=1)
    if viterbi_path:
        cents, _ = to_viterbi_cents(activation)
    else:
        cents = np.array([to_local_average_cents(activation[i]) for i in range(len(activation))])
    frequency = 10 * 2 ** (cents / 1200)
    frequency[np.isnan(frequency)] = 0
    time = np.arange(confidence.shape[0]) * step_ms / 1000.0
    return time, frequency, confidence, activation


def extract_f0(audio16k: np.ndarray, weights: Dict[str, np.ndarray], frame_period: float):
    """CrepeAcousticFeatureWrapper.extract_f0 (acoustic_feature_wrapper.py:66-80) after the resampling step."""
    t, f0, confidence, _ = predict(audio16k, weights, step_ms=frame_period, viterbi_path=True)
    voiced = (predict_voicing(confidence) == 1) | (confidence > 0.1)
    f0 = f0.copy()
    f0[~voiced] = 0
    return f0, t
