"""CREPE restatement (oracle/crepe.py; PARITY UNPINNED: neither the crepe package nor its weights are available here): the pieces
that can be pinned independently are -- the Viterbi decoder against exhaustive search, TensorFlow 'same' padding arithmetic, the local
cents average, the voicing HMM on hand-made confidences, the frame / normalisation step and the network's tensor shapes."""
import itertools

import numpy as np

from oracle import crepe as oc
from realtime_yukarin_b200 import synthetic


def test_viterbi_equals_exhaustive_search():
    rng = np.random.default_rng(0)
    for n, T in ((3, 5), (4, 4), (2, 7)):
        start = np.log(rng.dirichlet(np.ones(n)))
        trans = np.log(rng.dirichlet(np.ones(n), size=n))
        frame = np.log(rng.random((T, n)))
        path = oc.viterbi(start, trans, frame)
        best, best_p = -np.inf, None
        for p in itertools.product(range(n), repeat=T):
            s = start[p[0]] + frame[0, p[0]] + sum(trans[p[t - 1], p[t]] + frame[t, p[t]] for t in range(1, T))
            if s > best + 1e-12:
                best, best_p = s, p
        assert tuple(path) == best_p


def test_same_padding_is_tensorflows():
    assert oc.same_padding(1024, 512, 4) == (256, 254, 254)
    assert oc.same_padding(128, 64, 1) == (128, 31, 32)
    assert oc.same_padding(8, 64, 1) == (8, 31, 32)


def test_local_average_and_cents_mapping():
    s = np.zeros(360, np.float32); s[100] = 1.0
    assert abs(oc.to_local_average_cents(s) - oc.CENTS_MAPPING[100]) < 1e-9
    s[101] = 1.0
    assert abs(oc.to_local_average_cents(s) - 0.5 * (oc.CENTS_MAPPING[100] + oc.CENTS_MAPPING[101])) < 1e-9
    assert abs(oc.CENTS_MAPPING[1] - oc.CENTS_MAPPING[0] - 20.0) < 1e-9          # 20-cent bins
    assert abs(10 * 2 ** (oc.CENTS_MAPPING[0] / 1200) - 31.7) < 0.1               # C1 ~ 32.7 Hz region


def test_voicing_hmm_keeps_state_through_short_dips():
    conf = np.array([0.9] * 20 + [0.3] * 2 + [0.9] * 20 + [0.05] * 40)
    v = oc.predict_voicing(conf)
    assert np.all(v[:42] == 1)            # a two-frame dip does not leave the voiced state (self transition 0.99)
    assert np.all(v[-30:] == 0)


def test_frames_and_network_shapes():
    rng = np.random.default_rng(1)
    x = rng.standard_normal(4800).astype(np.float32)
    fr = oc.frames_of(x, 5.0)
    assert fr.shape == (61, 1024)
    assert np.allclose(fr.mean(1), 0, atol=1e-6) and np.allclose(fr.std(1), 1, atol=1e-4)
    assert np.allclose(fr[0, :512], fr[0, :512][0])          # the first frame starts in the zero padding (constant after centring)
    w = synthetic.make_crepe_params(0, 'tiny')
    act = oc.get_activation(x[:1600], w, 10.0)
    assert act.shape == (11, 360) and np.all((act > 0) & (act < 1))
    t, f0, conf, _ = oc.predict(x[:1600], w, 10.0)
    assert np.allclose(t, np.arange(11) * 0.01) and np.all(f0 > 30) and np.all(f0 < 2100)
    assert np.allclose(conf, act.max(1))


def test_product_decoder_tables_equal_the_oracles():
    """The host mirror (realtime_yukarin_b200/crepe.py) builds the HMM tables it uploads to the device exactly as the restatement does:
    the device's Viterbi sums are then bit-identical to the CPU's.  (The product module does not import the oracle; this test does.)"""
    from realtime_yukarin_b200 import crepe as pc
    ls, lt, (es, eo) = pc.pitch_hmm_tables()
    ols, olt, oemit = oc.pitch_hmm_tables()
    assert ls == ols[0] and np.all(ols == ols[0])
    assert np.array_equal(lt, olt)
    assert (es, eo) == (oemit[0], oemit[1])
    assert np.array_equal(np.linspace(0, 7180, 360) + 1997.3794084376191, oc.CENTS_MAPPING)
    # banded: transitions beyond +-11 bins are impossible
    assert np.isneginf(lt[0, 12]) and np.isfinite(lt[0, 11]) and np.isfinite(lt[200, 189]) and np.isneginf(lt[200, 188])


def test_product_extract_f0_applies_the_references_voicing_rule(monkeypatch):
    """realtime_yukarin_b200.crepe.extract_f0 = acoustic_feature_wrapper.py:66-80: frames are kept where the voicing HMM says voiced OR
    the confidence exceeds 0.1, everything else is zeroed (host logic only: the device call is replaced by canned outputs)."""
    from realtime_yukarin_b200 import crepe as pc
    f0 = np.array([100.0, 110.0, 120.0, 130.0])
    conf = np.array([0.05, 0.05, 0.2, 0.9], np.float32)
    voicing = np.array([0, 1, 0, 1], np.int32)

    def fake_predict(audio, sr, step_size=10.0, engine=None, details=False):
        t = np.arange(4) * step_size / 1000.0
        return (t, f0.copy(), conf, np.zeros((4, 360), np.float32), voicing, np.zeros(4, np.int32))

    monkeypatch.setattr(pc, 'predict', fake_predict)
    out, t = pc.extract_f0(np.zeros(100, np.float32), 24000, 5.0)
    assert np.array_equal(out, [0.0, 110.0, 120.0, 130.0])
    assert np.allclose(t, np.arange(4) * 0.005)


def test_network_forward_equals_a_plain_numpy_writing():
    """The torch-CPU forward of the restatement against a numpy-only one (sliding windows + einsum, explicit 'same' padding, pooling by
    reshape): pins the padding / stride / flatten conventions independently of torch's conv1d."""
    rng = np.random.default_rng(5)
    w = synthetic.make_crepe_params(1, 'tiny')
    x = rng.standard_normal(1200).astype(np.float32)
    frames = oc.frames_of(x, 10.0)[:3].astype(np.float64)
    h = frames[:, None, :]                                           # (frames, channels, time)
    for l in range(6):
        W = w[f'conv{l + 1}.W'].astype(np.float64)
        k, s = oc.WIDTHS[l], oc.STRIDES[l]
        n_out, left, right = oc.same_padding(h.shape[2], k, s)
        hp = np.pad(h, ((0, 0), (0, 0), (left, right)))
        win = np.lib.stride_tricks.sliding_window_view(hp, k, axis=2)[:, :, ::s][:, :, :n_out]      # (F, Cin, n_out, k)
        h = np.einsum('fcok,nck->fno', win, W) + w[f'conv{l + 1}.b'][None, :, None]
        h = np.maximum(h, 0.0)
        a = w[f'bn{l + 1}.gamma'].astype(np.float64) / np.sqrt(w[f'bn{l + 1}.var'].astype(np.float64) + oc.BN_EPS)
        h = h * a[None, :, None] + (w[f'bn{l + 1}.beta'] - w[f'bn{l + 1}.mean'] * a)[None, :, None]
        h = h.reshape(h.shape[0], h.shape[1], h.shape[2] // 2, 2).max(axis=3)
    flat = h.transpose(0, 2, 1).reshape(h.shape[0], -1)
    z = flat @ w['dense.W'].astype(np.float64).T + w['dense.b']
    ref = 1.0 / (1.0 + np.exp(-z))
    got = oc.get_activation(x, w, 10.0)[:3]
    assert np.abs(got - ref).max() < 2e-5
