"""CPU restatement of the two Chainer U-Nets -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Follows the upstream model definitions the reference reaches through
`acoustic_converter.convert` (realtime_voice_conversion/yukarin_wrapper/voice_changer.py:33) and
`super_resolution.convert` (voice_changer.py:41): pix2pix-style encoder/decoder with
Convolution / Deconvolution (cross-correlation, W layouts (Cout,Cin,k..) / (Cin,Cout,k..)),
BatchNormalization in test mode (eps 2e-5, avg_mean / avg_var), LeakyReLU(0.2) / ReLU, skip concat
on the channel axis.  PARITY UNPINNED: `yukarin` / `become_yukarin` are not vendored; topology per
SURVEY App. A.6 / A.7.  Two interchangeable back-ends compute the same float32 math:
  'numpy' : im2col + matmul, the way Chainer's CPU path does (small shapes / tests)
  'torch' : torch.nn.functional conv on CPU threads (fast; used for the timed CPU baseline)
"""
from typing import Dict, List

import numpy as np

BN_EPS = 2e-5


def load_npz(path) -> Dict[str, np.ndarray]:
    with np.load(str(path), allow_pickle=False) as z:
        d = {k: z[k] for k in z.files}
    if d and all(k.startswith('predictor/') for k in d):
        d = {k[len('predictor/'):]: v for k, v in d.items()}
    return d


# ---------------------------------------------------------------- numpy back-end (N-d via 2-d)
def _as2d(x, W, ndim):
    if ndim == 1:
        return x[:, None, :], W[:, :, None, :]
    return x, W


def _conv2d_np(x, W, stride, pad):
    """x (Cin,H,Wd), W (Cout,Cin,kh,kw); stride/pad tuples. float32 cross-correlation."""
    Cin, H, Wd = x.shape
    Cout, _, kh, kw = W.shape
    sh, sw = stride
    ph, pw = pad
    xp = np.pad(x, ((0, 0), (ph, ph), (pw, pw)))
    Ho = (H + 2 * ph - kh) // sh + 1
    Wo = (Wd + 2 * pw - kw) // sw + 1
    s = xp.strides
    cols = np.lib.stride_tricks.as_strided(xp, shape=(Cin, kh, kw, Ho, Wo), strides=(s[0], s[1], s[2], s[1] * sh, s[2] * sw))
    cols = np.ascontiguousarray(cols).reshape(Cin * kh * kw, Ho * Wo)
    out = W.reshape(Cout, -1).astype(np.float32) @ cols.astype(np.float32)
    return out.reshape(Cout, Ho, Wo)


def _deconv2d_np(x, W, stride, pad):
    """Chainer Deconvolution: W (Cin,Cout,kh,kw); out = (in-1)*s + k - 2p."""
    Cin, H, Wd = x.shape
    _, Cout, kh, kw = W.shape
    sh, sw = stride
    ph, pw = pad
    up = np.zeros((Cin, (H - 1) * sh + 1, (Wd - 1) * sw + 1), dtype=x.dtype)
    up[:, ::sh, ::sw] = x
    Wf = np.ascontiguousarray(W[:, :, ::-1, ::-1].transpose(1, 0, 2, 3))
    return _conv2d_np(up, Wf, (1, 1), (kh - 1 - ph, kw - 1 - pw))


def _conv_np(x, W, b, stride, pad, transposed, ndim):
    x2, W2 = _as2d(x, W, ndim)
    st = (stride, stride) if ndim == 2 else (1, stride)
    pd = (pad, pad) if ndim == 2 else (0, pad)
    y = _deconv2d_np(x2, W2, st, pd) if transposed else _conv2d_np(x2, W2, st, pd)
    if b is not None:
        y = y + b.astype(np.float32)[:, None, None]
    return y[:, 0, :] if ndim == 1 else y


# ---------------------------------------------------------------- torch back-end
def _conv_torch(x, W, b, stride, pad, transposed, ndim):
    import torch
    import torch.nn.functional as F
    xt = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))[None]
    Wt = torch.from_numpy(np.ascontiguousarray(W, dtype=np.float32))
    bt = None if b is None else torch.from_numpy(np.ascontiguousarray(b, dtype=np.float32))
    if ndim == 1:
        y = F.conv_transpose1d(xt, Wt, bt, stride=stride, padding=pad) if transposed else F.conv1d(xt, Wt, bt, stride=stride, padding=pad)
    else:
        y = F.conv_transpose2d(xt, Wt, bt, stride=stride, padding=pad) if transposed else F.conv2d(xt, Wt, bt, stride=stride, padding=pad)
    return y[0].numpy()


def _bn(x, p, prefix):
    shape = (-1,) + (1,) * (x.ndim - 1)
    g, be = p[prefix + '/gamma'].reshape(shape), p[prefix + '/beta'].reshape(shape)
    m, v = p[prefix + '/avg_mean'].reshape(shape), p[prefix + '/avg_var'].reshape(shape)
    return ((x - m) / np.sqrt(v + np.float32(BN_EPS)) * g + be).astype(np.float32)


def unet_forward(x: np.ndarray, p: Dict[str, np.ndarray], ndim: int, backend: str = 'numpy', return_all: bool = False):
    """x: (Cin, T) for ndim 1 or (Cin, H, W) for ndim 2, float32. Returns the decoder c7 output."""
    conv = _conv_np if backend == 'numpy' else _conv_torch
    x = np.asarray(x, dtype=np.float32)
    hs: List[np.ndarray] = []
    h = conv(x, p['encoder/c0/W'], p.get('encoder/c0/b'), 1, 1, False, ndim)
    h = np.where(h > 0, h, np.float32(0.2) * h).astype(np.float32)
    hs.append(h)
    for i in range(1, 8):
        h = conv(hs[i - 1], p[f'encoder/c{i}/c/W'], p.get(f'encoder/c{i}/c/b'), 2, 1, False, ndim)
        if f'encoder/c{i}/batchnorm/gamma' in p:
            h = _bn(h, p, f'encoder/c{i}/batchnorm')
        h = np.where(h > 0, h, np.float32(0.2) * h).astype(np.float32)
        hs.append(h)
    dec = []
    h = hs[-1]
    for i in range(8):
        if i > 0:
            h = np.concatenate([h, hs[-i - 1]], axis=0)
        if i < 7:
            h = conv(h, p[f'decoder/c{i}/c/W'], p.get(f'decoder/c{i}/c/b'), 2, 1, True, ndim)
            if f'decoder/c{i}/batchnorm/gamma' in p:
                h = _bn(h, p, f'decoder/c{i}/batchnorm')
            h = np.maximum(h, np.float32(0)).astype(np.float32)      # dropout is the identity at test time
        else:
            h = conv(h, p['decoder/c7/W'], p.get('decoder/c7/b'), 1, 1, False, ndim)
        dec.append(h)
    if return_all:
        return h, hs, dec
    return h


# ---------------------------------------------------------------- the two convert() wrappers
def stage1_convert(mc: np.ndarray, p: Dict[str, np.ndarray], backend: str = 'numpy') -> np.ndarray:
    """yukarin AcousticConverter.convert, feature part: mc (T, C) float32 -> (T, C) float32."""
    mc = np.asarray(mc, dtype=np.float32)
    C = mc.shape[1]
    in_mean, in_std = p.get('stats/in_mean', np.zeros(C, np.float32)), p.get('stats/in_std', np.ones(C, np.float32))
    out_mean, out_std = p.get('stats/out_mean', np.zeros(C, np.float32)), p.get('stats/out_std', np.ones(C, np.float32))
    x = ((mc - in_mean) / in_std).astype(np.float32).T            # (C, T)
    T = x.shape[1]
    pad = 128 - T % 128
    x = np.pad(x, [(0, 0), (0, pad)], mode='minimum')
    y = unet_forward(x, p, 1, backend)[:, :-pad]
    return (y.T * out_std + out_mean).astype(np.float32)


def stage2_convert(sp: np.ndarray, p: Dict[str, np.ndarray], backend: str = 'numpy') -> np.ndarray:
    """become_yukarin SuperResolution.convert: sp (T, 513) float32 -> (T, 513) float32."""
    x = np.asarray(sp, dtype=np.float32)
    pad = 128 - len(x) % 128
    x = np.pad(x, [(0, pad), (0, 0)], mode='minimum')
    x = np.log(x)[:, :-1][np.newaxis]
    y = unet_forward(x, p, 2, backend)[0]
    y = np.pad(y, [(0, 0), (0, 1)], mode='edge')
    y = np.exp(y)
    return y[:-pad].astype(np.float32)
