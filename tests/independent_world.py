"""Independent second writings (vectorised numpy, numpy.fft) of WORLD's CheapTrick and D4C from the published algorithm
(M. Morise, cheaptrick.cpp / d4c.cpp / common.cpp, v0.2.x) -- TEST INFRASTRUCTURE.  They exist only to cross-check
oracle/world_oracle.c: two writings made independently of each other must agree to rounding (tests/test_oracle.py).
The DECIDE points are the oracle's (no randn dither, `+ eps` instead of `+ |randn| * eps`)."""
import numpy as np

from oracle import world as W


def mround(v):
    return int(v + 0.5) if v > 0 else int(v - 0.5)

def interp1Q(x0, dx, y, xi):
    pos = (xi - x0) / dx
    base = pos.astype(int)
    frac = pos - base
    yy = np.append(y, y[-1])
    dy = yy[base + 1] - yy[base]
    dy[base + 1 >= len(y)] = 0.0
    return y[base] + dy * frac

def cheaptrick_np(x, fs, t, f0, fft=1024, q1=-0.15):
    n = len(x)
    f0_floor = 3.0 * fs / (fft - 3.0)
    nb = fft // 2 + 1
    out = np.empty((len(f0), nb))
    for fi, (tt, ff) in enumerate(zip(t, f0)):
        cf0 = 500.0 if ff <= f0_floor else ff
        half = mround(1.5 * fs / cf0)
        base = np.arange(-half, half + 1)
        idx = np.clip(mround(tt * fs + 0.001) + base, 0, n - 1)
        pos = base / 1.5 / fs
        win = 0.5 * np.cos(np.pi * pos * cf0) + 0.5
        win = win / np.sqrt(np.sum(win * win))
        w = x[idx] * win
        w = w - win * (w.sum() / win.sum())
        buf = np.zeros(fft); buf[:len(w)] = w
        ps = np.abs(np.fft.rfft(buf)) ** 2
        # DC correction
        upper = 2 + int(cf0 * fft / fs)
        lfa = np.arange(upper) * fs / fft
        rep = interp1Q(cf0 - lfa[0], -fs / fft, ps[:upper + 1], lfa[:upper - 1])
        ps = ps.copy()
        ps[:upper - 1] += rep
        # linear smoothing, width = f0 * 2 / 3
        width = cf0 * 2.0 / 3.0
        bnd = int(width * fft / fs) + 1
        mir = np.concatenate([ps[bnd:0:-1], ps, ps[-2:-bnd - 2:-1]])
        seg = np.cumsum(mir * fs / fft)
        fa = np.arange(nb) / fft * fs - width / 2.0
        origin = -(bnd - 0.5) * fs / fft
        low = interp1Q(origin, fs / fft, seg, fa)
        high = interp1Q(origin, fs / fft, seg, fa + width)
        sm = (high - low) / width
        sm = sm + 2.220446049250313e-16             # DECIDE 3: |randn| * eps -> eps
        # smoothing with recovery
        q = np.arange(1, nb) / fs
        sl = np.concatenate([[1.0], np.sin(np.pi * cf0 * q) / (np.pi * cf0 * q)])
        cl = np.concatenate([[1.0], (1 - 2 * q1) + 2 * q1 * np.cos(2 * np.pi * q * cf0)])
        logp = np.log(sm)
        full = np.concatenate([logp, logp[-2:0:-1]])
        cep = np.fft.fft(full).real
        cep_h = cep[:nb] * sl * cl / fft
        full2 = np.concatenate([cep_h, cep_h[-2:0:-1]])     # real even cepstrum -> real spectrum
        out[fi] = np.exp(np.fft.fft(full2).real[:nb])
    return out


def windowed(x, fs, f0, pos, blackman, ratio):
    n = len(x)
    half = mround(ratio * fs / f0 / 2.0)
    base = np.arange(-half, half + 1)
    idx = np.clip(mround(pos * fs + 0.001) + base, 0, n - 1)
    p = (2.0 * base / ratio) / fs
    win = 0.42 + 0.5 * np.cos(np.pi * p * f0) + 0.08 * np.cos(np.pi * p * f0 * 2) if blackman else 0.5 * np.cos(np.pi * p * f0) + 0.5
    w = x[idx] * win
    return w - win * (w.sum() / win.sum())

def dc_correction(sp, f0, fs, fft):
    upper = 2 + int(f0 * fft / fs)
    lfa = np.arange(upper) * fs / fft
    rep = interp1Q(f0 - lfa[0], -fs / fft, sp[:upper + 1], lfa[:upper - 1])
    out = sp.copy(); out[:upper - 1] += rep
    return out

def linear_smoothing(sp, width, fs, fft):
    nb = fft // 2 + 1
    bnd = int(width * fft / fs) + 1
    mir = np.concatenate([sp[bnd:0:-1], sp, sp[-2:-bnd - 2:-1]])
    seg = np.cumsum(mir * fs / fft)
    fa = np.arange(nb) / fft * fs - width / 2.0
    origin = -(bnd - 0.5) * fs / fft
    return (interp1Q(origin, fs / fft, seg, fa + width) - interp1Q(origin, fs / fft, seg, fa)) / width

def rfft_pad(w, fft):
    b = np.zeros(fft); b[:len(w)] = w
    return np.fft.rfft(b)

def d4c_np(x, fs, t, f0, fft_out=1024, threshold=0.85):
    nb_out = fft_out // 2 + 1
    fft = 2 ** (1 + int(np.log(4.0 * fs / 47.0 + 1) / np.log(2.0)))
    flt = 2 ** (1 + int(np.log(3.0 * fs / 40.0 + 1) / np.log(2.0)))
    nap = int(min(15000.0, fs / 2.0 - 3000.0) / 3000.0)
    wl = int(3000.0 * fft / fs) * 2 + 1
    tmp = np.arange(wl) / (wl - 1.0)
    nutt = 0.355768 - 0.487396 * np.cos(2 * np.pi * tmp) + 0.144232 * np.cos(4 * np.pi * tmp) - 0.012604 * np.cos(6 * np.pi * tmp)
    b0, b1, b2 = int(np.ceil(100.0 * flt / fs)), int(np.ceil(4000.0 * flt / fs)), int(np.ceil(7900.0 * flt / fs))
    out = np.empty((len(f0), nb_out))
    for i, (tt, ff) in enumerate(zip(t, f0)):
        unv = np.full(nb_out, 1 - 1e-12)
        if ff == 0:
            out[i] = unv; continue
        w = windowed(x, fs, max(ff, 40.0), tt, True, 3.0)
        ps = np.abs(rfft_pad(w, flt)) ** 2
        ps[:b0 + 1] = 0.0
        cs = np.cumsum(ps[:b2 + 1])
        if not (cs[b1] / cs[b2] > threshold):
            out[i] = unv; continue
        cf0 = max(47.0, ff)
        def centroid(pos):
            w = windowed(x, fs, cf0, pos, True, 4.0)
            lim = mround(2.0 * fs / cf0) * 2
            w = w.copy(); w[:lim + 1] = w[:lim + 1] / np.sqrt(np.sum(w[:lim + 1] ** 2))
            X = rfft_pad(w, fft)
            Y = rfft_pad(w * (np.arange(len(w)) + 1.0), fft)
            return Y.real * X.real + X.imag * Y.imag
        sc = dc_correction(centroid(tt - 0.25 / cf0) + centroid(tt + 0.25 / cf0), cf0, fs, fft)
        ps = np.abs(rfft_pad(windowed(x, fs, cf0, tt, False, 4.0), fft)) ** 2
        sps = linear_smoothing(dc_correction(ps, cf0, fs, fft), cf0, fs, fft)
        gd = linear_smoothing(sc / sps, cf0 / 2.0, fs, fft)
        gd = gd - linear_smoothing(gd, cf0, fs, fft)
        bnd = mround(fft * 8.0 / wl)
        hw = wl // 2
        coarse = np.empty(nap + 2); coarse[0] = -60.0; coarse[-1] = -1e-12
        for b in range(nap):
            c = int(3000.0 * (b + 1) * fft / fs)
            P = np.sort(np.abs(rfft_pad(gd[c - hw:c - hw + hw * 2 + 1] * nutt, fft)) ** 2)
            cs = np.cumsum(P)
            coarse[1 + b] = min(0.0, 10 * np.log10(cs[fft // 2 - bnd - 1] / cs[fft // 2]) + (cf0 - 100.0) / 50.0)
        axis = np.append(np.arange(nap + 1) * 3000.0, fs / 2.0)
        fx = np.arange(nb_out) * fs / fft_out
        out[i] = 10 ** (W.interp1(axis, coarse, fx) / 20.0)
    return out

