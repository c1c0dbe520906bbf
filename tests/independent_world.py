"""Independent second writings (vectorised numpy, numpy.fft) of WORLD's DIO, StoneMask, CheapTrick and D4C from the published
algorithm (M. Morise, dio.cpp / stonemask.cpp / cheaptrick.cpp / d4c.cpp / common.cpp, v0.2.x) -- TEST INFRASTRUCTURE.  They exist only to cross-check
oracle/world_oracle.c: two writings made independently of each other must agree to rounding (tests/test_oracle.py).
The DECIDE points are the oracle's (no randn dither, `+ eps` instead of `+ |randn| * eps`).
History: DIO, CheapTrick and D4C agreed on the first comparison; StoneMask's second writing first used the wrong harmonic weighting
(sum a_k f_k / k / sum a_k instead of sum a_k f_k / sum a_k k) -- the C oracle had WORLD's form."""
import math

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



# ------------------------------------------------------------------------------------ DIO + StoneMask
def interp1(x, y, xi):
    x = np.asarray(x); y = np.asarray(y)
    k = np.clip(np.searchsorted(x, xi, side='right'), 1, len(x) - 1)
    s = (xi - x[k - 1]) / (x[k] - x[k - 1])
    return y[k - 1] + s * (y[k] - y[k - 1])

def nuttall(n):
    t = np.arange(n) / (n - 1.0)
    return 0.355768 - 0.487396 * np.cos(2 * np.pi * t) + 0.144232 * np.cos(4 * np.pi * t) - 0.012604 * np.cos(6 * np.pi * t)

def zc_engine(x, fs):
    n = len(x)
    neg = np.nonzero((x[:-1] > 0.0) & (x[1:] <= 0.0))[0] + 1
    if len(neg) < 2:
        return np.zeros(0), np.zeros(0)
    fine = neg - x[neg - 1] / (x[neg] - x[neg - 1])
    return (fine[:-1] + fine[1:]) / 2.0 / fs, fs / (fine[1:] - fine[:-1])

def dio_np(x, fs, frame_period=5.0, f0_floor=71.0, f0_ceil=800.0, channels=2.0, allowed_range=0.1):
    n = len(x)
    nbands = 1 + int(np.log(f0_ceil / f0_floor) / np.log(2.0) * channels)
    bounds = f0_floor * 2.0 ** ((np.arange(nbands) + 1) / channels)
    y_length = 1 + n
    fft = 2 ** (int(np.log2(y_length + mround(fs / 50.0) * 2 + 1 + 4 * int(1.0 + fs / bounds[0] / 2.0))) + 1)
    y = np.zeros(fft); y[:n] = x
    y[:y_length] -= y[:y_length].sum() / y_length
    Y = np.fft.fft(y)
    N = mround(fs / 50.0) * 2 + 1
    lc = np.zeros(fft)
    lc[:N] = 0.5 - 0.5 * np.cos(np.arange(1, N + 1) * 2 * np.pi / (N + 1))
    lc[:N] = -lc[:N] / lc[:N].sum()
    h = (N - 1) // 2
    f2 = np.zeros(fft)
    f2[fft - h:] = lc[:h]
    f2[:N - h] = lc[h:N]
    f2[0] += 1.0
    Y = Y * np.fft.fft(f2)
    f0_length = int(1000.0 * n / fs / frame_period) + 1
    t = np.arange(f0_length) * frame_period / 1000.0
    cands = np.zeros((nbands, f0_length)); scores = np.zeros((nbands, f0_length))
    for b in range(nbands):
        hl = mround(fs / bounds[b] / 2.0)
        lp = np.zeros(fft); lp[:hl * 4] = nuttall(hl * 4)
        filt = np.fft.ifft(Y * np.fft.fft(lp)).real * fft          # WORLD's c2r inverse is unnormalised (scale cancels in zero crossings)
        sig = filt[hl * 2:hl * 2 + y_length].copy()
        ev = []
        ev.append(zc_engine(sig, fs)); ev.append(zc_engine(-sig, fs))
        d = sig[:-1] - sig[1:]
        ev.append(zc_engine(d, fs)); ev.append(zc_engine(-d, fs))
        if any(len(e[0]) - 2 <= 0 for e in ev):
            scores[b] = 100000.0; continue
        sets = np.stack([interp1(loc, itv, t) for loc, itv in ev])
        c = sets.mean(0)
        s = np.sqrt(((sets - c) ** 2).sum(0) / 3.0)
        bad = (c > bounds[b]) | (c < bounds[b] / 2.0) | (c > f0_ceil) | (c < f0_floor)
        c[bad] = 0.0; s[bad] = 100000.0
        cands[b] = c; scores[b] = s / (c + 1e-12)
    best = cands[np.argmin(scores, axis=0), np.arange(f0_length)]
    # FixF0Contour
    vr = int(0.5 + 1000.0 / frame_period / f0_floor) * 2 + 1
    if f0_length <= vr:
        return np.zeros(f0_length), t
    base = np.zeros(f0_length); base[vr:f0_length - vr] = best[vr:f0_length - vr]
    s1 = np.zeros(f0_length)
    for i in range(vr, f0_length):
        s1[i] = base[i] if abs((base[i] - base[i - 1]) / (1e-12 + base[i])) < allowed_range else 0.0
    s2 = s1.copy()
    c = (vr - 1) // 2
    for i in range(c, f0_length - c):
        if np.any(s1[i - c:i + c + 1] == 0):
            s2[i] = 0.0
    pos = [i for i in range(1, f0_length) if s2[i - 1] == 0 and s2[i] != 0]
    neg = [i - 1 for i in range(1, f0_length) if s2[i] == 0 and s2[i - 1] != 0]
    def select(cur, past, idx):
        ref = (cur * 3.0 - past) / 2.0
        col = cands[:, idx]
        bf = col[np.argmin(np.abs(ref - col))]
        return 0.0 if abs(1.0 - bf / ref) > allowed_range else bf
    s3 = s2.copy()
    for i, st in enumerate(neg):
        limit = f0_length - 1 if i == len(neg) - 1 else neg[i + 1]
        for j in range(st, limit):
            s3[j + 1] = select(s3[j], s3[j - 1], j + 1)
            if s3[j + 1] == 0: break
    s4 = s3.copy()
    for i in range(len(pos) - 1, -1, -1):
        limit = 1 if i == 0 else pos[i - 1]
        for j in range(pos[i], limit, -1):
            s4[j - 1] = select(s4[j], s4[j + 1], j - 1)
            if s4[j - 1] == 0: break
    return s4, t

def stonemask_np(x, fs, t, f0):
    n = len(x); out = np.zeros(len(f0))
    for i, (tt, f) in enumerate(zip(t, f0)):
        if f <= 40.0 or f > fs / 12.0: continue
        half = int(1.5 * fs / f + 1.0)
        wl = (2 * half + 1) / fs
        bt = -half / fs + np.arange(2 * half + 1) / fs
        fft = 2 ** (2 + int(np.log(half * 2 + 1.0) / np.log(2.0)))
        idx = mround((tt + bt[0]) * fs + 0.001) + np.arange(len(bt))
        tm = (idx - 1.0) / fs - tt
        mw = 0.42 + 0.5 * np.cos(2 * np.pi * tm / wl) + 0.08 * np.cos(4 * np.pi * tm / wl)
        dw = np.empty_like(mw)
        dw[0] = -mw[1] / 2.0; dw[1:-1] = -(mw[2:] - mw[:-2]) / 2.0; dw[-1] = mw[-2] / 2.0
        safe = np.clip(idx - 1, 0, n - 1)
        def spec(w):
            b = np.zeros(fft); b[:len(w)] = x[safe] * w
            return np.fft.rfft(b)
        M, D = spec(mw), spec(dw)
        P = np.abs(M) ** 2
        NI = M.real * D.imag - M.imag * D.real
        def fix(f_init, nh):
            k = np.array([mround(f_init * fft / fs * (j + 1)) for j in range(nh)])
            inst = np.where(P[k] == 0, 0.0, k * fs / fft + NI[k] / np.where(P[k] == 0, 1, P[k]) * fs / 2.0 / np.pi)
            amp = np.sqrt(P[k])
            return (amp * inst).sum() / ((amp * (np.arange(nh) + 1.0)).sum() + 1e-12)      # stonemask.cpp FixF0: harmonic k votes f_k / k, weighted by amplitude * k
        nh = min(int(fs / 2.0 / f), 6)
        tent = fix(f, 2)
        mean = 0.0 if (tent <= 0 or tent > f * 2) else fix(tent, nh)
        out[i] = f if abs(mean - f) > f * 0.2 else mean
    return out



# ------------------------------------------------------------------------------------ realtime synthesizer
class NumpyRealtimeSynth:
    """Second writing (numpy) of the realtime synthesizer contract documented in DESIGN.md DECIDE 9-11 / SURVEY A.8."""

    def __init__(self, fs, frame_period_ms, fft, block):
        self.fs, self.fp, self.n, self.B = fs, frame_period_ms / 1000.0, fft, block
        self.f0 = np.zeros(0); self.sp = np.zeros((0, fft // 2 + 1), np.float32); self.ap = np.zeros((0, fft // 2 + 1), np.float32)
        self.handoff = False; self.h_phase = 0.0; self.h_f0 = 0.0
        self.pulses = []            # (index, time, vuv)
        self.next_pulse = 0; self.done = 0
        self.buf = np.zeros(2 * block + fft)
        i = np.arange(fft // 2)
        d = 0.5 - 0.5 * np.cos(2 * np.pi * (i + 1.0) / (1.0 + fft // 2))
        self.dcr = d / d.sum()

    def add(self, f0, sp, ap):
        n = len(f0)
        before = len(self.f0) - 1                       # cumulative_frame before
        self.f0 = np.concatenate([self.f0, f0]); self.sp = np.concatenate([self.sp, sp]); self.ap = np.concatenate([self.ap, ap])
        cum = before + n
        if cum < 1:
            self.h_f0, self.handoff = f0[-1], True
            return
        first = cum - n
        start = max(0, math.ceil(first * self.fp * self.fs)); end = math.ceil(cum * self.fp * self.fs)
        ns = end - start
        hf = 1 if self.handoff else 0
        cum0 = max(first, 0)
        ct = np.array([cum0 * self.fp] * hf + [(i + cum0 + hf) * self.fp for i in range(n)])
        cf = np.array([self.h_f0] * hf + list(f0))
        cv = (cf != 0).astype(float)
        ta = (np.arange(ns) + start) / self.fs
        k = np.clip(np.searchsorted(ct, ta, side='right'), 1, len(ct) - 1)
        s = (ta - ct[k - 1]) / (ct[k] - ct[k - 1])
        vuv = ((cv[k - 1] + s * (cv[k] - cv[k - 1])) > 0.5).astype(float)
        if0 = np.where(vuv == 0, 500.0, cf[k - 1] + s * (cf[k] - cf[k - 1]))
        np_ = ns + hf
        inc = np.zeros(np_)
        inc[1:] = 2 * np.pi * if0[np.arange(1, np_) - hf] / self.fs
        tp = np.empty(np_)
        base = self.h_phase if hf else 2 * np.pi * if0[0] / self.fs
        for b0 in range(0, np_, 256):
            loc = np.cumsum(inc[b0:b0 + 256])
            tp[b0:b0 + 256] = base + loc
            base = base + loc[-1]
        self.h_phase = tp[-1]
        wp = np.fmod(tp, 2 * np.pi)
        for i in np.nonzero(np.abs(np.diff(wp)) > np.pi)[0]:
            t = ta[i] - hf / self.fs
            idx = int(t * self.fs + 0.5) if t * self.fs > 0 else int(t * self.fs - 0.5)
            li = min(max(idx - start, 0), ns - 1)
            self.pulses.append((idx, t, int(vuv[li] > 0.5)))
        self.h_f0, self.handoff = f0[-1], True

    def _min_phase(self, log_half):
        n, half = self.n, self.n // 2
        full = np.concatenate([log_half, log_half[-2:0:-1]])
        cep = np.fft.fft(full).real
        cep[1:half] *= 2.0; cep[half + 1:] = 0.0
        spec = np.fft.fft(cep)[:half + 1] / n
        return np.exp(spec.real) * np.exp(1j * spec.imag)

    def _response(self, p, noise_size):
        idx, t, vuv = self.pulses[p]
        n, half = self.n, self.n // 2
        cumf = len(self.f0) - 1
        fl = min(int(t / self.fp), cumf); ce = min(math.ceil(t / self.fp), cumf)
        w = t / self.fp - int(t / self.fp)
        clip = lambda a: np.clip(a.astype(float), 0.001, 0.999999999999)
        if fl == ce:
            s, a = np.abs(self.sp[fl].astype(float)), clip(self.ap[fl]) ** 2
        else:
            s = (1 - w) * np.abs(self.sp[fl].astype(float)) + w * np.abs(self.sp[ce].astype(float))
            a = ((1 - w) * clip(self.ap[fl]) + w * clip(self.ap[ce])) ** 2
        if vuv == 0 or a[0] > 0.999:
            per = np.zeros(n)
        else:
            per = np.fft.fftshift(np.fft.irfft(self._min_phase(np.log(s * (1 - a) + 1e-12) / 2), n) * n)
            dc = per[half:].sum()
            per = np.concatenate([np.zeros(half), per[half:] - dc * self.dcr])
        r = W.randn_stream(max(idx, 0), noise_size)
        nz = np.zeros(n); nz[:noise_size] = r - r.mean()
        with np.errstate(divide='ignore', invalid='ignore'):
            m = self._min_phase(np.log(s * a) / 2 if vuv else np.log(s) / 2)
            aper = np.fft.fftshift(np.fft.irfft(m * np.fft.rfft(nz), n) * n)
        return (per * np.sqrt(noise_size) + aper) / n

    def synthesis2(self):
        if not self.pulses or self.done + self.B >= self.pulses[-1][0]:
            return None
        B, n = self.B, self.n
        self.buf = np.concatenate([self.buf[B:], np.zeros(B)])
        while self.next_pulse < len(self.pulses):
            cur = self.pulses[self.next_pulse][0]
            if cur >= self.done + B:
                break
            noise_size = min(max(self.pulses[self.next_pulse + 1][0] - cur, 1), n)
            resp = self._response(self.next_pulse, noise_size)
            off = cur - self.done - n // 2 + 1
            lo = max(0, -off)
            self.buf[off + lo:off + n] += resp[lo:]
            self.next_pulse += 1
        self.done += B
        return self.buf[:B].copy()

    def decode(self, f0, sp, ap):
        self.add(f0, sp, ap)
        out = []
        while True:
            y = self.synthesis2()
            if y is None:
                break
            out.append(y)
        return np.concatenate(out) if out else np.zeros(0)




# ------------------------------------------------------------------------------------ Harvest (second writing of its array stages)
def harvest_stages_np(x, fs, f0_floor=71.0, f0_ceil=800.0):
    """numpy / scipy writing of Harvest up to the refined, pruned candidates and of its smoothing stage (harvest.cpp: GetWaveformAndSpectrum,
    GetRawF0Candidates, DetectOfficialF0Candidates, OverlapF0Candidates, RefineF0Candidates, RemoveUnreliableCandidates, SmoothF0Contour).
    The sequential contour tracking (FixStep1..4) is not rewritten here.  Returns dict(y, raw, cand, score, smooth) where smooth(best)
    maps a tracked 1 ms contour to the smoothed one."""
    from scipy import signal
    x = np.asarray(x, np.float64)
    r = mround(fs / 8000.0)
    afs = fs / r
    lo, hi = f0_floor * 0.9, f0_ceil * 1.1
    channels = 1 + int(np.log(hi / lo) / np.log(2.0) * 40.0)
    boundary = lo * 2.0 ** ((np.arange(channels) + 1) / 40.0)
    ylen = int(np.ceil(len(x) / r))
    # decimate: pad with the edge values, reflect 9 samples, cheby1(3, 0.05 dB) forwards and backwards, every r-th sample
    if r == 1:
        y = x.copy()
    else:
        lag = int(np.ceil(140.0 / r) * r)
        nx = np.concatenate([np.full(lag, x[0]), x, np.full(lag, x[-1])])
        b, a = signal.cheby1(3, 0.05, 0.8 / r)
        t = np.concatenate([2 * nx[0] - nx[9:0:-1], nx, 2 * nx[-1] - nx[-2:-11:-1]])
        t = signal.lfilter(b, a, t)[::-1]
        t = signal.lfilter(b, a, t)[::-1]
        nout = (len(nx) - 1) // r + 1
        nbeg = r - r * nout + len(nx)
        dec = t[np.arange(nbeg, len(nx) + 9, r) + 8]
        y = dec[lag // r: lag // r + ylen]
    y = y - y.mean()
    fft = 2 ** (int(np.log2(ylen + 5 + 2 * int(2.0 * afs / boundary[0]))) + 1)
    Y = np.fft.rfft(y, fft)
    nf = int(1000.0 * len(x) / fs) + 1
    tpos = np.arange(nf) / 1000.0
    raw = np.zeros((channels, nf))
    for c, bf in enumerate(boundary):
        h = mround(afs / bf * 2.0)
        bp = nuttall(2 * h + 1) * np.cos(2 * np.pi * bf * np.arange(-h, h + 1) / afs)
        f = np.fft.irfft(Y * np.fft.rfft(bp, fft), fft)[h + 1: h + 1 + ylen]
        d = -f[:-1] + f[1:]                                     # (-f)[i] - (-f)[i + 1]
        trains = [zc_engine(f, afs), zc_engine(-f, afs), zc_engine(d, afs), zc_engine(-d, afs)]
        if any(len(t_[0]) <= 2 for t_ in trains):
            continue
        v = np.mean([interp1(loc, itv, tpos) for loc, itv in trains], axis=0)
        v[(v > bf * 1.1) | (v < bf * 0.9) | (v > f0_ceil) | (v < f0_floor)] = 0.0
        raw[c] = v
    # candidates: mean over runs of >= 10 agreeing channels
    max_cand = mround(channels / 10.0) * 7
    cand = np.zeros((nf, max_cand))
    ncand = 0
    for i in range(nf):
        vuv = (raw[:, i] > 0).astype(int); vuv[0] = vuv[-1] = 0
        dv = np.diff(vuv)
        st, ed = np.nonzero(dv == 1)[0] + 1, np.nonzero(dv == -1)[0] + 1
        k = 0
        for s_, e_ in zip(st, ed):
            if e_ - s_ >= 10:
                cand[i, k] = raw[s_:e_, i].sum() / (e_ - s_)
                k += 1
        ncand = max(ncand, k)
    base = cand[:, :ncand].copy()
    for i in range(1, 4):                                        # overlap +-3 frames
        cand[i:, ncand * i: ncand * (i + 1)] = base[:-i]
        cand[:-i, ncand * (i + 3): ncand * (i + 4)] = base[i:]
    nc = ncand * 7
    # refinement (GetRefinedF0) on the decimated signal
    score = np.zeros_like(cand)
    n = len(y)
    for i in range(nf):
        for j in range(nc):
            f = cand[i, j]
            if f <= 0:
                cand[i, j] = 0.0
                continue
            half = int(1.5 * afs / f + 1.0)
            wl = (2 * half + 1) / afs
            fsz = 2 ** (2 + int(np.log(half * 2 + 1.0) / np.log(2.0)))
            idx = mround((tpos[i] - half / afs) * afs + 0.001) + np.arange(2 * half + 1)
            tm = (idx - 1.0) / afs - tpos[i]
            mw = 0.42 + 0.5 * np.cos(2 * np.pi * tm / wl) + 0.08 * np.cos(4 * np.pi * tm / wl)
            dw = np.empty_like(mw)
            dw[0] = -mw[1] / 2.0; dw[1:-1] = -(mw[2:] - mw[:-2]) / 2.0; dw[-1] = mw[-2] / 2.0
            seg = y[np.clip(idx - 1, 0, n - 1)]
            M, D = np.fft.rfft(seg * mw, fsz), np.fft.rfft(seg * dw, fsz)
            nh = min(int(afs / 2.0 / f), 6)
            k = np.minimum(np.array([mround(f * fsz / afs * (q + 1)) for q in range(nh)]), fsz // 2)
            P = np.abs(M[k]) ** 2
            NI = M[k].real * D[k].imag - M[k].imag * D[k].real
            inst = np.where(P == 0, 0.0, k * afs / fsz + NI / np.where(P == 0, 1, P) * afs / 2.0 / np.pi)
            amp = np.sqrt(P)
            rf = (amp * inst).sum() / ((amp * (np.arange(nh) + 1.0)).sum() + 1e-12)
            sc = 1.0 / (np.abs((inst / (np.arange(nh) + 1.0) - f) / f).sum() / nh + 1e-12)
            if rf < f0_floor or rf > f0_ceil or sc < 2.5:
                rf, sc = 0.0, 0.0
            cand[i, j], score[i, j] = rf, sc
    # removal of candidates that neither neighbouring frame supports within 5 %
    keep = cand.copy()
    for i in range(1, nf - 1):
        for j in range(nc):
            ref = keep[i, j]
            if ref == 0:
                continue
            e1 = min(1.0, np.abs(ref - keep[i + 1, :nc]).min() / ref)
            e2 = min(1.0, np.abs(ref - keep[i - 1, :nc]).min() / ref)
            if min(e1, e2) > 0.05:
                cand[i, j] = 0.0; score[i, j] = 0.0

    def smooth(best):
        """SmoothF0Contour: per voiced section of the 300-frame padded contour, edge-extended, Butterworth(2) forwards and backwards."""
        b = np.array([0.0078202080334971724, 0.015640416066994345, 0.0078202080334971724])
        a = np.array([1.0, -1.7347257688092754, 0.76600660094326412])
        pad = np.concatenate([np.zeros(300), best, np.zeros(300)])
        v = (pad > 0).astype(int); v[0] = v[-1] = 0
        dv = np.diff(v)
        st, ed = np.nonzero(dv == 1)[0] + 1, np.nonzero(dv == -1)[0]
        out = np.zeros(len(best))
        for s_, e_ in zip(st, ed):
            sec = pad.copy(); sec[:s_] = pad[s_]; sec[e_ + 1:] = pad[e_]
            z = signal.lfilter(b, a, sec)[::-1]
            z = signal.lfilter(b, a, z)[::-1]
            out[s_ - 300: e_ + 1 - 300] = z[s_: e_ + 1]
        return out

    return dict(y=y, raw=raw, cand=cand, score=score, nc=nc, smooth=smooth)


def harvest_fix_contour_np(cand, score, nc):
    """Second writing of Harvest's FixF0Contour (SearchF0Base, FixStep1..4) on the pruned candidates [frames][columns]."""
    cand, score = cand[:, :nc], score[:, :nc]
    nf = len(cand)

    def boundaries(f0):
        v = (f0 > 0).astype(int); v[0] = v[-1] = 0
        ch = np.nonzero(np.diff(v) != 0)[0] + 1
        return [(int(ch[i]), int(ch[i + 1]) - 1) for i in range(0, len(ch) - 1, 2)]      # (first, last) frame of every voiced section

    def select_best(ref, row, allowed):
        err = np.abs(ref - row) / ref
        ok = np.nonzero(err <= allowed)[0]
        if len(ok) == 0:
            return 0.0
        m = err[ok].min()
        return float(row[ok[err[ok] == m][-1]])                   # the last candidate with the minimal error

    base = np.where(score.max(axis=1) > 0, cand[np.arange(nf), score.argmax(axis=1)], 0.0)
    # step 1: rapid changes
    s1 = np.zeros(nf)
    for i in range(2, nf):
        if base[i] == 0:
            continue
        ref = base[i - 1] * 2 - base[i - 2]
        with np.errstate(divide='ignore', invalid='ignore'):
            jump = abs((base[i] - ref) / ref) > 0.008 and abs(base[i] - base[i - 1]) / base[i - 1] > 0.008
        s1[i] = 0.0 if jump else base[i]
    # step 2: short sections
    s2 = s1.copy()
    for a, b in boundaries(s1):
        if b - a < 6:
            s2[a:b + 1] = 0.0
    # step 3: extension along the candidates, selection of long sections, merge
    secs = boundaries(s2)
    rows = []
    for a, b in secs:
        row = np.zeros(nf); row[a:b + 1] = s2[a:b + 1]
        rows.append(row)
    ext = []
    for (a, b), row in zip(secs, rows):
        def walk(origin, last, step):
            cur, where, miss = row[origin], origin, 0
            for i in range(abs(last - origin) + 1):
                idx = origin + step * (i + 1)
                row[idx] = select_best(cur, cand[idx], 0.18)
                if row[idx] == 0.0:
                    miss += 1
                else:
                    cur, miss, where = row[idx], 0, idx
                if miss == 4:
                    break
            return where
        e_ = walk(b, min(nf - 2, b + 100), 1)
        s_ = walk(a, max(1, a - 100), -1)
        ext.append((s_, e_))
    sel, mean_f0 = [], 0.0
    for (s_, e_), row in zip(ext, rows):
        mean_f0 = (mean_f0 + row[s_:e_].sum()) / (e_ - s_)          # the running value carries over (published source)
        if 2200.0 / mean_f0 < e_ - s_:
            sel.append((s_, e_, row))
    s3 = s2.copy()
    if sel:
        order = list(range(len(sel)))
        for i in range(1, len(sel)):
            for j in range(i - 1, -1, -1):
                if sel[order[j]][0] > sel[order[i]][0]:
                    order[i], order[j] = order[j], order[i]
                else:
                    break
        s3 = sel[0][2].copy()
        b0, b1 = sel[0][0], sel[0][1]
        bounds = [[s_, e_] for s_, e_, _ in sel]
        for q in range(1, len(sel)):
            o = order[q]
            st2, ed2 = (b0, b1) if o == 0 else bounds[o]
            row = sel[o][2]
            if st2 - b1 > 0:
                s3[st2:ed2 + 1] = row[st2:ed2 + 1]
                b0, b1 = st2, ed2
            elif not (b0 <= st2 and b1 >= ed2):
                def sc(f0v, i):
                    hit = cand[i] == f0v
                    return float(score[i][hit].max()) if hit.any() else 0.0
                sc1 = sum(sc(s3[i], i) for i in range(st2, b1 + 1))
                sc2 = sum(sc(row[i], i) for i in range(st2, b1 + 1))
                frm = b1 if sc1 > sc2 else st2
                s3[frm:ed2 + 1] = row[frm:ed2 + 1]
                b1 = ed2
    # step 4: short gaps
    s4 = s3.copy()
    secs = boundaries(s3)
    for (a0, b0_), (a1, _) in zip(secs[:-1], secs[1:]):
        dist = a1 - b0_ - 1
        if dist >= 9:
            continue
        t0, t1 = s3[b0_] + 1, s3[a1] - 1
        s4[b0_ + 1:a1] = t0 + (t1 - t0) / (dist + 1.0) * np.arange(1, dist + 1)
    return s4
