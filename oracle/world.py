"""ctypes front-end of oracle/world_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the WORLD / SPTK arithmetic reached from
realtime_voice_conversion/yukarin_wrapper/vocoder.py:26-48 (analysis), voice_changer.py:38 (mc2sp)
and vocoder.py:72-120 (realtime synthesis).  PARITY UNPINNED (see world_oracle.c header).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / '_build' / 'libworld_oracle.so'
_SRC = _HERE / 'world_oracle.c'

c_double_p = ctypes.POINTER(ctypes.c_double)
c_float_p = ctypes.POINTER(ctypes.c_float)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_ll_p = ctypes.POINTER(ctypes.c_longlong)


def build(force: bool = False) -> Path:
    """gcc -O2 the C restatement into oracle/_build/ (a few hundred ms)."""
    if force or not _SO.exists() or _SO.stat().st_mtime < _SRC.stat().st_mtime:
        _SO.parent.mkdir(exist_ok=True)
        subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-std=c99', '-o', str(_SO), str(_SRC), '-lm'])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(_SO))
        _lib.wo_synth_create.restype = ctypes.c_void_p
        _lib.wo_synth_create.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        _lib.wo_synth_add.argtypes = [ctypes.c_void_p, c_double_p, ctypes.c_int, c_float_p, c_float_p]
        _lib.wo_synth_synthesis2.argtypes = [ctypes.c_void_p, c_double_p]
        _lib.wo_synth_destroy.argtypes = [ctypes.c_void_p]
        _lib.wo_synth_set_mode.argtypes = [ctypes.c_void_p, ctypes.c_int]
        _lib.wo_synth_skip_randn.argtypes = [ctypes.c_void_p, ctypes.c_longlong]
        _lib.wo_synth_pulse_count.argtypes = [ctypes.c_void_p]
        _lib.wo_synth_pulse_count.restype = ctypes.c_longlong
        _lib.wo_synth_get_pulses.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, c_ll_p, c_double_p, c_int_p]
        _lib.wo_randn_stream.argtypes = [ctypes.c_longlong, ctypes.c_int, c_double_p]
        _lib.wo_cheaptrick_fft_size.argtypes = [ctypes.c_int, ctypes.c_double]
    return _lib


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def cheaptrick_fft_size(fs: int, f0_floor: float = 71.0) -> int:
    return lib().wo_cheaptrick_fft_size(int(fs), float(f0_floor))


def fft(x: np.ndarray, sign: int = -1) -> np.ndarray:
    re = _f64(x.real).copy()
    im = _f64(x.imag).copy() if np.iscomplexobj(x) else np.zeros_like(re)
    lib().wo_fft(_dp(re), _dp(im), ctypes.c_int(len(re)), ctypes.c_int(sign))
    return re + 1j * im


def interp1(x, y, xi):
    x, y, xi = _f64(x), _f64(y), _f64(xi)
    out = np.empty_like(xi)
    lib().wo_interp1(_dp(x), _dp(y), ctypes.c_int(len(x)), _dp(xi), ctypes.c_int(len(xi)), _dp(out))
    return out


def randn_stream(start: int, n: int) -> np.ndarray:
    out = np.empty(n)
    lib().wo_randn_stream(start, n, _dp(out))
    return out


def dio(x, fs, frame_period=5.0, f0_floor=71.0, f0_ceil=800.0, debug=False):
    x = _f64(x)
    n = lib().wo_dio_num_frames(int(fs), len(x), ctypes.c_double(frame_period))
    t = np.empty(n)
    f0 = np.empty(n)
    nb = 1 + int(np.log(f0_ceil / f0_floor) / 0.69314718055994529 * 2.0)
    cand = np.empty((nb, n))
    score = np.empty((nb, n))
    lib().wo_dio(_dp(x), ctypes.c_int(len(x)), ctypes.c_int(int(fs)), ctypes.c_double(frame_period),
                 ctypes.c_double(f0_floor), ctypes.c_double(f0_ceil), _dp(t), _dp(f0), _dp(cand), _dp(score))
    if debug:
        return f0, t, cand, score
    return f0, t


def harvest_geometry(n, fs, f0_floor=71.0, f0_ceil=800.0):
    """(channels, basic 1 ms frames, decimated length, fft size, candidate columns, decimation ratio) of HarvestGeneral."""
    info = (ctypes.c_int * 6)()
    lib().wo_harvest_geometry(ctypes.c_int(int(n)), ctypes.c_int(int(fs)), ctypes.c_double(f0_floor), ctypes.c_double(f0_ceil), info)
    return tuple(int(v) for v in info)


def harvest(x, fs, frame_period=5.0, f0_floor=71.0, f0_ceil=800.0, debug=False):
    """pyworld.harvest restated (world_oracle.c, Harvest section).  debug=True also returns the intermediate arrays
    dict(y, raw, cand, score, best, basic, nc) used to localise a GPU mismatch stage by stage."""
    x = _f64(x)
    n = lib().wo_harvest_num_frames(int(fs), len(x), ctypes.c_double(frame_period))
    t, f0 = np.empty(n), np.empty(n)
    ch, nf, ylen, fft_size, maxc, ratio = harvest_geometry(len(x), fs, f0_floor, f0_ceil)
    dbg = dict(y=np.zeros(ylen), raw=np.zeros((ch, nf)), cand=np.zeros((nf, maxc)), score=np.zeros((nf, maxc)), best=np.zeros(nf),
               basic=np.zeros(nf))
    nc = ctypes.c_int(0)
    lib().wo_harvest_ex(_dp(x), ctypes.c_int(len(x)), ctypes.c_int(int(fs)), ctypes.c_double(frame_period), ctypes.c_double(f0_floor),
                        ctypes.c_double(f0_ceil), _dp(t), _dp(f0), _dp(dbg['y']), _dp(dbg['raw']), _dp(dbg['cand']), _dp(dbg['score']),
                        _dp(dbg['best']), _dp(dbg['basic']), ctypes.byref(nc))
    if debug:
        dbg['nc'] = int(nc.value)
        return f0, t, dbg
    return f0, t


def decimate(x, r):
    """matlabfunctions.cpp decimate() (zero-phase Chebyshev I of order 3, every r-th sample)."""
    x = _f64(x)
    y = np.zeros((len(x) - 1) // r + 1 + 16)      # the C loop also emits ceil(9 / r) - 1 samples of the reflected tail
    lib().wo_decimate(_dp(x), ctypes.c_int(len(x)), ctypes.c_int(int(r)), _dp(y))
    return y[:(len(x) - 1) // r + 1]


def stonemask(x, fs, t, f0):
    x, t, f0 = _f64(x), _f64(t), _f64(f0)
    out = np.empty_like(f0)
    lib().wo_stonemask(_dp(x), ctypes.c_int(len(x)), ctypes.c_int(int(fs)), _dp(t), _dp(f0), ctypes.c_int(len(f0)), _dp(out))
    return out


def cheaptrick(x, fs, t, f0, fft_size=None, q1=-0.15):
    x, t, f0 = _f64(x), _f64(t), _f64(f0)
    if fft_size is None:
        fft_size = cheaptrick_fft_size(fs)
    sp = np.empty((len(f0), fft_size // 2 + 1))
    lib().wo_cheaptrick(_dp(x), ctypes.c_int(len(x)), ctypes.c_int(int(fs)), _dp(t), _dp(f0), ctypes.c_int(len(f0)),
                        ctypes.c_int(fft_size), ctypes.c_double(q1), _dp(sp))
    return sp


def d4c(x, fs, t, f0, fft_size=None, threshold=0.85, debug=False):
    x, t, f0 = _f64(x), _f64(t), _f64(f0)
    if fft_size is None:
        fft_size = cheaptrick_fft_size(fs)
    ap = np.empty((len(f0), fft_size // 2 + 1))
    ap0 = np.empty(len(f0))
    lib().wo_d4c(_dp(x), ctypes.c_int(len(x)), ctypes.c_int(int(fs)), _dp(t), _dp(f0), ctypes.c_int(len(f0)),
                 ctypes.c_int(fft_size), ctypes.c_double(threshold), _dp(ap), _dp(ap0))
    if debug:
        return ap, ap0
    return ap


def freqt(c, order, alpha):
    c = _f64(c)
    out = np.empty(order + 1)
    lib().wo_freqt(_dp(c), ctypes.c_int(len(c) - 1), _dp(out), ctypes.c_int(order), ctypes.c_double(alpha))
    return out


def sp2mc(sp, order, alpha):
    sp = _f64(sp)
    T, nb = sp.shape
    mc = np.empty((T, order + 1))
    lib().wo_sp2mc(_dp(sp), ctypes.c_int(T), ctypes.c_int((nb - 1) * 2), ctypes.c_int(order), ctypes.c_double(alpha), _dp(mc))
    return mc


def mc2sp(mc, alpha, fftlen):
    mc = _f64(mc)
    T, d = mc.shape
    sp = np.empty((T, fftlen // 2 + 1))
    lib().wo_mc2sp(_dp(mc), ctypes.c_int(T), ctypes.c_int(d - 1), ctypes.c_double(alpha), ctypes.c_int(fftlen), _dp(sp))
    return sp


def frame_mse(wave, frame_length, hop, n_frames=None):
    w = np.ascontiguousarray(wave, dtype=np.float32)
    if n_frames is None:
        n_frames = 1 + len(w) // hop
    out = np.empty(n_frames)
    lib().wo_frame_mse(w.ctypes.data_as(c_float_p), ctypes.c_int(len(w)), ctypes.c_int(frame_length), ctypes.c_int(hop),
                       _dp(out), ctypes.c_int(n_frames))
    return out


class RealtimeSynthesizer:
    """Restatement of world4py's WorldSynthesizer + _InitializeSynthesizer/_AddParameters/_Synthesis2
    (call sites: realtime_voice_conversion/yukarin_wrapper/vocoder.py:79-103)."""

    CANON_RANDN, CANON_PHASE = 1, 2

    def __init__(self, fs, frame_period, fft_size, buffer_size, ring_frames=4096, canonical: int = 0):
        """canonical: bit 0 = WORLD's sequential randn() consumption, bit 1 = WORLD's single running-sum phase (DECIDE 10 / 11 undone);
        0 (default) = the variants the CUDA path reproduces bit for bit.  Only tests/test_oracle_canonical.py uses non-zero values."""
        self.buffer_size = buffer_size
        self.fft_size = fft_size
        self._h = lib().wo_synth_create(int(fs), float(frame_period), int(fft_size), int(buffer_size), int(ring_frames))
        if canonical:
            lib().wo_synth_set_mode(self._h, int(canonical))

    def skip_randn(self, n: int):
        lib().wo_synth_skip_randn(self._h, int(n))

    def add_parameters(self, f0, sp, ap) -> int:
        f0 = _f64(np.asarray(f0).ravel())
        sp = np.ascontiguousarray(sp, dtype=np.float32)
        ap = np.ascontiguousarray(ap, dtype=np.float32)
        return lib().wo_synth_add(self._h, _dp(f0), ctypes.c_int(len(f0)), sp.ctypes.data_as(c_float_p), ap.ctypes.data_as(c_float_p))

    def synthesis2(self):
        out = np.empty(self.buffer_size)
        ok = lib().wo_synth_synthesis2(self._h, _dp(out))
        return out if ok else None

    def pulses(self):
        n = lib().wo_synth_pulse_count(self._h)
        idx = np.empty(n, dtype=np.int64)
        tm = np.empty(n)
        vuv = np.empty(n, dtype=np.int32)
        if n:
            lib().wo_synth_get_pulses(self._h, 0, int(n), idx.ctypes.data_as(c_ll_p), _dp(tm), vuv.ctypes.data_as(c_int_p))
        return idx, tm, vuv

    def decode(self, f0, sp, ap) -> np.ndarray:
        """RealtimeVocoder.decode (vocoder.py:89-120): add, then drain whole blocks."""
        self.add_parameters(f0, sp, ap)
        ys = []
        while True:
            y = self.synthesis2()
            if y is None:
                break
            ys.append(y)
        return np.concatenate(ys) if ys else np.empty(0)

    def __del__(self):
        try:
            lib().wo_synth_destroy(self._h)
        except Exception:
            pass


def synthesize(f0, sp, ap, fs, frame_period=5.0, fft_size=None, return_pulses=False):
    """pyworld.synthesize restated (WORLD synthesis.cpp Synthesis(); call site
    realtime_voice_conversion/yukarin_wrapper/vocoder.py:50-62).  sp / ap enter as float32."""
    f0 = _f64(np.asarray(f0).ravel())
    sp = np.ascontiguousarray(sp, dtype=np.float32)
    ap = np.ascontiguousarray(ap, dtype=np.float32)
    if fft_size is None:
        fft_size = (sp.shape[1] - 1) * 2
    L = lib()
    L.wo_synthesize_length.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_int]
    n = L.wo_synthesize_length(len(f0), float(frame_period), int(fs))
    y = np.zeros(max(n, 0))
    cap = max(n, 1)
    idx = np.zeros(cap, dtype=np.int64)
    shift = np.zeros(cap)
    vuv = np.zeros(cap, dtype=np.int32)
    L.wo_synthesize.argtypes = [c_double_p, ctypes.c_int, c_float_p, c_float_p, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                c_double_p, ctypes.c_int, c_ll_p, c_double_p, c_int_p]
    L.wo_synthesize.restype = ctypes.c_int
    npulse = L.wo_synthesize(_dp(f0), len(f0), sp.ctypes.data_as(c_float_p), ap.ctypes.data_as(c_float_p), int(fft_size), float(frame_period),
                             int(fs), n, _dp(y), cap, idx.ctypes.data_as(c_ll_p), _dp(shift), vuv.ctypes.data_as(c_int_p))
    if return_pulses:
        return y, idx[:npulse], shift[:npulse], vuv[:npulse]
    return y


def stft_power_db_mean(wave, n_fft=2048, hop=512, amin=1e-10, top_db=80.0) -> float:
    """librosa.core.power_to_db(numpy.abs(librosa.stft(wave)) ** 2).mean() restated
    (realtime_voice_conversion/worker/decode_worker.py:56), fp64."""
    w = _f64(np.asarray(wave).ravel())
    L = lib()
    L.wo_stft_power_db_mean.restype = ctypes.c_double
    L.wo_stft_power_db_mean.argtypes = [c_double_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double]
    return float(L.wo_stft_power_db_mean(_dp(w), len(w), int(n_fft), int(hop), float(amin), float(top_db)))
