"""Pure-arithmetic WORLD constants shared by the host layer (no GPU, no library load)."""
import math


def cheaptrick_fft_size(sampling_rate: int, f0_floor: float = 71.0) -> int:
    """pyworld.get_cheaptrick_fft_size (call site: yukarin_wrapper/vocoder.py:83): 2^(1+floor(log2(3 fs / f0_floor + 1)))."""
    return 2 ** (1 + int(math.log(3.0 * sampling_rate / f0_floor + 1) / 0.69314718055994529))


def dio_num_frames(sampling_rate: int, n_samples: int, frame_period: float) -> int:
    return int(1000.0 * n_samples / sampling_rate / frame_period) + 1
