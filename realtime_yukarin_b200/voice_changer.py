"""VoiceChanger: stage-1 + stage-2 orchestration (voice_changer.py:8-42 of the reference)."""
import numpy

from .feature import AcousticFeature, AcousticFeatureWrapper
from .models import AcousticConverter, SuperResolution


class VoiceChanger(object):
    def __init__(self, acoustic_converter: AcousticConverter, super_resolution: SuperResolution, threshold: float = 60,
                 output_sampling_rate: int = None, fused: bool = False) -> None:
        if output_sampling_rate is None:
            output_sampling_rate = super_resolution.config.dataset.param.voice_param.sample_rate
        self.acoustic_converter = acoustic_converter
        self.super_resolution = super_resolution
        self.threshold = threshold
        self.output_sampling_rate = output_sampling_rate
        self.fused = fused      # True: one upload / one download through ryk_convert_window

    def convert_from_acoustic_feature(self, f_in: AcousticFeatureWrapper) -> AcousticFeature:
        if self.fused:
            return self._convert_fused(f_in)
        ac = self.acoustic_converter
        f_eff, effective = ac.separate_effective(wave=f_in.wave, feature=f_in, threshold=self.threshold)
        f_out = ac.convert(f_eff) if numpy.any(effective) else f_eff
        f_out = ac.combine_silent(effective=effective, feature=f_out)
        f_out = ac.decode_spectrogram(f_out)
        f_out.sp += 1e-16
        f_out.sp = self.super_resolution.convert(f_out.sp.astype(numpy.float32))
        return f_out

    def _convert_fused(self, f_in: AcousticFeatureWrapper) -> AcousticFeature:
        ac = self.acoustic_converter
        p = ac.config.dataset.acoustic_param
        from .world_consts import cheaptrick_fft_size
        out = ac.engine.convert_window(
            f_in.wave.wave, fs=p.sampling_rate, frame_length=p.fft_length, hop=p.sampling_rate * p.frame_period // 1000,
            threshold_db=self.threshold, f0=f_in.f0, ap=f_in.ap, mc=f_in.mc, voiced=f_in.voiced, order=p.order,
            alpha=p.alpha, fftlen=cheaptrick_fft_size(ac.out_sampling_rate))
        return AcousticFeature(f0=out['f0'][:, None], ap=out['ap'], sp=out['sp'], mc=out['mc'], voiced=out['voiced'][:, None])
