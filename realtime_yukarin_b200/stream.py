"""Stream plugin API: a time-indexed segment store plus the three concrete stages of the hot path.

Re-creation of realtime_voice_conversion/stream/{base_stream,encode_stream,convert_stream,
decode_stream,stream_wrapper}.py.  The framing contract that must hold bit-for-bit
(base_stream.py:32-79):
  * a window [start - extra, start + length + extra) is assembled from the stored segments in
    insertion order, zero/silent-padded where no segment covers it;
  * segment-relative indices are round((t - segment.start) * rate) and
    round(first + span * rate) with Python's float arithmetic and banker's rounding;
  * segments that merely touch the window are visited (the overlap test uses strict <);
  * each stage trims round(extra * rate) items from both ends of its *output*
    (encode: output rate, encode_stream.py:38; convert: input rate, convert_stream.py:40).
The window assembly is expressed as a list of (kind, ...) pieces (`plan_window`) so that the same
arithmetic can drive host arrays here and device-resident rings in `session.py`.
"""
from abc import abstractmethod
from typing import Generic, List, Tuple, TypeVar

import numpy

from .feature import AcousticFeature, AcousticFeatureWrapper, Wave
from .segment import (BaseSegmentMethod, FeatureSegmentMethod, FeatureWrapperSegmentMethod, Segment,
                      WaveSegmentMethod)

T_IN = TypeVar('T_IN')
T_OUT = TypeVar('T_OUT')

PAD = 'pad'
PICK = 'pick'


def plan_window(segments, rate, start_time: float, time_length: float, extra_time: float) -> List[Tuple]:
    """Pieces that make up the requested window: (PAD, n_items) or (PICK, segment, first, last)."""
    t0 = start_time - extra_time
    span = time_length + extra_time * 2
    t1 = t0 + span

    pieces: List[Tuple] = []
    cursor = t0          # time up to which the window has been filled
    left = span          # time still to fill
    finished = False
    for seg in segments:
        seg_start, seg_end = seg.start_time, seg.end_time
        if t1 < seg_start or seg_end < t0:
            continue
        if seg_start > cursor:
            gap = seg_start - cursor
            pieces.append((PAD, round(gap * rate)))
            left -= gap
            cursor = seg_start
        available = seg_end - cursor
        take = available if left > available else left
        first = round((cursor - seg_start) * rate)
        last = round(first + take * rate)
        pieces.append((PICK, seg, first, last))
        cursor += take
        left -= take
        if cursor >= t1:
            finished = True
            break
    if not finished:
        pieces.append((PAD, round((t1 - cursor) * rate)))
    return pieces


class BaseStream(Generic[T_IN, T_OUT]):
    def __init__(self, in_segment_method: BaseSegmentMethod, out_segment_method: BaseSegmentMethod):
        self.in_segment_method = in_segment_method
        self.out_segment_method = out_segment_method
        self.stream: List[Segment] = []

    def add(self, start_time: float, data: T_IN):
        self.stream.append(Segment(start_time=start_time, data=data, method=self.in_segment_method))

    def remove(self, end_time: float):
        self.stream = [s for s in self.stream if s.end_time > end_time]

    def fetch(self, start_time: float, time_length: float, extra_time: float) -> T_IN:
        method = self.in_segment_method
        parts = []
        for piece in plan_window(self.stream, method.sampling_rate, start_time, time_length, extra_time):
            if piece[0] == PAD:
                parts.append(method.pad(piece[1]))
            else:
                _, seg, first, last = piece
                parts.append(method.pick(seg.data, first, last))
        return method.concat(parts)

    @abstractmethod
    def process(self, start_time: float, time_length: float, extra_time: float) -> T_OUT:
        raise NotImplementedError()


class EncodeStream(BaseStream[numpy.ndarray, AcousticFeatureWrapper]):
    """wave -> WORLD features (encode_stream.py:11-42)."""

    def __init__(self, vocoder):
        p = vocoder.acoustic_param
        super().__init__(
            in_segment_method=WaveSegmentMethod(sampling_rate=p.sampling_rate),
            out_segment_method=FeatureWrapperSegmentMethod(
                sampling_rate=1000 // p.frame_period, wave_sampling_rate=p.sampling_rate,
                order=p.order, frame_period=p.frame_period),
        )
        self.vocoder = vocoder

    def process(self, start_time: float, time_length: float, extra_time: float) -> AcousticFeatureWrapper:
        samples = self.fetch(start_time=start_time, time_length=time_length, extra_time=extra_time)
        feature = self.vocoder.encode(Wave(wave=samples, sampling_rate=self.in_segment_method.sampling_rate))
        trim = round(extra_time * self.out_segment_method.sampling_rate)
        if trim > 0:
            feature = self.out_segment_method.pick(feature, trim, -trim)
        return feature


class ConvertStream(BaseStream[AcousticFeatureWrapper, AcousticFeature]):
    """WORLD features -> converted features with super-resolved spectrum (convert_stream.py:10-44)."""

    def __init__(self, voice_changer):
        ac = voice_changer.acoustic_converter.config.dataset.acoustic_param
        sr = voice_changer.super_resolution.config.dataset.param.acoustic_feature_param
        super().__init__(
            in_segment_method=FeatureWrapperSegmentMethod(
                sampling_rate=1000 // ac.frame_period, wave_sampling_rate=ac.sampling_rate,
                order=ac.order, frame_period=ac.frame_period),
            out_segment_method=FeatureSegmentMethod(
                sampling_rate=1000 // sr.frame_period, wave_sampling_rate=voice_changer.output_sampling_rate,
                order=sr.order),
        )
        self.voice_changer = voice_changer

    def process(self, start_time: float, time_length: float, extra_time: float) -> AcousticFeature:
        window = self.fetch(start_time=start_time, time_length=time_length, extra_time=extra_time)
        converted = self.voice_changer.convert_from_acoustic_feature(window)
        trim = round(extra_time * self.in_segment_method.sampling_rate)
        if trim > 0:
            converted = self.out_segment_method.pick(converted, trim, -trim)
        return converted


class DecodeStream(BaseStream[AcousticFeature, numpy.ndarray]):
    """converted features -> waveform through the realtime synthesizer (decode_stream.py:10-39)."""

    def __init__(self, vocoder):
        p = vocoder.acoustic_param
        super().__init__(
            in_segment_method=FeatureSegmentMethod(
                sampling_rate=1000 // p.frame_period, wave_sampling_rate=vocoder.out_sampling_rate, order=p.order),
            out_segment_method=WaveSegmentMethod(sampling_rate=vocoder.out_sampling_rate),
        )
        self.vocoder = vocoder

    def process(self, start_time: float, time_length: float, extra_time: float) -> numpy.ndarray:
        feature = self.fetch(start_time=start_time, time_length=time_length, extra_time=extra_time)
        wave = self.vocoder.decode(acoustic_feature=feature).wave
        wave[numpy.isnan(wave)] = 0      # sp == 0 silent frames synthesise NaN (decode_stream.py:38)
        return wave


class StreamWrapper(object):
    """Sequential driver: each call processes the next `time_length` seconds (stream_wrapper.py:4-18)."""

    def __init__(self, stream: BaseStream, extra_time: float):
        self.stream = stream
        self.extra_time = extra_time
        self._current_time = 0.

    def process_next(self, time_length: float):
        out = self.stream.process(start_time=self._current_time, time_length=time_length,
                                  extra_time=self.extra_time)
        self._current_time += time_length
        return out
