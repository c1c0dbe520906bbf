"""SegmentMethod plugin API: how a stream's payload type is measured, padded, sliced and joined.

Re-creation of realtime_voice_conversion/segment/{segment,wave_segment,feature_segment,
feature_wrapper_segment}.py.  Contract (segment.py:7-25): a method object carries the payload's
`sampling_rate` (items per second) and implements length / pad / pick / concat; a `Segment`
(segment.py:28-59) is a (start_time, data, method) triple whose end time is
start_time + length / sampling_rate.
"""
from abc import ABC, abstractmethod
from typing import Generic, Iterable, List, Optional, TypeVar

import numpy

from .feature import AcousticFeature, AcousticFeatureWrapper

T = TypeVar('T')


class BaseSegmentMethod(ABC, Generic[T]):
    def __init__(self, sampling_rate: int):
        self.sampling_rate = sampling_rate

    @abstractmethod
    def length(self, data: T) -> int: ...

    @abstractmethod
    def pad(self, width: int) -> T: ...

    @abstractmethod
    def pick(self, data: T, first: int, last: int) -> T: ...

    @abstractmethod
    def concat(self, datas: Iterable[T]) -> T: ...


class Segment(tuple, Generic[T]):
    """Immutable-ish record; also unpacks as a 3-tuple like the reference's tuple subclass."""

    def __new__(cls, start_time: float, data: T, method: BaseSegmentMethod):
        obj = super().__new__(cls, (start_time, data, method))
        obj.start_time = start_time
        obj.data = data
        obj.method = method
        return obj

    sampling_rate = property(lambda self: self.method.sampling_rate)
    length = property(lambda self: self.method.length(self.data))
    time_length = property(lambda self: self.length / self.sampling_rate)
    end_time = property(lambda self: self.time_length + self.start_time)


class WaveSegmentMethod(BaseSegmentMethod[numpy.ndarray]):
    """float32 mono samples (wave_segment.py:8-19)."""

    def length(self, data):
        return len(data)

    def pad(self, width):
        return numpy.zeros(shape=width, dtype=numpy.float32)

    def pick(self, data, first, last):
        return data[first:last]

    def concat(self, datas):
        return numpy.concatenate(list(datas))


class FeatureSegmentMethod(BaseSegmentMethod[AcousticFeature]):
    """Converted features on their way to the vocoder: keys f0/ap/sp/voiced (feature_segment.py:9-37)."""

    def __init__(self, sampling_rate: int, wave_sampling_rate: int, order: int):
        super().__init__(sampling_rate=sampling_rate)
        self.wave_sampling_rate = wave_sampling_rate
        self.order = order
        self._keys = ['f0', 'ap', 'sp', 'voiced']

    def length(self, data):
        return len(data.f0)

    def pad(self, width):
        sizes = AcousticFeature.get_sizes(sampling_rate=self.wave_sampling_rate, order=self.order)
        return AcousticFeature.silent(width, sizes=sizes, keys=self._keys)

    def pick(self, data, first, last):
        return data.pick(first, last, keys=self._keys)

    def concat(self, datas):
        return AcousticFeature.concatenate(list(datas), keys=self._keys)


class FeatureWrapperSegmentMethod(BaseSegmentMethod[AcousticFeatureWrapper]):
    """Analysis features + the sample-aligned waveform: keys f0/ap/mc/voiced
    (feature_wrapper_segment.py:10-49)."""

    def __init__(self, sampling_rate: int, wave_sampling_rate: int, order: int, frame_period: int,
                 keys: Optional[List[str]] = None):
        super().__init__(sampling_rate=sampling_rate)
        self.wave_sampling_rate = wave_sampling_rate
        self.order = order
        self.frame_period = frame_period
        self._keys = ['f0', 'ap', 'mc', 'voiced'] if keys is None else keys

    def length(self, data):
        return len(data.f0)

    def pad(self, width):
        sizes = AcousticFeature.get_sizes(sampling_rate=self.wave_sampling_rate, order=self.order)
        silent = AcousticFeatureWrapper.silent_wrapper(
            width, sizes=sizes, keys=self._keys, frame_period=self.frame_period,
            sampling_rate=self.wave_sampling_rate, wave_dtype=numpy.float32)
        return silent.astype_only_float_wrapper(numpy.float32)

    def pick(self, data, first, last):
        return data.pick_wrapper(first, last, keys=self._keys, frame_period=self.frame_period)

    def concat(self, datas):
        return AcousticFeatureWrapper.concatenate_wrapper(list(datas), keys=self._keys)
