"""run-time configuration record (re-creation of realtime_voice_conversion/config.py:8-71)."""
from dataclasses import dataclass
from enum import Enum
from pathlib import Path

import yaml


class VocodeMode(Enum):
    WORLD = 'world'
    CREPE = 'crepe'


_PATH_FIELDS = ('input_statistics_path', 'target_statistics_path', 'stage1_model_path', 'stage1_config_path',
                'stage2_model_path', 'stage2_config_path')


@dataclass(frozen=True)
class Config:
    input_device_name: str
    output_device_name: str
    input_rate: int
    output_rate: int
    frame_period: float
    buffer_time: float
    extract_f0_mode: VocodeMode
    vocoder_buffer_size: int
    input_scale: float
    output_scale: float
    input_silent_threshold: float
    output_silent_threshold: float
    encode_extra_time: float
    convert_extra_time: float
    decode_extra_time: float
    input_statistics_path: Path
    target_statistics_path: Path
    stage1_model_path: Path
    stage1_config_path: Path
    stage2_model_path: Path
    stage2_config_path: Path

    @property
    def in_audio_chunk(self) -> int:
        return round(self.input_rate * self.buffer_time)

    @property
    def out_audio_chunk(self) -> int:
        return round(self.output_rate * self.buffer_time)

    @staticmethod
    def from_yaml(path: Path) -> 'Config':
        with Path(path).open() as f:
            d = yaml.safe_load(f)
        kw = {}
        for name in Config.__dataclass_fields__:
            v = d[name]
            if name == 'extract_f0_mode':
                v = VocodeMode(v)
            elif name in _PATH_FIELDS:
                v = Path(v)
            kw[name] = v
        return Config(**kw)
