"""Model-loading entry point with the call surface of realtime_voice_conversion/converter/yukarin_converter.py:13-60
(`YukarinConverter.make_yukarin_converter(<six paths>)` -> object with `.acoustic_converter` and `.super_resolution`).

Loading a stage = parse its config.json, read the Chainer-layout npz, fold BatchNorm + bias into per-channel scale/shift and upload
the layers to the engine (models.py); the log-f0 statistics of both speakers go into the F0Converter that stage 1 carries."""
import logging
from pathlib import Path
from typing import NamedTuple

from . import models, params

_LOG = logging.getLogger('encode')          # the reference logs model loading under this name (yukarin_converter.py:31)
OUTPUT_RATE = 24000                         # yukarin_converter.py:45: stage 1 is told the vocoder's output rate


class StagePaths(NamedTuple):
    model: Path
    config: Path


class YukarinConverter(object):
    """The two loaded stages; `VoiceChanger(acoustic_converter=..., super_resolution=...)` consumes them."""

    def __init__(self, acoustic_converter: models.AcousticConverter, super_resolution: models.SuperResolution):
        self.acoustic_converter = acoustic_converter
        self.super_resolution = super_resolution

    @classmethod
    def _load_stage1(cls, paths: StagePaths, f0_converter: models.F0Converter) -> models.AcousticConverter:
        stage1 = models.AcousticConverter(config=params.create_from_json(paths.config), model_path=paths.model, gpu=0,
                                          f0_converter=f0_converter, out_sampling_rate=OUTPUT_RATE)
        _LOG.info('model 1 loaded!')
        return stage1

    @classmethod
    def _load_stage2(cls, paths: StagePaths) -> models.SuperResolution:
        stage2 = models.SuperResolution(config=params.create_sr_from_json(paths.config), model_path=paths.model, gpu=0)
        _LOG.info('model 2 loaded!')
        return stage2

    @staticmethod
    def make_yukarin_converter(input_statistics_path: Path, target_statistics_path: Path, stage1_model_path: Path,
                               stage1_config_path: Path, stage2_model_path: Path, stage2_config_path: Path) -> 'YukarinConverter':
        f0_converter = models.F0Converter(input_statistics=input_statistics_path, target_statistics=target_statistics_path)
        return YukarinConverter(
            acoustic_converter=YukarinConverter._load_stage1(StagePaths(stage1_model_path, stage1_config_path), f0_converter),
            super_resolution=YukarinConverter._load_stage2(StagePaths(stage2_model_path, stage2_config_path)))
