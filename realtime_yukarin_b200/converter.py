"""Model-loading entry point (re-creation of realtime_voice_conversion/converter/yukarin_converter.py:13-60)."""
import logging
from pathlib import Path

from .models import AcousticConverter, F0Converter, SuperResolution
from .params import create_from_json, create_sr_from_json


class YukarinConverter(object):
    def __init__(self, acoustic_converter: AcousticConverter, super_resolution: SuperResolution):
        self.acoustic_converter = acoustic_converter
        self.super_resolution = super_resolution

    @staticmethod
    def make_yukarin_converter(input_statistics_path: Path, target_statistics_path: Path, stage1_model_path: Path,
                               stage1_config_path: Path, stage2_model_path: Path, stage2_config_path: Path):
        logger = logging.getLogger('encode')
        f0_converter = F0Converter(input_statistics=input_statistics_path, target_statistics=target_statistics_path)
        acoustic_converter = AcousticConverter(config=create_from_json(stage1_config_path), model_path=stage1_model_path, gpu=0,
                                               f0_converter=f0_converter, out_sampling_rate=24000)
        logger.info('model 1 loaded!')
        super_resolution = SuperResolution(config=create_sr_from_json(stage2_config_path), model_path=stage2_model_path, gpu=0)
        logger.info('model 2 loaded!')
        return YukarinConverter(acoustic_converter=acoustic_converter, super_resolution=super_resolution)
