"""Parameter / config records the hot path reads.

These mirror the *field names* the reference touches on its third-party config objects:
  yukarin.param.AcousticParam      -> vocoder.py:30-36, encode_stream.py:18-24
  become_yukarin.param.Param       -> convert_stream.py:16,25-27, voice_changer.py:17
  yukarin.config.create_from_json / become_yukarin.config.sr_config.create_from_json
                                   -> check.py:31-32, converter/yukarin_converter.py:40,50
Only what the encode->convert->decode path needs is kept (no training/dataset options).
"""
import json
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional


@dataclass(frozen=True)
class AcousticParam:
    sampling_rate: int = 24000
    pad_second: float = 0.0
    threshold_db: Optional[float] = None
    frame_period: int = 5
    order: int = 8
    alpha: float = 0.466
    f0_floor: float = 71.0
    f0_ceil: float = 800.0
    fft_length: int = 1024
    dtype: str = 'float32'


@dataclass(frozen=True)
class VoiceParam:
    sample_rate: int = 24000
    top_db: Optional[float] = None
    pad_second: float = 0.0


@dataclass(frozen=True)
class AcousticFeatureParam:
    frame_period: int = 5
    order: int = 8
    alpha: float = 0.466
    f0_estimating_method: str = 'dio'


@dataclass(frozen=True)
class Param:
    voice_param: VoiceParam = field(default_factory=VoiceParam)
    acoustic_feature_param: AcousticFeatureParam = field(default_factory=AcousticFeatureParam)


# ---- stage-1 ("yukarin") config -------------------------------------------------------------

@dataclass(frozen=True)
class DatasetConfig:
    acoustic_param: AcousticParam = field(default_factory=AcousticParam)
    in_features: List[str] = field(default_factory=lambda: ['mc'])
    out_features: List[str] = field(default_factory=lambda: ['mc'])


@dataclass(frozen=True)
class ModelConfig:
    in_channels: int = 9
    out_channels: int = 9
    generator_base_channels: int = 64
    generator_extensive_layers: int = 8


@dataclass(frozen=True)
class Config:
    dataset: DatasetConfig = field(default_factory=DatasetConfig)
    model: ModelConfig = field(default_factory=ModelConfig)


def create_from_json(path) -> Config:
    d = json.loads(Path(path).read_text())
    ds = d.get('dataset', {})
    ap = ds.get('acoustic_param', ds.get('param', {}))
    known = {f for f in AcousticParam.__dataclass_fields__}
    acoustic_param = AcousticParam(**{k: v for k, v in ap.items() if k in known})
    dataset = DatasetConfig(
        acoustic_param=acoustic_param,
        in_features=list(ds.get('in_features', ['mc'])),
        out_features=list(ds.get('out_features', ['mc'])),
    )
    m = d.get('model', {})
    model = ModelConfig(
        in_channels=m.get('in_channels', 9),
        out_channels=m.get('out_channels', 9),
        generator_base_channels=m.get('generator_base_channels', 64),
        generator_extensive_layers=m.get('generator_extensive_layers', 8),
    )
    return Config(dataset=dataset, model=model)


# ---- stage-2 ("become_yukarin" super-resolution) config ------------------------------------

@dataclass(frozen=True)
class SRDatasetConfig:
    param: Param = field(default_factory=Param)


@dataclass(frozen=True)
class SRModelConfig:
    generator_base_channels: int = 64


@dataclass(frozen=True)
class SRConfig:
    dataset: SRDatasetConfig = field(default_factory=SRDatasetConfig)
    model: SRModelConfig = field(default_factory=SRModelConfig)


def create_sr_from_json(path) -> SRConfig:
    d = json.loads(Path(path).read_text())
    p = d.get('dataset', {}).get('param', {})
    vp = p.get('voice_param', {})
    afp = p.get('acoustic_feature_param', {})
    param = Param(
        voice_param=VoiceParam(**{k: v for k, v in vp.items() if k in VoiceParam.__dataclass_fields__}),
        acoustic_feature_param=AcousticFeatureParam(
            **{k: v for k, v in afp.items() if k in AcousticFeatureParam.__dataclass_fields__}),
    )
    m = d.get('model', {})
    return SRConfig(dataset=SRDatasetConfig(param=param),
                    model=SRModelConfig(generator_base_channels=m.get('generator_base_channels', 64)))
