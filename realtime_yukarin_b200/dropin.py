"""Import aliases that make this package answer to the names the reference's drivers import:

    realtime_voice_conversion.{config, stream[.base_stream|...], segment.*, yukarin_wrapper.*, converter.*, worker.utility}
    yukarin[.acoustic_feature|.wave|.param|.config|.f0_converter],  become_yukarin[.param|.config.sr_config]
    librosa.{load, output.write_wav} (wav I/O of check.py; only if librosa itself is not installed)

Nothing is copied: each alias module is a thin namespace whose attributes are this package's objects
(import sites: check.py:7-17, tests/test_*.py of the reference).
"""
import sys
import types


_skip = set()


def _module(name: str, **attrs):
    if name.partition('.')[0] in _skip:
        return None
    m = types.ModuleType(name)
    m.__ryk_alias__ = True
    m.__dict__.update(attrs)
    m.__path__ = []          # behave like a package so that submodule imports resolve through sys.modules
    sys.modules[name] = m
    parent, _, child = name.rpartition('.')
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


def _importable(top: str) -> bool:
    """True when a REAL package of that name can be imported (an alias this module registered earlier does not count)."""
    import importlib.util
    m = sys.modules.get(top)
    if m is not None:
        return not getattr(m, '__ryk_alias__', False)
    try:
        return importlib.util.find_spec(top) is not None
    except (ImportError, ValueError):
        return False


def install(force: bool = True) -> None:
    """Register the aliases.  force=True (default, what a drop-in switch wants): this package answers to the reference's names even
    if the original packages are installed.  force=False: a family (`realtime_voice_conversion`, `yukarin`, `become_yukarin`) is
    aliased only when no real package of that name is importable, so an existing installation is never shadowed."""
    from . import config, converter, feature, models, params, segment, stream, vocoder, voice_changer, wave_io, worker

    global _skip
    _skip = set() if force else {top for top in ('realtime_voice_conversion', 'yukarin', 'become_yukarin') if _importable(top)}

    rvc = 'realtime_voice_conversion'
    _module(rvc)
    _module(f'{rvc}.config', Config=config.Config, VocodeMode=config.VocodeMode)
    st = dict(BaseStream=stream.BaseStream, EncodeStream=stream.EncodeStream, ConvertStream=stream.ConvertStream,
              DecodeStream=stream.DecodeStream, StreamWrapper=stream.StreamWrapper)
    _module(f'{rvc}.stream', **st)
    _module(f'{rvc}.stream.base_stream', BaseStream=stream.BaseStream)
    _module(f'{rvc}.stream.encode_stream', EncodeStream=stream.EncodeStream)
    _module(f'{rvc}.stream.convert_stream', ConvertStream=stream.ConvertStream)
    _module(f'{rvc}.stream.decode_stream', DecodeStream=stream.DecodeStream)
    _module(f'{rvc}.stream.stream_wrapper', StreamWrapper=stream.StreamWrapper)
    _module(f'{rvc}.segment')
    _module(f'{rvc}.segment.segment', BaseSegmentMethod=segment.BaseSegmentMethod, Segment=segment.Segment)
    _module(f'{rvc}.segment.wave_segment', WaveSegmentMethod=segment.WaveSegmentMethod)
    _module(f'{rvc}.segment.feature_segment', FeatureSegmentMethod=segment.FeatureSegmentMethod)
    _module(f'{rvc}.segment.feature_wrapper_segment', FeatureWrapperSegmentMethod=segment.FeatureWrapperSegmentMethod)
    _module(f'{rvc}.yukarin_wrapper')
    _module(f'{rvc}.yukarin_wrapper.vocoder', Vocoder=vocoder.Vocoder, RealtimeVocoder=vocoder.RealtimeVocoder)
    _module(f'{rvc}.yukarin_wrapper.voice_changer', VoiceChanger=voice_changer.VoiceChanger,
            AcousticFeatureWrapper=feature.AcousticFeatureWrapper)
    _module(f'{rvc}.yukarin_wrapper.acoustic_feature_wrapper', AcousticFeatureWrapper=feature.AcousticFeatureWrapper,
            CrepeAcousticFeatureWrapper=vocoder.CrepeAcousticFeatureWrapper)
    _module(f'{rvc}.worker')
    _module(f'{rvc}.worker.utility', Item=worker.Item, init_logger=worker.init_logger)
    _module(f'{rvc}.converter')
    _module(f'{rvc}.converter.yukarin_converter', YukarinConverter=converter.YukarinConverter)

    # `librosa` as check.py uses it (check.py:80 load, check.py:112 output.write_wav) -- only when the real package is absent
    try:
        import librosa  # noqa: F401
    except ImportError:
        def _load(path, sr=22050, **_):       # librosa.load's default rate; check.py:80 always passes sr=
            w = wave_io.load_wave(path, sr)
            return w.wave, w.sampling_rate
        _module('librosa', load=_load)
        _module('librosa.output', write_wav=lambda path, y, sr, **_: wave_io.write_wav(path, y, sr))

    _module('yukarin', AcousticConverter=models.AcousticConverter, AcousticFeature=feature.AcousticFeature, Wave=feature.Wave)
    _module('yukarin.acoustic_feature', AcousticFeature=feature.AcousticFeature)
    _module('yukarin.wave', Wave=feature.Wave)
    _module('yukarin.param', AcousticParam=params.AcousticParam)
    _module('yukarin.config', create_from_json=params.create_from_json, Config=params.Config)
    _module('yukarin.f0_converter', F0Converter=models.F0Converter)
    _module('become_yukarin', SuperResolution=models.SuperResolution)
    _module('become_yukarin.param', Param=params.Param, VoiceParam=params.VoiceParam,
            AcousticFeatureParam=params.AcousticFeatureParam)
    _module('become_yukarin.config')
    _module('become_yukarin.config.sr_config', create_from_json=params.create_sr_from_json, SRConfig=params.SRConfig)
