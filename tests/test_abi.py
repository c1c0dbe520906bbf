"""The C-ABI boundary without a GPU: the shared library loads, exports every function include/ryk.h
declares (and the Python binding's list agrees with the header), and the product path fails loudly --
no CPU fallback -- when there is no device."""
import ctypes
import re
from pathlib import Path

import pytest

from realtime_yukarin_b200 import engine as eng

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / 'include' / 'ryk.h'


def declared_functions():
    text = re.sub(r'/\*.*?\*/', '', HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r'\b(ryk_[a-z0-9_]+)\s*\(', text)))


def test_header_cites_reference_interfaces():
    text = HEADER.read_text()
    for needle in ('vocoder.py:79-87', 'vocoder.py:99', 'voice_changer.py:24-42', 'acoustic_feature_wrapper.py:28-33'):
        assert needle in text


def test_library_exports_every_declared_symbol():
    lib = eng.load_library()
    names = declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f'{n} is declared in ryk.h but not exported by libryk.so'
    assert sorted(eng.EXPORTED_SYMBOLS) == names
    assert lib.ryk_abi_version() == 1
    assert lib.ryk_world_num_frames(7200, 24000, ctypes.c_double(5.0)) == 61


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    with pytest.raises(eng.RykError):
        eng.Engine(device=0)


def test_product_package_does_not_import_the_oracle():
    pkg = ROOT / 'realtime_yukarin_b200'
    for path in list(pkg.rglob('*.py')) + list(pkg.rglob('*.cu')) + list(pkg.rglob('*.h')) + list(pkg.rglob('*.cuh')):
        text = path.read_text()
        assert 'import oracle' not in text and 'from oracle' not in text and 'world_oracle' not in text, path


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/ryk.h must compile as C99 on its own (no C++ / CUDA / torch types), and every entry point
    must be callable from C with the declared prototype (compile + link a C translation unit against libryk.so)."""
    import subprocess
    src = tmp_path / 'use_ryk.c'
    src.write_text(
        '#include "ryk.h"\n'
        'int main(void) {\n'
        '  ryk_engine* e = 0; int n = 0; double pw = 0; int keep = 0; float y[4];\n'
        '  if (ryk_abi_version() != 1) return 1;\n'
        '  if (ryk_world_synthesize_length(200, 5.0, 24000) != 24000) return 2;\n'
        '  if (ryk_resample_length(147, 80, 147) != 80) return 3;\n'
        '  if (ryk_world_num_frames(7200, 24000, 5.0) != 61) return 4;\n'
        '  (void)e; (void)n; (void)pw; (void)keep; (void)y;\n'
        '  return 0;\n'
        '}\n')
    lib_dir = ROOT / 'realtime_yukarin_b200' / 'csrc'
    exe = tmp_path / 'use_ryk'
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Werror', '-pedantic', '-I', str(ROOT / 'include'), str(src), '-o', str(exe),
                           '-L', str(lib_dir), '-lryk', f'-Wl,-rpath,{lib_dir}', '-Wl,-rpath,/usr/local/cuda/lib64'])
    assert subprocess.call([str(exe)]) == 0          # pure host-side entry points: no GPU needed
