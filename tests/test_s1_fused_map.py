"""CPU check of the fused stage-1 kernel's index arithmetic (csrc/s1_map.h): tests/s1_fused_emulate.cpp replays the kernel's loops
(staging, ldmatrix addresses, B-fragment packing, mma.sync m16n8k16 fragment layouts, split-K partials, epilogue coordinates) on the
CPU and compares every k4 layer with a direct (transposed) convolution in the model file's weight layout."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope='module')
def emulator(tmp_path_factory):
    exe = tmp_path_factory.mktemp('s1emu') / 's1_fused_emulate'
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-o', str(exe), str(ROOT / 'tests' / 's1_fused_emulate.cpp')])
    return exe


@pytest.mark.parametrize('tp1,cluster', [(128, 16), (256, 8), (384, 16), (384, 8), (512, 16), (640, 16), (1024, 16), (2048, 16)])
def test_fused_stage1_index_maps(emulator, tp1, cluster):
    r = subprocess.run([str(emulator), '64', str(tp1), str(cluster)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count(' ok') == 14
