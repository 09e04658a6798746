"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/navgym_b200.h declares.  No compute calls (there is no GPU here)."""
import ctypes as C
import os
import re

import pytest

from nav_gym_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    _lib.build()
    return _lib.load()


def _declared():
    src = open(os.path.join(ROOT, 'include', 'navgym_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(navgym_[a-z0-9_]+)\s*\(', src)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.EXPORTS) == names


def test_struct_layout_matches(lib):
    assert lib.navgym_sizeof_step_args() == C.sizeof(_lib.StepArgs)
    assert lib.navgym_sizeof_map() == C.sizeof(_lib.MapT)
    assert lib.navgym_abi_version() == 1


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError):
        _lib.require_device()
    from nav_gym_b200.batched_env import BatchedNavGym
    import numpy as np
    m = dict(data=np.zeros((8, 8), np.int8), origin=(0, 0), resolution=0.05, width=8, height=8)
    with pytest.raises(RuntimeError):
        BatchedNavGym(1, [m])


def test_sass_is_sm100a():
    out = os.popen('cuobjdump -lelf %s 2>/dev/null' % _lib.SO).read()
    assert 'sm_100a' in out


def test_plain_c_client_links_and_runs(lib, tmp_path):
    """include/navgym_b200.h is valid C99 and a C program linked against the shared library
    sees the same ABI (struct sizes, no-op empty batch, host helper) -- the view a cgo / JNI /
    FFI host has of the boundary."""
    import shutil
    import subprocess
    from nav_gym_b200 import _lib
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('no gcc')
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / 'abi_client')
    subprocess.check_call([gcc, '-std=c99', '-Wall', '-Wextra', '-pedantic', '-Werror',
                           '-I', os.path.dirname(_lib.HDR), os.path.join(here, 'abi_client.c'),
                           '-o', exe, '-L', os.path.dirname(_lib.SO), '-lnavgym_b200',
                           '-Wl,-rpath,' + os.path.dirname(_lib.SO)])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert 'abi ok' in out.stdout
