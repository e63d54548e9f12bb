"""The C-ABI library builds for sm_100a, loads, and exports exactly the symbols include/varsep.h declares
(no compute call is made: this runs on machines without a GPU)."""
import ctypes
import os
import re
import subprocess

from spatiotemporal_variable_separation_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'varsep.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(vs_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from spatiotemporal_variable_separation_b200.csrc.build import build
    lib_path = build()
    lib = ctypes.CDLL(lib_path)
    declared = header_functions()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/varsep.h but not exported'
    exported = subprocess.run(['nm', '-D', '--defined-only', lib_path], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r'\bT (vs_[a-z0-9_]+)$', exported, flags=re.M)))
    assert exported == declared, (set(exported) ^ set(declared))


def test_python_binding_covers_the_header():
    assert sorted(_lib.SIGNATURES) == header_functions()
    lib = _lib.load()
    assert lib.vs_abi_version() == 1
    assert lib.vs_last_error() is not None
    assert _lib.launch_count() >= 0


def test_sass_contains_blackwell_tensor_core_and_tma_instructions():
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG (B200_PROFILING.md)."""
    sass = subprocess.run(['cuobjdump', '-sass', _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ('UTCHMMA', 'LDTM', 'UTMALDG'):
        assert mnemonic in sass, mnemonic
    assert 'sm_100a' in sass
