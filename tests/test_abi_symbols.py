'''
The C ABI: every function include/mod16_b200.h declares is exported by the built library
and bound by mod16_b200/_abi.py; flag words agree between the header and the binding; the
product package never reaches into oracle/.  No compute calls (no GPU needed).
'''
import ctypes
import os
import re

from mod16_b200 import _abi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'mod16_b200.h')


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(mod16_[a-z0-9_]+)\s*\(', src)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    lib = ctypes.CDLL(path)
    names = declared_functions()
    assert 'mod16_forward_f32' in names and 'mod16_forward_f64' in names
    for name in names:
        assert hasattr(lib, name), 'library does not export ' + name
    assert sorted(_abi.EXPORTS) == names, 'binding and header disagree'


def test_version_and_error_strings():
    lib = _abi.load()
    assert lib.mod16_abi_version() == _abi.ABI_VERSION
    assert lib.mod16_error_string(0) == b'ok'
    assert b'NULL' in lib.mod16_error_string(-1)
    # argument validation happens before any CUDA call
    assert lib.mod16_forward_f32(None, None, None, 1, 1, 0, None) == -1


def test_flag_words_match_header():
    src = open(HEADER).read()
    macros = dict((k, int(v, 16)) for k, v in
                  re.findall(r'#define\s+MOD16_([A-Z_]+)\s+(0x[0-9A-Fa-f]+)u', src))
    for k, v in macros.items():
        if hasattr(_abi, k):
            assert getattr(_abi, k) == v, k
    assert _abi.GATE_FLAGS['global'] == macros['GATE_GLOBAL']
    assert ctypes.sizeof(_abi.Driver) == 32
    assert ctypes.sizeof(_abi.Params) == 8 + 11 * 13 * 8 + 11 * 8 + 11 * 8


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'mod16_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', text, flags=re.M), f
                assert 'mod16_oracle' not in text, f
