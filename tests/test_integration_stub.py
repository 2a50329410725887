'''
INTEGRATION.md shows the ctypes stub a maintainer of the reference would add
(`mod16/_b200.py`).  This test cuts that very code block out of the document and runs it on
the GPU against the oracle, so the document cannot drift from the C ABI.
'''
import os
import re

import numpy as np
import pytest

from mod16_b200 import build, synthetic
from oracle import mod16_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_the_documented_ctypes_stub_runs_and_matches(cuda_lib):
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    blocks = re.findall(r'```python\n(.*?)```', text, flags=re.S)
    stub = [b for b in blocks if 'ctypes.CDLL' in b and 'def _et(' in b]
    assert len(stub) == 1
    code = stub[0].replace("ctypes.CDLL('libmod16_b200.so')", 'ctypes.CDLL(%r)' % build.LIB_PATH)
    scope = {}
    exec(compile(code, 'INTEGRATION.md', 'exec'), scope)
    T, N = 12, 1024
    d = synthetic.host_drivers(T, N, 501, dtype=np.float32)
    drv = [d[k] for k in synthetic.DRIVER_NAMES]
    params = [-8., 8.8, 650., 4400., .02, .02, 1e-5, .0055, 60., 95., 250.]
    got = scope['_et'](params, *drv)
    d64 = [v.astype(np.float64) if isinstance(v, np.ndarray) else v for v in drv]
    with np.errstate(all='ignore'):
        want = orc.static_et(params, *d64)
    assert got.shape == (T, N)
    assert np.all(np.abs(got - want) <= np.maximum(1e-5 * np.abs(want), 1e-3))
