'''
TEST INFRASTRUCTURE ONLY -- stand-in for the external package ``mod17``
(PyPI / GitHub arthur-e/MOD17, pinned by the reference only as ``mod17>=0.1.1``,
/root/reference/pyproject.toml:22), which is NOT present in /root/reference and
cannot be installed here (no network).

The reference imports exactly one name from it on the forward path:
``from mod17 import linear_constraint`` (/root/reference/mod16/__init__.py:104;
call sites :325-327 and :1148-1150).  This module restates the published
behaviour of that one function so that the UNMODIFIED reference can be imported
in the build container by ``oracle/make_golden.py``.  It never travels into the
product path.

Published behaviour (MOD17 docs, "linear_constraint"): returns a ramp function
f(x) that rises 0 -> 1 between ``xmin`` and ``xmax``; ``form='reversed'`` falls
1 -> 0; ``form='binary'`` is a step at ``xmax``.  Pinned numerically by the
reference's own known answers: tests/tests.py:113-121 (1.248 kg m-2 h-1 needs
the reversed VPD ramp's interior) and the ramp values listed in
tests/verification/transpiration.c:154-155.
'''
import numpy as np


def linear_constraint(xmin, xmax, form=None):
    assert form is None or form in ('reversed', 'binary')
    assert form == 'binary' or np.any(xmax >= xmin)
    if form == 'reversed':
        return lambda x: np.where(
            x >= xmax, 0, np.where(x < xmin, 1, 1 - np.divide(
                np.subtract(x, xmin), xmax - xmin)))
    if form == 'binary':
        return lambda x: np.where(x == xmax, 1, 0)
    return lambda x: np.where(
        x >= xmax, 1, np.where(x < xmin, 0, np.divide(
            np.subtract(x, xmin), xmax - xmin)))
