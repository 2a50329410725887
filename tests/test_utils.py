'''
Host utilities that sit either side of the path: the BPLUT reader / writer
(reference mod16/utils.py:81-140), the dominant-PFT vote (:29-78), the dense parameter
table the kernels gather from, and `MOD16Collection61` (mod16/models.py:27-51).  CPU only.
(The reader is also checked against the reference's own `restore_bplut` by
oracle/make_golden.py in the build container.)
'''
import io
import os

import numpy as np
import pytest

from mod16_b200 import MOD16, utils


def test_restore_bplut_layout():
    b = utils.restore_bplut(utils.BPLUT_C51)
    assert set(MOD16.required_parameters) <= set(b)
    for k in MOD16.required_parameters:
        assert b[k].shape == (13,) and b[k].dtype == np.float64
        assert np.isnan(b[k][0]) and np.isnan(b[k][11])
    assert np.isnan(b['beta']).all()                     # the C5.1 table has no beta row
    # PFT 7 (open shrubland) of the C5.1 table, as quoted in tests/verification/transpiration.c
    assert [b[k][7] for k in ('tmin_close', 'tmin_open', 'vpd_open', 'vpd_close')] == \
        [-8.0, 8.8, 650.0, 4400.0]
    t = utils.bplut_table(b, MOD16.required_parameters, default_beta=250.0)
    assert t.shape == (11, 13) and (t[10, [1, 7, 12]] == 250.0).all() and np.isnan(t[0, 0])


def test_write_bplut_round_trip(tmp_path):
    b = utils.restore_bplut(utils.BPLUT_C51)
    b['beta'] = np.where(np.isnan(b['beta']), 250.0, b['beta'])
    b['csl'] = b['csl'] * 1.5
    path = os.path.join(tmp_path, 'bplut.csv')
    utils.write_bplut(b, path)
    back = utils.restore_bplut(path)
    for k in MOD16.required_parameters:
        valid = list(utils.PFT_VALID)
        assert np.array_equal(back[k][valid], b[k][valid]), k
    # from an open stream, with a short file
    few = utils.restore_bplut(io.StringIO(open(path).read()), nrows=3)
    assert not np.isnan(few['tmin_close'][1]) and np.isnan(few['beta']).all()


def test_pft_dominant_votes_and_protocol_overrides():
    m = np.array([[1, 1, 4, 0, 11], [0, 11, 13, 255, 0], [7, 7, 10, 10, 10], [5, 5, 5, 5, 5]])
    assert list(utils.pft_dominant(m)) == [1.0, 0.0, 10.0, 5.0]
    sites = ['US-A10', 'CA-SF2', 'XX-001', 'US-NGC']
    assert list(utils.pft_dominant(m, site_list=sites)) == [0.0, 3.0, 10.0, 3.0]
    assert utils.pft_dominant(m).dtype == np.float32


def test_collection61_model_parameters():
    from mod16_b200.models import MOD16Collection61
    model = MOD16Collection61(7)
    assert model.tmin_close == -8.0 and model.vpd_close == 4400.0 and model.beta == 250
    with pytest.raises(AssertionError):
        MOD16Collection61(11)
    with pytest.raises(KeyError):
        MOD16({'tmin_close': 1.0})
    assert MOD16.required_parameters == [
        'tmin_close', 'tmin_open', 'vpd_open', 'vpd_close', 'gl_sh', 'gl_wv', 'g_cuticular',
        'csl', 'rbl_min', 'rbl_max', 'beta']
