'''
BPLUT I/O and land-cover helpers -- the host-side data boundary of the forward path.

Mirrors the interface of the reference's ``mod16/utils.py`` (``restore_bplut`` :81-117,
``write_bplut`` :120-145, ``pft_dominant`` :29-78, ``BPLUT_FIELD_LOOKUP`` :15-27) so that
a parameter dictionary built here is interchangeable with one built there: a dict of 11
float64 arrays of length 13 indexed by PFT code, NaN at codes 0 and 11 and for rows the
file does not carry (the C5.1 table has no ``beta`` row).

``bplut_table`` turns such a dict into the dense 11 x 13 table the CUDA kernels gather
from by land-cover code (row order = ``MOD16.required_parameters``).
'''
import csv
import os
from collections import Counter, OrderedDict
import numpy as np

PFT_VALID = (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12)
N_PFT_CODES = 13
DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')
BPLUT_C51 = os.path.join(
    DATA_DIR, 'MOD16_BPLUT_C5.1_05deg_MCD43B_Albedo_MERRA_GMAO.csv')

# CSV row label -> parameter name (reference mod16/utils.py:15-27)
BPLUT_FIELD_LOOKUP = OrderedDict([
    ('Tmin_min(C)', 'tmin_close'), ('Tmin_max(C)', 'tmin_open'),
    ('VPD_min(Pa)', 'vpd_open'), ('VPD_max(Pa)', 'vpd_close'),
    ('gl_sh((m/s)', 'gl_sh'), ('gl_e_wv(m/s)', 'gl_wv'),
    ('g_cuticular(m/s)', 'g_cuticular'), ('Cl(m/s)', 'csl'),
    ('RBL_MIN(s/m)', 'rbl_min'), ('RBL_MAX(s/m)', 'rbl_max'),
    ('beta', 'beta')])


def restore_bplut(path_or_buffer, nrows=11):
    '''
    Reads a BPLUT CSV (first column = row label, then one column per PFT in the order
    ENF, EBF, DNF, DBF, MF, CSH, OSH, WSV, SAV, GRS, CRO) into a dict of float64[13]
    arrays indexed by MCD12Q1 LC_Type2 code.  Unknown row labels are ignored; rows
    beyond ``nrows`` are not read.
    '''
    if hasattr(path_or_buffer, 'read'):
        rows = list(csv.reader(path_or_buffer))
    else:
        with open(path_or_buffer, 'r', newline='') as stream:
            rows = list(csv.reader(stream))
    table = dict((name, np.full((N_PFT_CODES,), np.nan))
                 for name in BPLUT_FIELD_LOOKUP.values())
    codes = list(PFT_VALID)
    for row in rows[1:1 + nrows]:
        if not row or row[0] not in BPLUT_FIELD_LOOKUP:
            continue
        values = [float(v) if v.strip() != '' else np.nan for v in row[1:1 + len(codes)]]
        table[BPLUT_FIELD_LOOKUP[row[0]]][codes[:len(values)]] = values
    return table


def write_bplut(params_dict, output_path, template=BPLUT_C51):
    'Writes a BPLUT dict back to CSV with the header of the C5.1 table.'
    with open(template, 'r', newline='') as stream:
        header = next(csv.reader(stream))
    with open(output_path, 'w', newline='') as stream:
        writer = csv.writer(stream)
        writer.writerow(header)
        for label, name in BPLUT_FIELD_LOOKUP.items():
            writer.writerow([label] + [params_dict[name][p] for p in PFT_VALID])


def pft_dominant(pft_map, site_list=None, valid_pft=PFT_VALID):
    '''
    Majority-vote PFT per site from an (N x M) sub-grid of land-cover codes; 0 where a
    site has no valid code.  With ``site_list`` the Cal/Val protocol overrides apply
    (US-A10 -> 0; CA-SF2, CA-SF3, US-NGC -> 3), as in the reference (utils.py:62-78).
    '''
    pft_map = np.asarray(pft_map)
    out = np.zeros(pft_map.shape[0], np.float32)
    for i, cells in enumerate(pft_map):
        votes = Counter(c for c in cells.tolist() if c in valid_pft)
        if votes:
            out[i] = votes.most_common()[0][0]
    if site_list is not None:
        site_list = list(site_list)
        if 'US-A10' in site_list:
            out[site_list.index('US-A10')] = 0
        for sid in ('CA-SF2', 'CA-SF3', 'US-NGC'):
            if sid in site_list:
                out[site_list.index(sid)] = 3
    return out


def bplut_table(params_dict, order, default_beta=None):
    '''
    Dense float64 (11, 13) table in ``order`` (= ``MOD16.required_parameters``) for the
    kernels' land-cover gather.  ``default_beta`` replaces an all-NaN ``beta`` row, which
    is how every reference caller treats the C5.1 table (models.py:49-50,
    calibration.py:807-808).
    '''
    table = np.full((len(order), N_PFT_CODES), np.nan)
    for i, name in enumerate(order):
        row = np.asarray(params_dict[name], np.float64).ravel()
        table[i, :row.size] = row[:N_PFT_CODES]
    if default_beta is not None:
        b = list(order).index('beta')
        table[b, np.isnan(table[b])] = default_beta
    return table
