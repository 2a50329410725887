'''
Pre-configured model variants (mirrors the reference's ``mod16/models.py``).
'''
import numpy as np

from .model import MOD16, PFT_VALID
from .utils import BPLUT_C51, restore_bplut

PFT_ALL = {
    'Evergreen Needleleaf Forest (ENF)': 1, 'Evergreen Broadleaf Forest (EBF)': 2,
    'Deciduous Needleleaf Forest (DNF)': 3, 'Deciduous Broadleaf Forest (DBF)': 4,
    'Mixed Forest (MF) ': 5, 'Closed Shrublands (CSH)': 6, 'Open Shrublands (OSH)': 7,
    'Woody Savannas (WSV)': 8, 'Savannas (SAV)': 9, 'Grasslands (GRS)': 10,
    'Croplands (CRO)': 12,
}


class MOD16Collection61(MOD16):
    '''
    The MOD16 Collection 6.1 model for one Plant Functional Type: the C5.1 BPLUT column
    of ``pft`` with ``beta = 250`` where the table has none (reference models.py:36-51).
    '''

    def __init__(self, pft):
        assert pft in PFT_VALID, 'Not a recognized numeric PFT code; should be one of: ' \
            + ','.join(map(str, PFT_VALID))
        table = restore_bplut(BPLUT_C51)
        params = dict((key, table[key][pft]) for key in table.keys())
        if np.isnan(params['beta']):
            params['beta'] = 250
        super().__init__(params=params)
