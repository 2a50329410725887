'''
TEST / BENCH INFRASTRUCTURE ONLY -- materialises the UNMODIFIED reference forward model
under oracle/_ref/ so that it can be timed on the GPU box's host cores
(`bench.py --impl reference`, `cpu_baseline.kind = "reference"`) and imported by tests there.

    python oracle/make_ref.py        (in the build container, where /root/reference exists)

The reference is pure Python (MIT licence, /root/reference/LICENSE).  Its forward path lives
in three files -- mod16/__init__.py (the MOD16 class and the physics helpers), mod16/utils.py
(BPLUT reader) and mod16/models.py (MOD16Collection61) -- plus the BPLUT CSV; they are copied
byte for byte (sha256 recorded in oracle/_ref/MANIFEST.json) next to oracle/refshim/mod17,
the stand-in for the one external function the path imports (`mod17.linear_constraint`,
mod16/__init__.py:104).  calibration.py / sensitivity.py are not copied (pymc, SALib, h5py
are not installed; they are callers of the path, not part of it).

oracle/_ref/ is git-ignored: reference sources never enter this repository's history.  It is
NOT gpurun-ignored, so it travels to the GPU box like the built shared library.  Nothing in
the product (mod16_b200/) imports it.
'''
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
DST = os.path.join(HERE, '_ref')
FILES = ['mod16/__init__.py', 'mod16/utils.py', 'mod16/models.py', 'LICENSE']


def available():
    return os.path.exists(os.path.join(REF, 'mod16', '__init__.py'))


def build():
    if not available():
        raise SystemExit('%s is not present: oracle/_ref can only be made in the build container' % REF)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    files = list(FILES)
    data_dir = os.path.join(REF, 'mod16', 'data')
    files += ['mod16/data/' + f for f in sorted(os.listdir(data_dir)) if f.endswith('.csv')]
    for rel in files:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, 'rb') as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    shim = os.path.join(HERE, 'refshim', 'mod17', '__init__.py')
    os.makedirs(os.path.join(DST, 'mod17'), exist_ok=True)
    shutil.copyfile(shim, os.path.join(DST, 'mod17', '__init__.py'))
    manifest['mod17/__init__.py'] = 'oracle/refshim/mod17/__init__.py (stand-in, this repository)'
    with open(os.path.join(DST, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': REF, 'files': manifest}, f, indent=1)
    return DST


def load():
    '''
    Imports the reference package from oracle/_ref (None when it has not been materialised).
    Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may call this.
    '''
    if not os.path.exists(os.path.join(DST, 'mod16', '__init__.py')):
        return None
    if DST not in sys.path:
        sys.path.insert(0, DST)
    import mod16
    if not os.path.abspath(mod16.__file__).startswith(DST):
        raise ImportError('a different `mod16` is already imported: %s' % mod16.__file__)
    return mod16


if __name__ == '__main__':
    print(build())
