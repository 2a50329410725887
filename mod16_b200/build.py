'''
In-tree build of the CUDA shared library (sm_100a only).

    python -m mod16_b200.build [--force]

produces ``mod16_b200/lib/libmod16_b200.so`` with nvcc; the file is git-ignored but
travels to the GPU box with the repository snapshot.  There is no CPU fallback: if the
library is missing, ``mod16_b200._abi`` raises at first use.
'''
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
LIB_DIR = os.path.join(PKG, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libmod16_b200.so')
SOURCES = sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith('.cuh')) + [
    os.path.join('..', '..', 'include', 'mod16_b200.h')]
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
    '--expt-relaxed-constexpr', '-Xcompiler', '-fPIC', '-Xptxas', '-v',
    '--fmad=true',
]


HASH_PATH = LIB_PATH + '.srchash'


def source_hash():
    'sha256 over the sources, headers and compiler flags the library is built from.'
    import hashlib
    h = hashlib.sha256(' '.join(FLAGS).encode())
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), 'rb') as f:
            h.update(name.encode() + b'\0' + f.read())
    return h.hexdigest()


def is_stale():
    '''
    True when the library is missing or was built from other sources than the ones in the
    tree (content hash recorded next to the binary -- modification times do not survive a
    checkout or the snapshot that carries the binary to the GPU box).
    '''
    if not (os.path.exists(LIB_PATH) and os.path.exists(HASH_PATH)):
        return True
    with open(HASH_PATH) as f:
        return f.read().strip() != source_hash()


def build(force=False, verbose=False):
    'Compiles the library if a source is newer than it.  Returns the library path.'
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    # one nvcc per translation unit, all at once, then one link
    obj_dir = os.path.join(LIB_DIR, 'obj')
    os.makedirs(obj_dir, exist_ok=True)
    from concurrent.futures import ThreadPoolExecutor

    def compile_one(src):
        obj = os.path.join(obj_dir, src[:-3] + '.o')
        cmd = [NVCC] + FLAGS + ['-c', '-o', obj, os.path.join(CSRC, src)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return obj, ' '.join(cmd) + '\n' + res.stdout + res.stderr, res.returncode
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        done = list(pool.map(compile_one, SOURCES))
    log = '\n'.join(d[1] for d in done)
    failed = any(d[2] != 0 for d in done)
    if not failed:
        cmd = [NVCC, '-shared', '-o', LIB_PATH] + [d[0] for d in done]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log += '\n' + ' '.join(cmd) + '\n' + res.stdout + res.stderr
        failed = res.returncode != 0
    with open(os.path.join(LIB_DIR, 'build.log'), 'w') as f:
        f.write(log)
    if failed:
        raise RuntimeError('nvcc failed:\n' + log[-6000:])
    with open(HASH_PATH, 'w') as f:
        f.write(source_hash() + '\n')
    if verbose:
        print(log)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
