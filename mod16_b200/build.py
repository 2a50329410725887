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
SOURCES = ['mod16_forward.cu', 'mod16_forward_fast.cu', 'mod16_aggregate.cu',
           'mod16_objective.cu', 'mod16_elementwise.cu']
HEADERS = ['mod16_math.cuh', 'mod16_fast.cuh', 'mod16_common.cuh', 'mod16_pipe.cuh',
           'mod16_objective.cuh', os.path.join('..', '..', 'include', 'mod16_b200.h')]
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
    '--expt-relaxed-constexpr', '-Xcompiler', '-fPIC', '-Xptxas', '-v',
    '--fmad=true', '-shared', '--threads', '0',
]


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > built for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    'Compiles the library if a source is newer than it.  Returns the library path.'
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [NVCC] + FLAGS + ['-o', LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(LIB_DIR, 'build.log'), 'w') as f:
        f.write(' '.join(cmd) + '\n' + log)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + log[-6000:])
    if verbose:
        print(log)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
