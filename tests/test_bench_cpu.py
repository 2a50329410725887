'''
bench.py on the CPU side: the reference arm runs here (it times the unmodified reference from
oracle/_ref, or the numpy port when that has not been materialised), carries every key the
contract asks for, and its `config` is the object the GPU arm prints for the same arguments.
'''
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_and_same_config():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--steps', '1', '--warmup', '0', '--cpu-pixels', '256', '--T', '8'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step',
                'higher_is_better', 'config', 'cpu_baseline', 'e2e', 'dtype'):
        assert key in line, key
    assert line['impl'] == 'reference' and line['value'] > 0
    cb = line['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] >= 1
    assert set(cb['legs']) == {'f64_stock_promotion', 'f32_throughout'}
    assert line['value'] == max(cb['legs'].values())
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
    have_ref = os.path.exists(os.path.join(ROOT, 'oracle', '_ref', 'mod16', '__init__.py'))
    assert cb['kind'] == ('reference' if have_ref else 'port')
    # the same `config` as the GPU arm builds for these arguments
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    args = argparse.Namespace(workload='tile', pixels=bench.TILE_N, T=8, kind='physical',
                              dtype='float32', gate='global', gpus=1)
    assert line['config'] == bench.bench_config(args, 1, bench.TILE_N, 8)


def test_reference_copy_is_byte_identical_when_present():
    'oracle/_ref is the UNMODIFIED reference: its manifest hashes match the files.'
    import hashlib
    ref = os.path.join(ROOT, 'oracle', '_ref')
    man = os.path.join(ref, 'MANIFEST.json')
    if not os.path.exists(man):
        import pytest
        pytest.skip('oracle/_ref has not been materialised (python oracle/make_ref.py)')
    files = json.load(open(man))['files']
    assert 'mod16/__init__.py' in files
    for rel, digest in files.items():
        if len(digest) != 64:
            continue
        with open(os.path.join(ref, rel), 'rb') as f:
            assert hashlib.sha256(f.read()).hexdigest() == digest, rel
        src = os.path.join('/root/reference', rel)
        if os.path.exists(src):
            with open(src, 'rb') as f:
                assert hashlib.sha256(f.read()).hexdigest() == digest, rel
