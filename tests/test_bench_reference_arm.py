"""bench.py --impl reference (the CPU arm the driver runs next to the GPU arm): one valid JSON line with the
contract's keys, on a reduced workload so the CPU suite stays fast."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '1', '--batch', '4', '--length', '6'], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['unit'] == 'sentences/s' and line['higher_is_better'] is True
    assert line['value'] > 0 and line['ms_per_step'] > 0 and line['n_gpus'] == 1
    # 'reference' = the unmodified reference staged under oracle/_ref (oracle/make_ref.py); 'port' only without it
    staged = os.path.isdir(os.path.join(ROOT, 'oracle', '_ref', 'cliora', 'net'))
    assert line['cpu_baseline']['kind'] == ('reference' if staged else 'port') and line['cpu_baseline']['cores'] >= 1
    assert len(out.stdout.strip().splitlines()) == 1      # the reference's own prints must not reach stdout
    assert line['cpu_baseline']['value'] == line['value'] == line['e2e']['value']
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in line['config'] and line['metric'].startswith('train sentences/sec')


def test_non_rank0_reference_arm_is_silent():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ''


def test_staged_reference_is_byte_identical_to_the_checkout():
    """oracle/_ref is a byte-for-byte copy of the reference package (oracle/make_ref.py): every file listed in its
    manifest hashes to the recorded value, and -- where the checkout is present -- to the checkout's file."""
    import hashlib
    import pytest
    ref = os.path.join(ROOT, 'oracle', '_ref')
    if not os.path.exists(os.path.join(ref, 'MANIFEST.json')):
        pytest.skip('oracle/_ref not staged')
    man = json.load(open(os.path.join(ref, 'MANIFEST.json')))
    assert len(man['files']) > 20
    for rel, digest in man['files'].items():
        assert hashlib.sha256(open(os.path.join(ref, rel), 'rb').read()).hexdigest() == digest, rel
        src = os.path.join('/root/reference', rel)
        if os.path.exists(src):
            assert hashlib.sha256(open(src, 'rb').read()).hexdigest() == digest, rel
