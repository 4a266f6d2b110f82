"""CPU (no GPU needed): bench.py's reference arm prints ONE JSON line with the contract's keys, rank 0 alone works under
a multi-rank launch, and the B200 arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_AVAILABLE = os.path.isdir('/root/reference') or os.path.exists(os.path.join(REPO, 'oracle', '_ref', 'dreamgnn_reference.zip'))


def _run(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                           '--cpu-scale', '0.0005'] + extra, capture_output=True, text=True, cwd=REPO, env=e, timeout=600)


@pytest.mark.skipif(not REF_AVAILABLE, reason='neither /root/reference nor the staged archive is present')
def test_reference_arm_prints_one_contract_line():
    r = _run(['--gpus', '1'])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout                              # everything else goes to stderr
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'aggregated_edges_per_sec' and d['unit'] == 'GE/s'
    assert d['higher_is_better'] is True and d['n_gpus'] == 1 and d['steps'] == 1 and d['warmup'] == 0
    assert d['vs_baseline'] is None and d['data'] == 'synthetic' and 'workload' in d['config'] and 'model' not in d['config']
    assert d['value'] > 0 and d['ms_per_step'] > 0
    cb = d['cpu_baseline']
    assert cb['kind'] == 'reference' and cb['cores'] >= 1 and cb['value'] == d['value'] and 'train()' in cb['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 'GE/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['gpu_launches'] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(['--gpus', '2'], env={'RANK': '1', 'LOCAL_RANK': '1', 'WORLD_SIZE': '2'})
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_b200_arm_refuses_to_run_without_cuda():
    import torch as th
    if th.cuda.is_available():
        pytest.skip('a CUDA device is present')
    r = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--steps', '1', '--warmup', '0'], capture_output=True,
                       text=True, cwd=REPO, timeout=600)
    assert r.returncode != 0 and 'no CPU fallback' in r.stderr and r.stdout.strip() == ''
