"""Host-side legs of bench.py that run without a GPU: the cpu_baseline / --impl reference leg (the oracle port timed on
the host cores), the helpers behind the JSON line, and the line's contract keys for the reference arm."""
import json
import subprocess
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_helpers():
    assert bench.median([3.0, 1.0, 2.0]) == 2.0 and bench.median([4.0, 1.0]) in (1.0, 2.5, 4.0)
    peak, tc_peak, src = bench.measured_peaks()
    assert peak > 1000 and tc_peak > 100 and isinstance(src, str)
    traffic = bench.measured_traffic('yelp-lightgcn')                   # per-launch DRAM bytes from the committed ncu captures
    assert traffic['igcn_tc_candidates'] > 1e6 and all(v > 0 for v in traffic.values() if not isinstance(v, str))
    assert bench.measured_traffic('no-such-workload') == {}
    # algorithmic bytes of one full layer: nnz*(4+4) + (n+1)*8 + read X + write Y (+ add operands)
    n, nnz, D = 1000, 50000, 64
    assert bench.spmm_bytes(n, nnz, D, 0) == nnz * 8 + (n + 1) * 8 + 2 * n * D * 4
    assert bench.spmm_bytes(n, nnz, D, 2) - bench.spmm_bytes(n, nnz, D, 0) == 2 * n * D * 4


@pytest.mark.parametrize('name', ['gowalla-igcn', 'yelp-lightgcn'])
def test_cpu_baseline_leg_on_a_small_graph(name, monkeypatch):
    """cpu_block (the oracle port: full train steps + a bounded evaluation sample) on the 3000 x 4000 shape standing in
    for the workload's own: finite numbers, the documented keys, an honest sample description."""
    shape, kind, l2_reg, dropout = bench.WORKLOADS[name]
    monkeypatch.setitem(bench.WORKLOADS, name, ('small', kind, l2_reg, dropout))
    ds = bench.build_dataset('small', 'cpu')
    cfg = bench.workload_config(name, ds)
    assert cfg['workload'] == name and cfg['n_users'] == 3000 and cfg['n_items'] == 4000 and cfg['batch'] == 2048
    assert cfg['steps_per_epoch'] == -(-len(ds) // 2048)
    cpu = bench.cpu_block(ds, name, steps=1, warmup=0, eval_batches=1)
    assert cpu['kind'] == 'port' and cpu['unit'] == 'ms' and cpu['cores'] >= 1
    assert 0 < cpu['ms_per_step'] < 60e3 and abs(cpu['value'] - cpu['ms_per_step'] * cfg['steps_per_epoch']) < 1e-6 * cpu['value']
    assert cpu['eval_users_per_s'] > 0 and 'oracle/restate.py' in cpu['sample']


def test_reference_arm_prints_one_contract_line(tmp_path):
    """`bench.py --impl reference` on rank != 0 prints nothing and exits 0 (the driver launches it under torchrun too)."""
    p = subprocess.run([sys.executable, 'bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0'], cwd=ROOT,
                       env={**__import__('os').environ, 'RANK': '1', 'WORLD_SIZE': '2'}, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, timeout=300)
    assert p.returncode == 0 and p.stdout.decode().strip() == ''
