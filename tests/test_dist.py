"""Multi-rank tests: world_size-2 gloo on CPU for the host-side sharding logic, and (GPU box with at
least two devices) the row-sharded propagation / user-sharded evaluation against the single-GPU path."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _launch(mode, nproc, timeout):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(nproc),
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()),
           os.path.join(ROOT, 'tests', 'dist_worker.py'), mode]
    p = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout)
    out = p.stdout.decode()
    assert p.returncode == 0 and 'dist_worker: all ok' in out, out[-4000:]
    return out


def test_host_sharding_logic_world2_gloo():
    _launch('cpu', 2, 300)


def test_shard_bounds_and_split_single_process():
    import numpy as np
    from igcn_cf_b200 import dist as idist
    rowptr = np.concatenate([[0], np.cumsum(np.array([5, 0, 0, 100, 3, 3, 3, 50, 1, 1]))])
    for world in (1, 2, 3, 8):
        b = idist.shard_bounds(rowptr, world)
        assert b[0] == 0 and b[-1] == 10 and len(b) == world + 1 and np.all(np.diff(b) >= 0)
    assert [idist.split_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert idist.split_range(0, 0, 2) == (0, 0)


@pytest.mark.gpu
def test_row_sharded_path_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs at least two GPUs on the box')
    _launch('gpu', 2, 600)
