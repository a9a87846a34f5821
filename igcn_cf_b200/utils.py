"""Small host utilities with the reference's names (reference utils.py:12-29, 126-151)."""
import os
import random
import sys

import numpy as np
import torch


def set_seed(seed=0):
    """Seed python / numpy / torch (CPU and CUDA) like utils.py:12-20."""
    random.seed(seed)
    os.environ['PYTHONHASHSEED'] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True


class Unbuffered:
    """Flush-on-write stream wrapper (utils.py:138-151)."""

    def __init__(self, stream):
        self.stream = stream

    def write(self, data):
        self.stream.write(data)
        self.stream.flush()

    def writelines(self, datas):
        self.stream.writelines(datas)
        self.stream.flush()

    def __getattr__(self, attr):
        return getattr(self.stream, attr)


def init_run(log_path, seed):
    """Seed and redirect stdout/stderr to <log_path>/log.txt (utils.py:23-29)."""
    set_seed(seed)
    os.makedirs(log_path, exist_ok=True)
    f = Unbuffered(open(os.path.join(log_path, 'log.txt'), 'w'))
    sys.stderr = f
    sys.stdout = f


class AverageMeter:
    """Running weighted mean (utils.py:126-135)."""

    def __init__(self):
        self.avg = 0.
        self.sum = 0.
        self.count = 0.

    def update(self, val, n=1):
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count
