"""Top-level alias so reference launchers (`from dataset import ...`) pick up the B200 implementation."""
from igcn_cf_b200.dataset import *  # noqa: F401,F403
from igcn_cf_b200 import dataset as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
