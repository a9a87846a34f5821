"""Top-level alias so reference launchers (`from utils import ...`) pick up the B200 implementation."""
from igcn_cf_b200.utils import *  # noqa: F401,F403
from igcn_cf_b200 import utils as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
