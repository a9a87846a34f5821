"""Import the UNMODIFIED reference (/root/reference) on CPU.  TEST INFRASTRUCTURE ONLY.

Works only where /root/reference exists (the build container).  It is used by
`tests/golden/make_golden.py` to produce the committed golden vectors and by optional
container-only cross-checks; nothing that runs on the GPU box calls it.

The reference dispatches classes by name through sys.modules['model'|'trainer'|'dataset']
(/root/reference/model.py:19, trainer.py:18, dataset.py:12), so its files must be imported
under exactly those top-level names.
"""
import importlib
import os
import sys

REFERENCE_ROOT = '/root/reference'
_NAMES = ('utils', 'dataset', 'model', 'trainer', 'config')


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'model.py'))


def load():
    """Return dict name -> reference module (utils, dataset, model, trainer, config)."""
    if not available():
        raise RuntimeError('reference tree not present at ' + REFERENCE_ROOT)
    shim = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shim')
    for p in (REFERENCE_ROOT, shim):
        if p in sys.path:
            sys.path.remove(p)
    sys.path[:0] = [shim, REFERENCE_ROOT]
    for n in _NAMES:
        mod = sys.modules.get(n)
        if mod is not None and not getattr(mod, '__file__', '').startswith(REFERENCE_ROOT):
            raise RuntimeError('a different top-level module named %r is already imported' % n)
    return {n: importlib.import_module(n) for n in _NAMES}
