"""TEST INFRASTRUCTURE ONLY -- not part of the product.

`oracle/` holds (1) a tiny `dgl` shim that lets the unmodified reference under
/root/reference run on CPU in the build container (`oracle/shim`, `oracle/ref_loader.py`)
and (2) a CPU restatement of the reference's hot path (`oracle/restate.py`) that travels
to the GPU box, where /root/reference does not exist.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package.  The product (`igcn_cf_b200/`) never does: it fails loudly
when the CUDA library is missing.

Parity pin: the reference ships no tests or golden vectors, so the pins are
outputs of the reference itself executed here through the shim; the generating script is
`tests/golden/make_golden.py` and the vectors live in `tests/golden/*.npz`.
`tests/test_oracle_golden.py` checks the restatement against them.
"""
