"""igcn_cf_b200 -- B200 (sm_100a) implementation of the INMO / IGCN hot path.

`model`, `trainer`, `dataset`, `utils`, `config` mirror the reference's top-level modules of the
same names; `dropin/` at the repository root re-exports them under those top-level names.
Compute lives in libigcn_b200.so (igcn_cf_b200/csrc, C ABI in include/igcn_b200.h); importing this
package does not need a GPU, calling a model does (there is no CPU fallback).
"""
__version__ = '0.1.0'
