"""CPU oracle for the RVQ / GRVQ quantize-codec path.  TEST INFRASTRUCTURE ONLY.

This package restates, operation for operation, the arithmetic of the reference
(jacquelm/AcademiCodec) on CPU tensors.  It exists so that the CUDA path can be checked
against something that runs anywhere.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it, and only as the checker
or the timed CPU baseline -- never as a fallback for the product path.
`academicodec_b200/` must not import anything from here (tests/test_cpu_host.py::test_package_never_imports_oracle).

Why torch-CPU and not C: the reference is 100 % Python/PyTorch and its arithmetic *is* a
sequence of ATen calls (`@`, `max`, `argmin`, `F.embedding`, `one_hot`, `mse_loss`); the most
faithful restatement issues the same calls in the same order, so the fp32 rounding matches
the reference bit for bit on the same host.  `adjudicate.py` adds an fp64 numpy arbiter.

Parity pinning: the reference ships no golden vectors.  The oracle is pinned
against the *live* reference imported in the dev container: `tests/golden/make_golden.py`
runs the unmodified reference modules on seeded inputs and commits their outputs under
`tests/golden/`; `tests/test_oracle_golden.py` requires the oracle to reproduce them exactly.
"""
