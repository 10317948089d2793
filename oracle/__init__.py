"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the FitSNAP linear-fit hot path.

Nothing in `fitsnap_b200/` (the product) may import this package.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` use it, and only as the checker or as the timed CPU baseline.

Contents
--------
* `linear_fit.py`  -- numpy/scipy/sklearn restatement of the reference's row
  assembly (`_collect_lammps`) and solver prologue + third-party solve calls.
* `ref_driver.py`  -- drives the UNMODIFIED reference from `/root/reference` with a
  fake `lammps` module.  Works only in the build container (the reference tree does
  not exist on the GPU box); used to pin `linear_fit.py` and to generate
  `tests/golden/*` via `make_golden.py`.

Parity status: PINNED.  `linear_fit.py` is checked (tests/test_oracle.py) against
(1) the reference's own golden triple `Descriptors/Truth-Ref/Weights.npy ->
Ta_pot.snapcoeff`, and (2) fixtures produced by running the unmodified reference
classes in this container (`tests/golden/*.npz`, generator `oracle/make_golden.py`).
RIDGE / LASSO / apply_transpose have no golden in the reference's own test-suite
(SURVEY 8c); they are pinned only by (2).
"""
