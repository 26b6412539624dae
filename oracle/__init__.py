"""TEST INFRASTRUCTURE ONLY.

CPU restatement ("oracle") of the VSRD silhouette-rendering hot path.  Nothing in the
product package (`vsrd_b200/`, `vsrd/`) may import from here; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs do.

Parity status: PINNED.  The reference ships no tests or golden vectors, so the
restatement is pinned against outputs of the *unmodified reference modules* imported in the
build container (`tests/golden/make_golden.py`, fixtures under `tests/golden/*.npz`) and
re-checked by `tests/test_oracle_golden.py`.
"""
