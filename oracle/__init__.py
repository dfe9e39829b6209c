"""CPU oracle for the e2e_asr hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under `e2e_asr_b200/` imports this package.  Allowed importers: `tests/`,
`__graft_entry__.smoke()`, and `bench.py`'s `cpu_baseline` / `--impl reference`
legs (as the checker / the timed CPU reference, never as the product path).
See oracle/model.py for the parity-pinning statement.
"""
