"""TEST INFRASTRUCTURE ONLY — CPU restatement (oracle) of MILLION's PQ KV-cache hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import anything from this package.  The product (`million_b200/`) never does.
"""
