"""TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy / scipy-LAPACK) of the inference-tools GpRegressor hot path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
package; the product (`inference_tools_b200`) never does.
"""
