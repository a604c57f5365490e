"""ORACLE package — test infrastructure only (see oracle/README.md).

Nothing under omni3d_b200/ may import this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
"""
