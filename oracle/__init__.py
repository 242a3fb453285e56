"""TEST INFRASTRUCTURE ONLY - CPU oracles for the join path (see oracle/gcre_oracle.c, oracle/ref_shim.cpp).

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only.
"""
