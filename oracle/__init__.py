"""CPU oracle for the KKT factor/solve path (test infrastructure, NOT the product).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  See ``oracle/ldl_oracle.c``.
"""
from .ldl_oracle import LDLFactStruct, amd_order, build_oracle, load_oracle  # noqa: F401
