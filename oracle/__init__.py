"""CPU oracle for the ray-casting hot path.  TEST INFRASTRUCTURE ONLY.

Importable from ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline
legs of ``bench.py`` -- never from ``pyqsm_b200`` (the product).  PARITY
UNPINNED: see the header of ``qsmrt_oracle.c``.
"""
from .oracle import OracleScene, build_oracle, INVALID_ID  # noqa: F401
