"""TEST INFRASTRUCTURE ONLY.

CPU oracle for the MoCoPCI point-set neighbourhood hot path (see oracle/oracle.h). Only tests/,
``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs may
import this package; nothing under ``mocopci_b200/`` does, and the product fails loudly when
its CUDA library is missing instead of falling back to anything here.
"""
