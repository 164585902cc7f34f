"""CPU oracle for the deep_cartograph CV hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the
product package ``deep_cartograph_b200``; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker.
"""
from .cv_oracle import *  # noqa: F401,F403
