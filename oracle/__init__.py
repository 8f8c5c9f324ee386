"""CPU oracle for the Quantik hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package.  The
product package (``quantik_b200``) never does.
"""
from .oracle import *  # noqa: F401,F403
