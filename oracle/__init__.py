"""CPU oracle package — TEST INFRASTRUCTURE ONLY (see oracle/ht_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this.  The product package `headtrackr_b200` never does.
"""
from .binding import *  # noqa: F401,F403
