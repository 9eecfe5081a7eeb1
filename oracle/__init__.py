"""CPU oracle (TEST INFRASTRUCTURE). Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product package
lab_1806_vec_db_b200 never does."""
from .oracle_py import *  # noqa: F401,F403
