"""CPU oracle for the Yota gradient-tape hot path.  TEST INFRASTRUCTURE ONLY:
importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs; never from the product package."""
