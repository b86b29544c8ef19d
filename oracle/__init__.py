"""CPU oracle for the clustering hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference)
may import this package; nothing under pacybara_b200/ does.  See pcb_oracle.c.
"""
