"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see ea_oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this.  The product package edge_alignment_b200 never does.
"""
