"""TEST INFRASTRUCTURE ONLY: CPU restatement of libepic's log-space harmonic path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (epic_b200) never does.
"""
