"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's hot path.

Nothing under oracle/ is imported by the product package (comorag_b200/); only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may use it, and only as the checker or the timed CPU baseline.
"""
