"""TEST INFRASTRUCTURE ONLY. CPU restatement of the reference hot path (see unet_oracle.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package,
and only as the checker or the timed CPU baseline - never on the product path.
"""
