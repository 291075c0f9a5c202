"""CPU oracle for dpose_core's footprint-cost path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package; the product (dpose_b200/) never does.
"""
