"""CPU oracle for the MBAR hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import anything from this package.  The product package
(pymbar_b200) never imports it and has no CPU fallback.
"""
