"""CPU oracle for the block-simplex hot path -- TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
The product package never imports this module (tests/test_boundary.py greps for it).
"""
