"""CPU oracle for the rANS / CDF / coupling hot path.  TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
