"""CPU oracle for the smoke step -- TEST INFRASTRUCTURE ONLY (see smoke_oracle.c).

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only.
"""
from .oracle import *  # noqa: F401,F403
