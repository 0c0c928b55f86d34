"""CPU oracle of the quartet-resolution hot path.  TEST INFRASTRUCTURE ONLY (see oracle.py)."""
