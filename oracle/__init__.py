"""CPU oracle (test infrastructure only) -- see ref_search.py / oracle.c headers."""
