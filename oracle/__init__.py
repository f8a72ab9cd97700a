"""CPU oracle package (test infrastructure only; see similarity_oracle.py header)."""
