"""CPU oracle package -- test infrastructure only (see apgd_oracle.py header)."""
