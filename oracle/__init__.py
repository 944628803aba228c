"""CPU oracle: test infrastructure only (see oracle/lda_oracle.py)."""
