"""CPU oracle (test infrastructure only - see tmk_oracle.c header)."""
