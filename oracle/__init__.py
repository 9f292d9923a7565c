"""CPU oracle -- test infrastructure only (see oracle/cf_oracle.c)."""
