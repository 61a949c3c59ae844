"""CPU oracle (test infrastructure only; see oracle/vip_oracle.py)."""
