"""CPU oracle (test infrastructure only; never imported by dlux_b200)."""
