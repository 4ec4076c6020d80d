"""CPU oracle package — TEST INFRASTRUCTURE ONLY (see oracle/mpopis_oracle.h)."""
