"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle/gcb_oracle.h)."""
