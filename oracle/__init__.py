"""CPU oracle for the RM2 hot path -- TEST INFRASTRUCTURE ONLY (see oracle/rm2_oracle.h)."""
