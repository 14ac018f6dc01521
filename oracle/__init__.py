"""CPU oracle for the NPDR hot path -- TEST INFRASTRUCTURE ONLY (see npdr_oracle.c header)."""
