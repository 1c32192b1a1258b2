"""CPU oracle for the pygenray hot path -- TEST INFRASTRUCTURE ONLY (see prt_oracle.c)."""
