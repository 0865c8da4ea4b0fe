"""CPU oracle (test infrastructure only).  See oracle/linemod_oracle.cpp header."""
