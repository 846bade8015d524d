"""TEST INFRASTRUCTURE ONLY: CPU oracle for nlmc_b200 (see oracle.py, nlmc_oracle.c)."""
