"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/cosmo_oracle.c header).  Never imported by the package."""
