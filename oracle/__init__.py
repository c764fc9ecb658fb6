"""CPU oracle (test infrastructure only -- see oracle/oracle.py).  Never imported by flowreg3d_b200."""
