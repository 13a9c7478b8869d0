"""TEST INFRASTRUCTURE ONLY: CPU oracle of the pastml ACR hot path (see oracle/pastml_oracle.c, cpu_oracle.py)."""
