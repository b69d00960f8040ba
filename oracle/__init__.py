"""CPU oracle + legacy-reference doors.  TEST INFRASTRUCTURE ONLY (see oracle/oracle.py)."""
