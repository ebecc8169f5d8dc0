"""TEST INFRASTRUCTURE ONLY — CPU oracle for the phoebus GRMHD hot path (see oracle/oracle.h)."""
