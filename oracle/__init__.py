"""CPU checker of the B200 path: test infrastructure only (see oracle/README.md)."""
