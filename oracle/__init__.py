"""CPU oracle for the B200 path-tracing core — TEST INFRASTRUCTURE ONLY (see rt_oracle.c)."""
