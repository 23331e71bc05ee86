"""Empty stub (oracle harness only)."""
