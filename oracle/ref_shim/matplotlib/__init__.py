"""Test-infrastructure stub: lets the read-only reference import without matplotlib. Plotting is a no-op."""
