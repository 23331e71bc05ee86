"""TEST INFRASTRUCTURE: import stub (the reference's 3D path generator imports the 3D viewer toolkits at module scope)."""
