"""Import shim (test infrastructure only): see ../cupy/__init__.py."""
