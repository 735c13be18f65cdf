"""Shim package (test infrastructure): see ../README.md."""
