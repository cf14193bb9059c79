"""Test infrastructure: see the package docstring."""
