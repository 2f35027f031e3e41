"""Test stub of langchain_core (see tests/stubs/README.md)."""
