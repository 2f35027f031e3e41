"""Test stub of langchain (see tests/stubs/README.md)."""
