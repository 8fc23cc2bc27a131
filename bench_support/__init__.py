"""Callers and synthetic inputs for bench.py (not part of the product package)."""
