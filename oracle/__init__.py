"""Test infrastructure only: CPU restatement of the reference algorithms (checker, never the product path)."""
