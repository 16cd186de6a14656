"""Oracle-only stand-in for the un-vendored espnet==202402 dependency (test infrastructure)."""
