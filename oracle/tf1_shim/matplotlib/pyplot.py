"""stub (see package docstring)."""
