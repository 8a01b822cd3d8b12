"""stub: the reference imports matplotlib at module scope (utils.py:2) but never uses it on the hot path."""
