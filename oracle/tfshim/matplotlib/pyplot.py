"""Empty stand-in (see matplotlib/__init__.py)."""
