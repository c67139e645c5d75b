"""Import shim, see matplotlib/__init__.py."""
