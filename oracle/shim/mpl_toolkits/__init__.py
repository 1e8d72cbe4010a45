"""Import shim (TEST INFRASTRUCTURE ONLY): see matplotlib/__init__.py."""
