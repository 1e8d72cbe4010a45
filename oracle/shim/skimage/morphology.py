"""Import shim (TEST INFRASTRUCTURE ONLY): see skimage/__init__.py."""
