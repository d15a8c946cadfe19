"""Stand-in for the `diffusers` package: TEST INFRASTRUCTURE ONLY (see ../README.md)."""
__version__ = "0.36.0+shim"
