"""Minimal stand-in for the parts of PyG the reference's hot-path modules import
(TEST INFRASTRUCTURE ONLY — see oracle/ref_stub/README.md)."""
from . import data, loader, transforms, utils  # noqa: F401
