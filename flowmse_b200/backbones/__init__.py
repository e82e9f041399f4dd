"""Backbone registry + the B200 NCSN++ (mirror of /root/reference/flowmse/backbones/__init__.py, shared.py:10)."""
from .shared import BackboneRegistry
from .ncsnpp import NCSNpp

__all__ = ["BackboneRegistry", "NCSNpp"]
