"""BackboneRegistry (/root/reference/flowmse/backbones/shared.py:10)."""
from ..util.registry import Registry

BackboneRegistry = Registry("Backbone")
