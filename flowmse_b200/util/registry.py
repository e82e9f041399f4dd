"""Plugin slots of the drop-in: a mapping from a public name to the class implementing it.

Contract taken from /root/reference/flowmse/util/registry.py:5-34, which the reference's ODE solvers, ODEs and backbones
register themselves with: ``register(name)`` is a class decorator, a second registration under the same name replaces the
first with a warning, ``get_by_name`` raises ``ValueError`` for an unknown name, ``get_all_names`` lists what is there.
The implementation is a thin dict subclass so that registries can also be inspected like mappings (``name in reg``).
"""
from __future__ import annotations

import warnings
from typing import Dict, Iterable, Type


class Registry(Dict[str, Type]):
    def __init__(self, managed_thing: str):
        super().__init__()
        self.managed_thing = managed_thing          # noun used in messages: "ODE", "ODEsolver", "Backbone"

    def _complain(self, what: str, name: str) -> str:
        return f"{self.managed_thing} with name '{name}' {what}"

    def register(self, name: str):
        """``@registry.register("euler")`` above a class definition."""
        def bind(cls: Type) -> Type:
            if name in self:
                warnings.warn(self._complain("doubly registered, old class will be replaced.", name))
            self[name] = cls
            return cls
        return bind

    def get_by_name(self, name: str) -> Type:
        if name not in self:
            raise ValueError(self._complain("unknown.", name))
        return self[name]

    def get_all_names(self) -> Iterable[str]:
        return list(self)
