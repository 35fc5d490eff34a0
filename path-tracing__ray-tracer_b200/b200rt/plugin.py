"""Renderer plug-in interface — the drop-in boundary.

Mirrors reference ``renderers/base_renderer.py:7-51``: an abstract
``BaseRenderer`` (``render(scene, camera, settings) -> PIL.Image``,
``get_capabilities()``, ``get_name()``, ``supports()``) and a class-level
``RendererFactory`` registry (``register`` / ``create(name, **kwargs)`` raising
``ValueError`` on unknown names / ``list_available``).

When this package is imported from inside a reference checkout (``renderers``
is importable) the reference's own classes are re-exported instead, so the B200
renderers register themselves into the *reference's* factory and
``python main.py -r b200_path_tracer`` works unchanged (``main.py:26,75``).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Dict, List


def _reference_plugin():
    import sys
    before = set(sys.modules)
    try:
        from renderers.base_renderer import BaseRenderer as B, RendererFactory as F  # type: ignore
        if hasattr(F, "register") and hasattr(F, "create"):
            return B, F
    except Exception:
        pass
    # an unrelated top-level package that happens to be called ``renderers`` must not stay imported on our
    # account: it would shadow a reference checkout put on sys.path later
    for m in set(sys.modules) - before:
        if m == "renderers" or m.startswith("renderers."):
            sys.modules.pop(m, None)
    return None


_ref = _reference_plugin()

if _ref is not None:
    BaseRenderer, RendererFactory = _ref
else:
    class BaseRenderer(ABC):
        def __init__(self, name: str):
            self.name = name

        @abstractmethod
        def render(self, scene, camera, settings):
            """Render ``scene`` through ``camera``; returns an RGB ``PIL.Image`` (row 0 = top)."""

        @abstractmethod
        def get_capabilities(self) -> List[str]:
            """Feature strings this renderer supports."""

        def get_name(self) -> str:
            return self.name

        def supports(self, feature: str) -> bool:
            return feature in self.get_capabilities()

    class RendererFactory:
        _renderers: Dict[str, type] = {}

        @classmethod
        def register(cls, name: str, renderer_class) -> None:
            cls._renderers[name] = renderer_class

        @classmethod
        def create(cls, name: str, **kwargs) -> "BaseRenderer":
            if name not in cls._renderers:
                raise ValueError(f"Unknown renderer: {name}")
            return cls._renderers[name](**kwargs)

        @classmethod
        def list_available(cls) -> List[str]:
            return list(cls._renderers.keys())
