"""``create_semantics(syntax, backend=...)`` with the ``b200`` backend plugged in.

The reference's dispatch is ``pystl.semantics.create_semantics``
(pystl/semantics/__init__.py:108-125): it normalises the names through two alias
tables (``:76-91``), gates on the backend alias BEFORE the registry is consulted
(``:101-105``) and then calls ``registry.create`` (pystl/semantics/registry.py:36-52).
So a new backend needs both a ``registry.register(...)`` per syntax and an entry in
``_BACKEND_ALIASES`` -- :func:`install` does exactly that when the reference package
is importable.  On a machine without ``pystl`` (the GPU box) the same dispatch is
provided here, restricted to the one backend this package implements, with the
reference's error behaviour (KeyError for unknown syntax / backend, TypeError for
unknown kwargs).
"""

from __future__ import annotations

from typing import Any, Callable, Dict, Tuple

from .semantics import (
    B200AgmRobustSemantics,
    B200ClassicRobustSemantics,
    B200CumulativeSemantics,
    B200DgmsrSemantics,
    B200SmoothRobustSemantics,
)

BACKEND = "b200"
BACKEND_ALIASES = {"b200": BACKEND, "cuda": BACKEND, "sm_100a": BACKEND, "blackwell": BACKEND}
SYNTAX_ALIASES = {  # pystl/semantics/__init__.py:76-83
    "classical": "classical",
    "smooth": "smooth",
    "cumulative": "cumulative",
    "dgmsr": "dgmsr",
    "agm": "agm",
    "arithmetic_geometric_mean": "agm",
}
FACTORIES: Dict[str, Callable[..., Any]] = {
    "classical": B200ClassicRobustSemantics,
    "smooth": B200SmoothRobustSemantics,
    "cumulative": B200CumulativeSemantics,
    "dgmsr": B200DgmsrSemantics,
    "agm": B200AgmRobustSemantics,
}


def _norm(value: str) -> str:
    return value.strip().lower().replace("-", "_")


def create_semantics(syntax: str, *, backend: str = BACKEND, **kwargs: Any):
    """Same call as the reference's; ``backend`` defaults to (and must be an alias of) ``"b200"``
    unless the reference is installed, in which case its own backends resolve too."""
    s = _norm(syntax)
    if s not in SYNTAX_ALIASES:
        raise KeyError(f"Unknown syntax '{syntax}'. Available: {sorted(set(SYNTAX_ALIASES.values()))}")
    b = _norm(backend)
    if b in BACKEND_ALIASES:
        return FACTORIES[SYNTAX_ALIASES[s]](**kwargs)
    try:
        import pystl.semantics as ref  # type: ignore
    except ImportError:
        raise KeyError(f"Unknown backend '{backend}'. Available: ['{BACKEND}']") from None
    return ref.create_semantics(syntax, backend=backend, **kwargs)


def install() -> bool:
    """Register the backend in the reference's registry and alias table (idempotent).

    Returns False when ``pystl`` is not importable (nothing to plug into)."""
    try:
        import pystl.semantics as ref  # type: ignore
    except ImportError:
        return False
    for syntax, factory in FACTORIES.items():
        ref.registry.register(syntax=syntax, backend=BACKEND, factory=factory)
    for alias, target in BACKEND_ALIASES.items():
        ref._BACKEND_ALIASES.setdefault(alias, target)  # the gate at semantics/__init__.py:101-105
    return True
