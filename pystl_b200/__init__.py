"""pystl_b200 -- B200-native (sm_100a) backend for PySTL's robustness evaluation.

    import pystl_b200                      # registers backend="b200" when `pystl` is installed
    from pystl import create_semantics     # ... or: from pystl_b200 import create_semantics
    sem = create_semantics("smooth", backend="b200", temperature=0.5)
    rho = sem.robustness_trace(phi, signal)     # signal [B, T, D] on the GPU -> rho [B, T]

See DESIGN.md for the kernel design and INTEGRATION.md for the C ABI.
"""

from .formula import (
    Always,
    And,
    Eventually,
    Formula,
    Interval,
    Not,
    Or,
    Predicate,
    Until,
    affine_predicate,
    from_spec,
    gt,
    lt,
    BASELINE_SPECS,
    BASELINE_SHAPES,
)
from .registry import BACKEND, create_semantics, install
from .semantics import (
    B200AgmRobustSemantics,
    B200ClassicRobustSemantics,
    B200CumulativeRobustness,
    B200CumulativeSemantics,
    B200DgmsrSemantics,
    B200SmoothRobustSemantics,
)
from .ctstl import B200CtstlSemantics, CumulativeTime
from .distributed import shard_bounds

REGISTERED_IN_REFERENCE = install()

__all__ = [
    "Formula", "Predicate", "Not", "And", "Or", "Always", "Eventually", "Until", "Interval",
    "affine_predicate", "gt", "lt", "from_spec", "BASELINE_SPECS", "BASELINE_SHAPES",
    "create_semantics", "install", "BACKEND", "shard_bounds",
    "B200ClassicRobustSemantics", "B200SmoothRobustSemantics", "B200CumulativeSemantics",
    "B200CumulativeRobustness", "B200AgmRobustSemantics", "B200DgmsrSemantics",
    "B200CtstlSemantics", "CumulativeTime",
]
