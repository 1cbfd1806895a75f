"""prosstt_b200 - PROSSTT's simulation hot path on NVIDIA B200 (sm_100a).

Same Python API as the reference package `prosstt` (tree.Tree, simulation.*, sim_utils.*,
count_model.*, tree_utils.*); the numeric work runs in hand-written CUDA kernels behind a
C ABI (include/prosstt_b200.h, prosstt_b200/libprosstt_b200.so).  No CPU fallback.

    from prosstt_b200 import tree, simulation as sim, sim_utils as sut, count_model as cm

To run an unmodified reference script (`from prosstt import ...`):

    import prosstt_b200; prosstt_b200.install_as_prosstt()
"""
import sys

__version__ = "0.1.0"

from prosstt_b200 import _native  # noqa: F401  (binding only; the library loads lazily)
from prosstt_b200 import tree, tree_utils, count_model, sim_utils, simulation, sharding, stats  # noqa: F401


def install_as_prosstt():
    """Register this package under the name `prosstt` so that scripts written for the
    reference (e.g. examples/generate_simN.py) import it unchanged."""
    me = sys.modules[__name__]
    sys.modules["prosstt"] = me
    for name in ("tree", "tree_utils", "count_model", "sim_utils", "simulation"):
        sys.modules["prosstt." + name] = getattr(me, name)
    return me
