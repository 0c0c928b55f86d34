"""tetrad_b200 -- B200-native quartet-resolution hot path of eaton-lab/tetrad.

Only the per-quartet path of ``tetrad/src/resolve_quartets.py`` is implemented here, as
hand-written sm_100a CUDA behind a C ABI (``include/tetrad_b200.h``); see DESIGN.md.
"""
__version__ = "0.1.0"

from .resolve_quartets import (  # noqa: F401,E402
    QuartetResolver,
    infer_resolved_quartets,
    new_infer_resolved_quartets,
)
from .tetrad import Tetrad  # noqa: F401,E402
