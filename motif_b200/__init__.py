"""motif_b200 -- Blackwell (sm_100a) implementation of MoTIF's per-pixel inference hot path.

Drop-in surfaces (same names, arguments and error behaviour as the reference):

* ``motif_b200.softsplat_cp``        ``FunctionSoftsplat`` / ``Softsplat``          (models/softsplat_cp.py)
* ``motif_b200.softsplat_max_cp``    ``FunctionSoftsplat`` / ``Softsplat_Max``      (models/softsplat_max_cp.py)
* ``motif_b200.softsplat_count_cp``  ``FunctionSoftsplat`` / ``Softsplat_Count``    (models/softsplat_count_cp.py)
* ``motif_b200.correlation``         ``FunctionCorrelation`` / ``ModuleCorrelation`` (OpticalFlow/correlation.py)
* ``motif_b200.decoder``             ``SpaceTimeDecoder`` -- Ours.py:659-858 from resident LR latents
* ``motif_b200.luna_tokis``          ``install(model)`` -- binds the decoder into a reference ``LunaTokis``

All compute goes through ``libmotif_b200.so`` (``include/motif_b200.h``); there is no fallback.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
