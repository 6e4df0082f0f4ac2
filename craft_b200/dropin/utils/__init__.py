"""Drop-in shim of the reference's `utils` package: `utils.utils` is craft_b200's; every other submodule
(flow_viz, frame_utils, augmentor) is looked up in the `utils` directories found further down sys.path."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
for _d in sys.path:
    _cand = os.path.join(os.path.abspath(_d or "."), "utils")
    if os.path.isdir(_cand) and os.path.abspath(_cand) != _here and _cand not in __path__:
        __path__.append(_cand)
