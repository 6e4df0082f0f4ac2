"""Drop-in shim of core/utils/utils.py: InputPadder, coords_grid, upflow8, forward_interpolate, print0 come
from craft_b200; the rest (bilinear_sampler, ...) falls through to the reference's file when present."""
from craft_b200.utils.utils import *          # noqa: F401,F403
from craft_b200.utils import utils as _impl
from _fallthrough import extend as _extend

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
_extend("utils.utils", globals(), __file__)
