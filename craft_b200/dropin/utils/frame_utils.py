"""Drop-in shim of core/utils/frame_utils.py: the .flo / KITTI-png writers and readers are craft_b200's
(craft_b200/utils/frame_utils.py); read_gen and the remaining readers fall through to the reference."""
from craft_b200.utils.frame_utils import *          # noqa: F401,F403
from craft_b200.utils import frame_utils as _impl
from _fallthrough import extend as _extend

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
_extend("utils.frame_utils", globals(), __file__)
