"""Drop-in shim: `from corr import ...` as the reference's drivers write it resolves to craft_b200.corr
(same class names, constructor arguments, state-dict keys and forward() signatures); names the hot path
does not replace fall through to the reference's own core/corr.py when that is on sys.path too."""
from craft_b200.corr import *          # noqa: F401,F403
from craft_b200 import corr as _impl
from _fallthrough import extend as _extend

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
_extend("corr", globals(), __file__)
