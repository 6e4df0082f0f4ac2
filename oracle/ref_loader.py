"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (craft_b200/).

Loads the *unmodified* reference (askerlee/craft) from /root/reference so that
tests/ and oracle/make_golden.py can execute it as the parity oracle.  Only
usable in the build container: /root/reference does not exist on the GPU box,
which is why its outputs are frozen into tests/golden/ by make_golden.py.

Reference entry points exercised:
  core/network.py:26   CRAFT(args)
  core/network.py:164  CRAFT.forward(image1, image2, iters, flow_init, upsample, test_mode)
"""
import argparse
import contextlib
import io
import os
import sys
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
_STAGED = os.path.join(_HERE, "_ref", "reference")      # untracked copy made by __graft_entry__.build() for the GPU box


def _pick_root():
    env = os.environ.get("CRAFT_REFERENCE_ROOT")
    for cand in (env, "/root/reference", _STAGED):
        if cand and os.path.isfile(os.path.join(cand, "core", "network.py")):
            return cand
    return env or "/root/reference"


REF_ROOT = _pick_root()
REF_CORE = os.path.join(REF_ROOT, "core")
LOCAL_WEIGHTS = os.path.join(os.path.dirname(_HERE), "tests", "golden", "_local", "craft-sintel-model.pth")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_CORE, "network.py"))


from craft_b200.testing import craft_args, smooth_pair, synthetic_pair  # noqa: E402,F401  (shared input helpers)


@contextlib.contextmanager
def _ref_on_path():
    """Temporarily put the reference's core/ first on sys.path and hide any
    same-named modules of ours (network, corr, setrans, ...)."""
    names = ["network", "corr", "setrans", "setrans_ablation", "gma", "update",
             "extractor", "utils", "utils.utils", "raft"]
    saved = {n: sys.modules.pop(n) for n in names if n in sys.modules}
    sys.path.insert(0, REF_CORE)
    try:
        yield
    finally:
        sys.path.remove(REF_CORE)
        for n in names:
            sys.modules.pop(n, None)
        sys.modules.update(saved)


def load_checkpoint_tensors(path):
    """torch.load of a reference checkpoint through a RESTRICTED unpickler: the files under /root/reference
    are untrusted public content and the checkpoints are full training states (optimizer, scheduler, the
    trainer's Logger) in a pickle dialect torch's weights_only loader cannot parse.  Only classes from
    torch / numpy / collections / argparse (plus builtins.getattr, which the file uses to rebuild dtypes) can
    be instantiated; anything else -- e.g. __main__.Logger -- becomes an inert placeholder, os/subprocess/...
    raise."""
    import pickle
    import torch

    class _Inert:
        def __init__(self, *a, **k):
            pass

        def __setstate__(self, st):
            pass

    ok_roots = ("torch", "numpy", "collections", "argparse", "_codecs")

    class _Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            root = module.split(".")[0]
            if module == "__builtin__":          # python-2 era protocol
                module = "builtins"
            if root in ok_roots or (module, name) in (("builtins", "getattr"), ("builtins", "set"), ("builtins", "dict"),
                                                      ("builtins", "list"), ("builtins", "tuple"), ("builtins", "int"),
                                                      ("builtins", "float"), ("builtins", "complex"), ("builtins", "slice")):
                return super().find_class(module, name)
            if root in ("__main__", "train", "train_ddp", "evaluate"):      # the trainer's own bookkeeping classes
                return _Inert
            raise pickle.UnpicklingError("refusing to unpickle %s.%s from an untrusted checkpoint" % (module, name))

    class _Pickle:
        __name__ = "pickle"
        Unpickler = _Unpickler

        @staticmethod
        def load(f, **kw):
            return _Unpickler(f, **kw).load()
    return torch.load(path, map_location="cpu", weights_only=False, pickle_module=_Pickle)


def load_reference_modules():
    """Returns a dict of the reference's modules (network, corr, setrans, gma, update, utils)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    with _ref_on_path(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import network, corr, setrans, gma, update, extractor  # noqa
        import utils.utils as uu
        return dict(network=network, corr=corr, setrans=setrans, gma=gma,
                    update=update, extractor=extractor, utils=uu)


def build_reference_model(args=None, checkpoint="craft-sintel.pth", seed=1234, quiet=True):
    """CRAFT(args) from the reference with optional checkpoint, eval mode, CPU fp32."""
    import torch
    mods = load_reference_modules()
    args = args or craft_args()
    torch.manual_seed(seed)
    sink = io.StringIO()
    with warnings.catch_warnings(), contextlib.redirect_stdout(sink if quiet else sys.stdout):
        warnings.simplefilter("ignore")
        model = mods["network"].CRAFT(args)
    if checkpoint:
        path = os.path.join(REF_ROOT, "checkpoints", checkpoint)
        if os.path.isfile(path):
            ck = load_checkpoint_tensors(path)
            sd = ck["model"] if "model" in ck else ck
        elif checkpoint == "craft-sintel.pth" and os.path.isfile(LOCAL_WEIGHTS):
            sd = torch.load(LOCAL_WEIGHTS, map_location="cpu")       # the staged tree carries no checkpoints
        else:
            raise FileNotFoundError(path)
        sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in sd.items()}
        missing, unexpected = model.load_state_dict(sd, strict=False)
        model._load_report = (list(missing), list(unexpected))
    model.eval()
    return model, mods
