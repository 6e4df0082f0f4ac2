"""Helper of the drop-in shim directory (see README in __init__ of this directory's modules).

The reference's drivers do `sys.path.append('core'); from network import CRAFT` (evaluate.py:1-2,14;
train.py:2-3,20).  Putting THIS directory on sys.path ahead of the reference's core/ makes those
unmodified imports resolve to craft_b200.  Drivers also import names the hot path does not replace
(`from update import BasicUpdateBlock` in raft.py, `from extractor import SmallEncoder`, ...): every
public name a shim module lacks is taken from the same-named module found FURTHER DOWN sys.path (the
reference's own core/, when the driver put it there), so those imports keep working too."""
import importlib.util
import os
import sys


def extend(module_name, namespace, shim_file):
    here = os.path.dirname(os.path.abspath(shim_file))
    rel = os.path.join(*module_name.split(".")) + ".py"
    # walk up from the shim file to the shim root (utils/utils.py sits one level down)
    root = here
    for _ in range(module_name.count(".")):
        root = os.path.dirname(root)
    for d in sys.path:
        d_abs = os.path.abspath(d or ".")
        if d_abs == root:
            continue
        cand = os.path.join(d_abs, rel)
        if not os.path.isfile(cand):
            continue
        spec = importlib.util.spec_from_file_location("_craft_ref_" + module_name.replace(".", "_"), cand)
        mod = importlib.util.module_from_spec(spec)
        try:
            spec.loader.exec_module(mod)
        except Exception as e:        # a missing optional dependency of the reference must not break the shim
            namespace.setdefault("__fallthrough_error__", repr(e))
            return None
        for k, v in vars(mod).items():
            if not k.startswith("_") and k not in namespace:
                namespace[k] = v
        namespace["__fallthrough__"] = cand
        return mod
    return None
