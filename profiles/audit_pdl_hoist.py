"""Audit: no global load may be scheduled before griddepcontrol.wait (SASS: ACQBULK) in any kernel of the
library.  nvcc treats loads through `const T* __restrict__` / __ldg as invariant and is free to hoist
them above the wait -- they then read the predecessor's buffers before it has finished (this is how the
clamp gate read a stale score maximum under CUDA-graph replay in round 2).
usage: python profiles/audit_pdl_hoist.py [lib.so]   -> exit status 1 when an offending load exists"""
import re
import subprocess
import sys


def audit(lib):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    bad, name, seen_wait, pre = {}, None, False, []
    for ln in sass.splitlines():
        if "Function :" in ln:
            name, seen_wait, pre = ln.split("Function :")[1].strip(), False, []
            continue
        if name is None:
            continue
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if not m:
            continue
        ins = m.group(2).strip()
        if "ACQBULK" in ins:
            if pre and not seen_wait:
                bad[name] = list(pre)
            seen_wait = True
        elif not seen_wait and re.search(r"\b(LDG|LD\.E|ATOMG|REDG|RED\.E|UTMALDG|STG)\b", ins):
            pre.append(m.group(1) + ": " + ins)
    return bad


if __name__ == "__main__":
    lib = sys.argv[1] if len(sys.argv) > 1 else "craft_b200/libcraft_b200.so"
    bad = audit(lib)
    for k, v in bad.items():
        print(k)
        for x in v:
            print("    ", x)
    print("%d kernel(s) with global memory traffic before griddepcontrol.wait" % len(bad))
    sys.exit(1 if bad else 0)
