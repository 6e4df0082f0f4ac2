"""Summarise a CRAFT_GEMM_TRACE file (csrc/capi.cu): per launch, the phase marks of CTA (0,0) in clocks and the
spread of CTA start / end times (globaltimer, ns) over the grid.
usage: python profiles/gemm_trace_report.py gpurun_out/<trace>.txt"""
import sys

NAMES = ["entry", "pre-wait", "pdl_wait done", "setup sync", "1st TMA issued", "last TMA issued", "1st stage full",
         "last MMA issued", "epi pre-work done", "acc complete", "epilogue done", "teardown sync"]


def main(path):
    recs, cur = [], None
    for ln in open(path):
        t = ln.split()
        if not t:
            continue
        if t[0] == "gemm":
            cur = dict(head=ln.strip(), marks=None, ctas=[])
            recs.append(cur)
        elif t[0] == "marks":
            cur["marks"] = [int(x) for x in t[1:]]
        elif t[0] == "cta":
            cur["ctas"].append(tuple(int(x) for x in t[1:]))
    seen = {}
    for r in recs:          # keep the LAST record of every distinct launch shape (warm caches)
        seen[r["head"]] = r
    for head, r in seen.items():
        print(head)
        m = r["marks"]
        print("  " + "  ".join("%s=%d" % (n, v) for n, v in zip(NAMES, m) if v >= 0))
        st = [c[1] for c in r["ctas"] if c[1]]
        en = [c[2] for c in r["ctas"] if c[2]]
        if st and en:
            t0 = min(st)
            dur = sorted(e - s for (_, s, e, _) in r["ctas"] if s and e)
            print("  grid: %d CTAs on %d SMs; start spread %d ns; first end %d, last end %d ns after first start; "
                  "CTA duration min/median/max %d/%d/%d ns" % (len(st), len(set(c[3] for c in r["ctas"])), max(st) - t0,
                                                               min(en) - t0, max(en) - t0, dur[0], dur[len(dur) // 2], dur[-1]))


if __name__ == "__main__":
    main(sys.argv[1])
