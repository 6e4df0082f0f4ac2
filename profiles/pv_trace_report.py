"""Pretty-print the CRAFT_PV_TRACE timeline of attn_pv CTA (0,0,0): clock64 deltas per tile and role.
usage: python profiles/pv_trace_report.py gpurun_out/pv_trace.txt"""
import sys
allrows = [list(map(int, l.split())) for l in open(sys.argv[1])]
rows = [r for r in allrows if r[0] < 4]
spans = [r for r in allrows if r[0] == 9]
t0 = min(v for r in rows for v in r[2:] if v > 0)
names = {0: "MMA  [S(j) issued by S-warp, -, PV-warp: p_full, v_full, PV issued]",
         1: "SM g0 [loop top, s_full acquired, S loaded, P computed, P stored + arrived, o_full wait, o_full acquired, O written]",
         2: "SM g1 [loop top, s_full acquired, S loaded, P computed, P stored + arrived, o_full wait, o_full acquired, O written]",
         3: "TMA  [K issued, V issued]"}
for role in range(4):
    print(names[role])
    for r in rows:
        if r[0] != role or not any(r[2:]):
            continue
        print("  tile %2d: " % r[1] + " ".join("%7d" % (v - t0) if v else "      -" for v in r[2:10]))
if spans:
    s0 = min(r[2] for r in spans)
    ends = sorted(r[3] - s0 for r in spans)
    starts = sorted(r[2] - s0 for r in spans)
    print("CTA spans (ns, relative to the first start): starts %d..%d, ends min %d / median %d / max %d" % (
        starts[0], starts[-1], ends[0], ends[len(ends) // 2], ends[-1]))
