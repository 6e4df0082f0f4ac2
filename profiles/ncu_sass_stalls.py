"""Per-instruction stall samples of an ncu report's SASS view:
   ncu -i X.ncu-rep --page source --csv --print-source sass > sass.csv; python profiles/ncu_sass_stalls.py sass.csv [N]
Prints the totals per stall reason, per opcode, and the N most-sampled instructions with their index."""
import csv
import sys
from collections import Counter


def main(path, n=40):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    body = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
    si = hdr.index("# Samples")
    ie = hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[si]) for r in body)
    print("instructions %d, samples %d, warp-instr executed %d" % (len(body), tot, sum(int(r[ie]) for r in body)))
    reasons = Counter()
    for r in body:
        for i in stall_cols:
            reasons[hdr[i]] += int(r[i])
    print("stall reasons:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in reasons.most_common(10)))
    ops = Counter()
    for r in body:
        ops[r[1].split()[0] if not r[1].split()[0].startswith("@") else r[1].split()[1]] += int(r[si])
    print("by opcode:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in ops.most_common(16)))
    print("top instructions (index, samples, share, executed, main stall, text):")
    for idx, r in sorted(enumerate(body), key=lambda kv: -int(kv[1][si]))[:n]:
        main_stall = max(stall_cols, key=lambda i: int(r[i]))
        print("  %5d %6d %5.1f%% %9d  %-22s %s" % (idx, int(r[si]), 100.0 * int(r[si]) / tot, int(r[ie]), hdr[main_stall], r[1].strip()[:70]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
