"""Top stall-sample source lines of an ncu report:  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv;
python profiles/ncu_source_top.py src.csv [N]"""
import csv
import sys


def num(x):
    try:
        return int(float(x.replace(",", "")))
    except Exception:
        return 0


def main(path, n=25):
    rows = list(csv.reader(open(path)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "File Path":
            cur = {"file": r[1], "rows": []}
            blocks.append(cur)
        elif r and r[0] == "Line No" and cur is not None:
            cur["hdr"] = r
        elif cur is not None and "hdr" in cur and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    for b in blocks:
        h = b["hdr"]
        si, ie = h.index("# Samples"), h.index("Instructions Executed")
        tot = sum(num(r[si]) for r in b["rows"])
        if tot == 0:
            continue
        print("%s : %d sampled rows, %d samples" % (b["file"], len(b["rows"]), tot))
        for r in sorted(b["rows"], key=lambda r: -num(r[si]))[:n]:
            print("  line %5s  samples %6d (%4.1f%%)  warp-instr %9d | %s" % (r[0], num(r[si]), 100.0 * num(r[si]) / tot, num(r[ie]), r[1].strip()[:105]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
