"""Stall-sample breakdown from `ncu -i rep --page source --csv` output (one kernel)."""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    si, src = hdr.index("# Samples"), hdr.index("Source")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[hi + 1:]:
        if r and r[0] == "Kernel Name":
            break  # next kernel of the report
        if len(r) == len(hdr) and r[si].isdigit():
            data.append(r)
    tot = sum(int(r[si]) for r in data)
    print("total samples", tot, "instructions", len(data))
    for i in stall_cols:
        s = sum(int(r[i]) for r in data)
        if s > tot * 0.01:
            print(f"{hdr[i]:25s} {s:8d} {100 * s / tot:5.1f}%")
    print("top instructions by samples")
    for idx, r in sorted(enumerate(data), key=lambda x: -int(x[1][si]))[:top]:
        reasons = sorted(((int(r[i]), hdr[i]) for i in stall_cols), reverse=True)[:2]
        print(f"{idx:5d} {int(r[si]):6d} {100 * int(r[si]) / tot:5.2f}% {r[src].strip()[:64]:64s} {reasons}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
