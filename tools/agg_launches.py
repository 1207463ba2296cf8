"""Aggregate an `ncu --csv` launch list: per kernel (name, grid, block) count, mean time, extra metrics."""
import collections
import csv
import sys


def main(path, top=30):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    agg = collections.OrderedDict()
    for r in rows:
        try:
            name, metric, val = r[4], r[-3], float(r[-1].replace(",", ""))
        except ValueError:
            continue
        short = name.split("(")[0][-70:]
        key = (short, r[7], r[8])
        d = agg.setdefault(key, collections.defaultdict(list))
        d[metric].append(val)
    tot = sum(sum(d["gpu__time_duration.sum"]) for d in agg.values())
    for k, d in sorted(agg.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"]))[:top]:
        t = d["gpu__time_duration.sum"]
        extra = ""
        if "dram__bytes_read.sum" in d:
            rd, wr = d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
            # ncu prints bytes in varying units in csv? keep raw means
            extra += f" rd={sum(rd)/len(rd):.4g} wr={sum(wr)/len(wr):.4g}"
        if "launch__registers_per_thread" in d:
            extra += f" regs={d['launch__registers_per_thread'][0]:.0f}"
        if "smsp__inst_executed.sum" in d:
            extra += f" inst={sum(d['smsp__inst_executed.sum'])/len(t):.4g}"
        print(f"{sum(t)/1e3:10.1f}us n={len(t):3d} avg={sum(t)/len(t)/1e3:9.1f}us{extra}  {k}")
    print(f"total {tot/1e3:.1f} us")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
