"""Summarise a chrome trace written by `SB200_TRACE=prefix python bench.py ...` (torch.profiler / CUPTI):
GPU activities of the LAST profiled step in time order with the idle gaps between them, busy time per
stream, and the totals per kernel name.

    python tools/trace_summary.py gpurun_out/trace_rank0.json [--all]
"""
import json
import sys
from collections import defaultdict


def main(path, show_all=False):
    ev = json.load(open(path))["traceEvents"]
    gpu = [e for e in ev if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    gpu.sort(key=lambda e: e["ts"])
    if not gpu:
        print("no GPU activity in the trace")
        return
    # steps are delimited by the fused velocity kernel (last kernel of a step on the main stream)
    ends = [e["ts"] + e["dur"] for e in gpu if "sb_velocity" in e["name"]]
    t0, t1 = (ends[-2], ends[-1]) if len(ends) >= 2 else (gpu[0]["ts"], gpu[-1]["ts"] + gpu[-1]["dur"])
    step = [e for e in gpu if t0 <= e["ts"] < t1]
    print(f"last profiled step: {(t1 - t0) / 1e3:.3f} ms, {len(step)} GPU activities")
    by_stream = defaultdict(float)
    by_name = defaultdict(lambda: [0, 0.0])
    for e in step:
        by_stream[e["args"].get("stream", e.get("tid"))] += e["dur"]
        k = by_name[e["name"][:70]]
        k[0] += 1
        k[1] += e["dur"]
    print("busy per stream [ms]:", {k: round(v / 1e3, 3) for k, v in by_stream.items()})
    # union of busy intervals over all streams -> idle time of the GPU inside the step
    busy, cur_end = 0.0, t0
    for e in step:
        s, f = max(e["ts"], cur_end), e["ts"] + e["dur"]
        if f > s:
            busy += f - s
            cur_end = f
    print(f"GPU busy (any stream) {busy / 1e3:.3f} ms, idle {(t1 - t0 - busy) / 1e3:.3f} ms")
    print("per kernel name:")
    for name, (n, d) in sorted(by_name.items(), key=lambda kv: -kv[1][1]):
        print(f"  {d / 1e3:8.3f} ms  n={n:3d}  {name}")
    if show_all:
        print("timeline (start offset us, duration us, stream, name):")
        prev_end = t0
        for e in step:
            gap = e["ts"] - prev_end
            print(f"  {e['ts'] - t0:9.1f} {e['dur']:8.1f} s{e['args'].get('stream', '?'):<3} gap={gap:7.1f}  {e['name'][:80]}")
            prev_end = max(prev_end, e["ts"] + e["dur"])


if __name__ == "__main__":
    main(sys.argv[1], "--all" in sys.argv)
