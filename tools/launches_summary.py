"""ncu launch list (CSV of `--metrics gpu__time_duration.sum`) -> markdown table of ONE device-resident bench step.
   python tools/launches_summary.py gpurun_out/launches_TAG.csv TAG step_ms > profiles/TAG_launches_c2.md

bench.py --steps 1 --warmup 1 runs the resident step twice (warm-up + timed), then the e2e arm and the per-stage
timings; the first resident step is the launches from the first `slab_mean_kernel` up to (excluding) the second."""
import csv
import sys
from collections import OrderedDict


def main():
    path, tag, step_ms = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    names = [r[ki] for r in rows]
    starts = [i for i, nm in enumerate(names) if "slab_mean_kernel" in nm]
    # a step has one slab_mean launch per upload slab-free resident Gramian: use the span between the first two
    # groups of the marker kernel that are separated by a collapse kernel
    ends = [i for i, nm in enumerate(names) if "collapse_median" in nm]
    first, last = starts[0], ends[0]
    agg = OrderedDict()
    for r in rows[first:last + 1]:
        nm = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(nm, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi]) / 1e6
    total = sum(v[1] for v in agg.values())
    nl = sum(v[0] for v in agg.values())
    print(f"# ncu launch list, one device-resident bench step of config 2 (500x512x512, ncomp=20) -- {tag}\n")
    print("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv python bench.py "
          "--steps 1 --warmup 1 --no-cpu`")
    print(f"(cold-cache, serialised launches: compare SHARES, not absolutes; raw CSV was gpurun_out/launches_{tag}.csv)\n")
    print("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|")
    for nm, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{nm[:64]}` | {c} | {ms:.3f} | {ms / c * 1e3:.2f} | {100 * ms / total:.1f}% |")
    print(f"\nTotal kernel time {total:.2f} ms over {nl} launches (CUDA-event step time without the profiler: "
          f"{step_ms} ms, {tag}_bench_c2.json).")


if __name__ == "__main__":
    main()
