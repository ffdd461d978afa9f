"""Summarises an ncu `--metrics gpu__time_duration.sum --csv` launch list: the LAST denoising step (from the last
timestep_features launch to the cfg_step kernel), grouped by kernel and by (kernel, grid)."""
import collections
import csv
import re
import sys


def short(name):
    m = re.search(r"asva::(\w+)(<[^>]*>)?", name)
    if m:
        return m.group(1) + (m.group(2) or "")
    m = re.search(r"at::native::(\w+)", name)
    return "torch:" + (m.group(1) if m else name[:40])


def main(path, top=45):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = [(short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], float(r["Metric Value"]) / 1e3)
            for r in csv.DictReader(lines)]
    starts = [i for i, r in enumerate(rows) if r[0].startswith("timestep_features")]
    ends = [i for i, r in enumerate(rows) if r[0].startswith("cfg_step")]
    b = ends[-1]
    a = [i for i in starts if i < b][-1]  # the capture may end inside a later, incomplete step
    step = rows[a:b + 1]
    total = sum(r[3] for r in step)
    print(f"# last step: launches {a}..{b} ({len(step)} kernels), sum of kernel durations {total / 1e3:.3f} ms "
          f"(ncu: serialised, cold-ish caches - compare shares)")
    by = collections.defaultdict(lambda: [0, 0.0])
    for k, g, bl, t in step:
        by[k][0] += 1
        by[k][1] += t
    print("\n## by kernel\n| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {t:.1f} | {100 * t / total:.1f}% |")
    byg = collections.defaultdict(lambda: [0, 0.0])
    for k, g, bl, t in step:
        byg[(k, g)][0] += 1
        byg[(k, g)][1] += t
    print(f"\n## top {top} (kernel, grid)\n| kernel | grid | launches | total us | avg us | share |\n|---|---|---|---|---|---|")
    for (k, g), (n, t) in sorted(byg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"| {k} | {g} | {n} | {t:.1f} | {t / n:.1f} | {100 * t / total:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
