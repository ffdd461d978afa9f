"""Joins the timed config-5 attention sweep (tools/attn_probe.py --sweep) with an ncu metrics pass over the same shapes
launched twice each (--sweep --once): adds tensor-pipe %, DRAM GB/s and the ncu duration of the second launch.
    python tools/join_attn_sweep.py timed.md labels.md ncu.csv > profiles/r2_attn_sweep.md"""
import collections
import csv
import sys


def main(timed, labels, ncu_csv):
    trows = [l.rstrip("\n") for l in open(timed) if l.startswith("| ") and not l.startswith("| kind")]
    lrows = [l for l in open(labels) if l.startswith("| ") and not l.startswith("| kind")]
    assert len(trows) == len(lrows), (len(trows), len(lrows))
    with open(ncu_csv) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.OrderedDict()
    for r in csv.DictReader(lines):
        if "attn_tc_kernel" not in r["Kernel Name"]:
            continue
        per.setdefault(r["ID"], {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    launches = list(per.values())
    assert len(launches) == 2 * len(trows), (len(launches), len(trows))
    print("# BASELINE.json config 5: attention microbenchmark sweep on one B200 (heads 8, CFG batch 2, 12 frames unless "
          "stated). `us`, TFLOP/s and GB/s: CUDA-graph replays timed with CUDA events (algorithmic 4 Nq Nk d FLOPs, "
          "Q+O+K+V bytes). Tensor-pipe % and DRAM GB/s: one `ncu --metrics` launch of the same shape (cold L2, so its "
          "DRAM rate is the upper bound of what the replayed launch pulls).\n")
    print("| kind | tokens / frames / keys | C | d | us | core TFLOP/s | Q+O+KV GB/s | tensor pipe % (ncu) | DRAM GB/s (ncu) | ncu us |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for i, t in enumerate(trows):
        m = launches[2 * i + 1]
        dur = m.get("gpu__time_duration.sum", 0.0)  # ns
        by = m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        tp = m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", float("nan"))
        unit = 1.0
        print(f"{t} {tp:.1f} | {by * unit / max(dur, 1.0):.0f} | {dur / 1e3:.1f} |")


if __name__ == "__main__":
    main(*sys.argv[1:4])
