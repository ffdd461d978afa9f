"""Reads `ncu --set full` reports (gpurun_out/*.ncu-rep) with `ncu -i ... --page raw --csv` and prints one markdown row
per report: duration, tensor-pipe %, DRAM bytes / throughput, L2 throughput, issue-slot use, registers, shared memory.
    python tools/ncu_summary.py name=path.ncu-rep ... > profiles/<file>.md"""
import csv
import io
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "us",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor %",
    "dram__bytes_read.sum": "DRAM rd MB",
    "dram__bytes_write.sum": "DRAM wr MB",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "DRAM %",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "L2 %",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue %",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps %",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "sm__cycles_elapsed.avg.per_second": "SM GHz",
}


def read(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in WANT or h == "Kernel Name":
            d[h] = (v, u)
    return d


def main():
    cols = list(WANT.values())
    print("| launch | kernel | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for arg in sys.argv[1:]:
        name, path = arg.split("=", 1)
        d = read(path)
        cells = []
        for k in WANT:
            v, u = d.get(k, ("", ""))
            try:
                x = float(v.replace(",", ""))
            except ValueError:
                cells.append(v)
                continue
            if k == "gpu__time_duration.sum":
                x = x / 1e3 if u in ("ns", "nsecond") else (x if u in ("us", "usecond") else x * 1e3)
                cells.append(f"{x:.1f}")
            elif k.startswith("dram__bytes"):
                scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                cells.append(f"{x * scale:.2f}")
            elif k == "sm__cycles_elapsed.avg.per_second":
                scale = {"hz": 1e-9, "Khz": 1e-6, "Mhz": 1e-3, "Ghz": 1.0}.get(u, 1.0)
                cells.append(f"{x * scale:.2f}")
            elif k.startswith("launch__"):
                cells.append(f"{int(x)}")
            else:
                cells.append(f"{x:.1f}")
        kn = d.get("Kernel Name", ("?", ""))[0]
        kn = kn[kn.find("asva::") + 6:] if "asva::" in kn else kn
        print(f"| {name} | `{kn[:44]}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
