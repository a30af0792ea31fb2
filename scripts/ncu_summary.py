"""Turn an .ncu-rep (ncu --set full) into the short per-kernel table committed under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.md
"""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__cycles_elapsed.max", "cycles"),
]
cols = [(hdr.index(k), n) for k, n in want if k in hdr]
ki = hdr.index("Kernel Name")
print("| kernel | " + " | ".join(n for _, n in cols) + " |")
print("|---|" + "---|" * len(cols))
seen = {}
for r in data:
    name = r[ki].split("(")[0].split("::")[-1]
    seen[name] = seen.get(name, 0) + 1
    if seen[name] > 2:
        continue
    vals = []
    for i, _ in cols:
        v = r[i]
        try:
            f = float(v.replace(",", ""))
            v = f"{f:.4g}"
        except ValueError:
            pass
        vals.append(f"{v} {units[i]}".strip())
    print(f"| {name} | " + " | ".join(vals) + " |")
