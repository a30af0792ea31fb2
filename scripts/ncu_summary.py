"""Turn an .ncu-rep (ncu --set full) into the short per-kernel table committed under profiles/, and (--json) into
profiles/ncu_dram_bytes.json, the per-launch DRAM traffic that bench.py reports as `roofline.traffic`.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.md
    python scripts/ncu_summary.py --json profiles/ncu_dram_bytes.json gpurun_out/prof.ncu-rep [more.ncu-rep ...]
"""
import csv, io, json, os, subprocess, sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor inst"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__inst_executed.sum", "warp inst"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__cycles_elapsed.max", "cycles"),
]
KEY = {"gma_aggregate_kernel": "gma_aggregate", "corr_lookup_ws_kernel": "corr_lookup", "corr_lookup_reg_kernel": "corr_lookup_1pair",
       "pcblock_ffn1_kernel": "pcblock_ffn1", "corr_gemm_pair_kernel": "corr_gemm",
       "corr_gemm_kernel": "corr_gemm_1cta", "gma_stats_kernel": "gma_stats", "gma_cast_kernel": "gma_cast",
       "corr_pack_kernel": "corr_pack", "gma_proj_kernel": "gma_proj"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def rows_of(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    return rows[0], rows[1], rows[2:]


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def table(rep):
    hdr, units, data = rows_of(rep)
    cols = [(hdr.index(k), n) for k, n in WANT if k in hdr]
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
            f = num(r[i])
            vals.append((f"{f:.4g}" if f is not None else r[i]) + (" " + units[i] if units[i] else ""))
        print(f"| {name} | " + " | ".join(vals) + " |")


def dram_json(out, reps):
    acc = {}
    for rep in reps:
        hdr, units, data = rows_of(rep)
        ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        for r in data:
            name = r[ki].split("(")[0].split("::")[-1].split("<")[0]
            key = KEY.get(name)
            if key is None:
                continue
            b = num(r[ri]) * UNIT[units[ri]] + num(r[wi]) * UNIT[units[wi]]
            acc.setdefault(key, []).append(b)
    res = {"source": "ncu --set full --clock-control none, " + ", ".join(os.path.basename(r) for r in reps) +
                     " (cold-cache, serialised launches; mean over the captured launches of each kernel)",
           "kernels": {k: sum(v) / len(v) for k, v in acc.items()},
           "launches": {k: len(v) for k, v in acc.items()}}
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "--json":
        dram_json(sys.argv[2], sys.argv[3:])
    else:
        table(sys.argv[1])
