"""Summarise an .ncu-rep (read here with `ncu -i ... --page raw --csv`) into a small JSON for profiles/.
Usage: python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/NAME.json [algorithmic_bytes_per_launch]"""
import csv, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg", "launch__occupancy_limit_shared_mem"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}

def main():
    rep, out = sys.argv[1], sys.argv[2]
    algo = float(sys.argv[3]) if len(sys.argv) > 3 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    launches = []
    for d in data:
        rec = {"kernel": d[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    rec[k] = {"value": float(d[i].replace(",", "")), "unit": units[i]}
                except ValueError:
                    rec[k] = {"value": d[i], "unit": units[i]}
        launches.append(rec)
    def b(rec, k):
        v = rec.get(k)
        return None if v is None else v["value"] * SCALE.get(v["unit"], 1.0)
    dram = [(b(r, "dram__bytes_read.sum") or 0) + (b(r, "dram__bytes_write.sum") or 0) for r in launches]
    summary = {"report": rep, "command": "see profiles/README.md", "launches": launches,
               "dram_bytes_per_launch": sum(dram) / len(dram) if dram else None,
               "algorithmic_bytes_per_launch": algo,
               "traffic_over_algorithmic": (sum(dram) / len(dram) / algo) if (dram and algo) else None}
    json.dump(summary, open(out, "w"), indent=1)
    print(json.dumps({k: summary[k] for k in ("dram_bytes_per_launch", "traffic_over_algorithmic")}))

main()
