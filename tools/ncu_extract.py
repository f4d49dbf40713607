"""Key metrics of `ncu --set full` reports as a markdown table + profiles/ncu_traffic.json entries.
usage: python tools/ncu_extract.py WORKLOAD stage=report.ncu-rep ..."""
import csv
import json
import os
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time (us)"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp instr"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("sm__inst_executed_pipe_tensor.sum", "tensor pipe instr")]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def main():
    workload = sys.argv[1]
    traffic_path = os.path.join("profiles", "ncu_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    traffic.setdefault(workload, {})
    for arg in sys.argv[2:]:
        stage, rep = arg.split("=")
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units, data = rows[0], rows[1], rows[2:]
        kn = hdr.index("Kernel Name")
        print(f"\n### {stage}: `{data[0][kn][:70]}` ({len(data)} launches captured, {os.path.basename(rep)})\n")
        print("| metric | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |")
        print("|---|" + "---:|" * len(data))
        dram = [0.0] * len(data)
        for key, label in WANT:
            if key not in hdr:
                continue
            i = hdr.index(key)
            vals = []
            for j, r in enumerate(data):
                v = r[i]
                if key.startswith("dram__bytes"):
                    b = float(v.replace(",", "")) * UNIT.get(units[i], 1.0)
                    dram[j] += b
                    vals.append(f"{b / 1e6:.2f} MB")
                elif key in ("launch__grid_size", "launch__block_size", "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum"):
                    vals.append(f"{float(v.replace(',', '')):,.0f}")
                else:
                    vals.append(f"{float(v.replace(',', '')):.2f}")
            print(f"| {label} | " + " | ".join(vals) + " |")
        traffic[workload][stage] = {"dram_bytes_per_launch": int(sum(dram) / len(dram)), "launches_captured": len(data), "report": os.path.basename(rep)}
    json.dump(traffic, open(traffic_path, "w"), indent=1)


if __name__ == "__main__":
    main()
