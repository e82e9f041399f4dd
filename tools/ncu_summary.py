"""Extract the per-launch metrics the profiles/ summaries quote from an Nsight Compute report:
python tools/ncu_summary.py report.ncu-rep [out.json]   (runs `ncu -i ... --page raw --csv`, no GPU needed)."""
import csv, io, json, subprocess, sys

KEYS = {
    "gpu__time_duration.sum": "duration_us_under_ncu",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct_of_active_cycles",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct_of_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct_of_peak",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct_of_peak_alt",
    "sm__inst_issued.avg.per_cycle_active": "ipc_issued_per_sm",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__cycles_elapsed.avg.per_second": "sm_clock_ghz_under_ncu",
    "lts__t_bytes.sum": "l2_bytes",
    "smsp__inst_executed.sum": "warp_instructions",
}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        rec = {"kernel": r[idx["Kernel Name"]][:110], "id": r[idx["ID"]]}
        for k, name in KEYS.items():
            if k in idx and r[idx[k]] not in ("", "n/a"):
                v = float(r[idx[k]].replace(",", ""))
                u = units[idx[k]]
                if name.endswith("_MB"):
                    v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                if name == "duration_us_under_ncu":
                    v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
                rec[name] = round(v, 4)
        out.append(rec)
    text = json.dumps(out, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    else:
        print(text)


if __name__ == "__main__":
    main()
