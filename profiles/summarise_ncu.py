"""Summarise an `ncu --set full` report of the clip kernel into (a) a small metric table (CSV, committed) and
(b) profiles/clip_kernel_traffic.json, the per-clip-step DRAM traffic bench.py puts into `roofline.traffic`.

    python profiles/summarise_ncu.py gpurun_out/clip_kernel_r02.ncu-rep profiles/r02_clip_kernel_ncu_full_summary.csv \
        --clips 148 --steps 12 --note "round-2 kernel"

Reads the report with `ncu -i <rep> --page raw --csv` (works without a GPU)."""
import argparse
import csv
import io
import json
import os
import subprocess

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "sm__inst_executed_pipe_uniform.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out_csv")
    ap.add_argument("--clips", type=int, required=True)
    ap.add_argument("--steps", type=int, required=True)
    ap.add_argument("--note", default="")
    ap.add_argument("--kernel", default="clip_kernel")
    ap.add_argument("--all", action="store_true", help="keep every metric, not just the short list")
    ap.add_argument("--no-traffic-json", action="store_true", help="do not rewrite clip_kernel_traffic.json (captures of other launch shapes)")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    data = [r for r in rows[2:] if any(a.kernel in c for c in r)]
    if not data:
        raise SystemExit("no launch of %s in the report" % a.kernel)
    r = data[-1]
    vals = {h: (u, v) for h, u, v in zip(hdr, units, r)}
    with open(a.out_csv, "w") as f:
        f.write(f"# ncu --set full --clock-control none -k {a.kernel}: {a.note}; {a.clips} clips x {a.steps} DDPM steps\n")
        f.write("metric,unit,value\n")
        for k, (u, v) in vals.items():
            short = k.split(".")[0]
            variant_ok = (".max" not in k and ".min" not in k and ".sum.p" not in k and "pct_of_peak_sustained_elapsed" not in k) or k in KEEP
            if a.all or k in KEEP or (variant_ok and (short.startswith("sm__inst_executed_pipe_") or "tensor" in short)):
                zero = v.replace(",", "").replace(".", "").strip("0") == ""
                if v not in ("", "n/a") and (k in KEEP or (not zero and ".peak_sustained" not in k and ".per_second" not in k)):
                    f.write(f"{k},{u},{v}\n")

    def num(k):
        u, v = vals[k]
        x = float(v.replace(",", ""))
        return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
    if "dram__bytes_read.sum" in vals and not a.no_traffic_json:
        per = (num("dram__bytes_read.sum") + num("dram__bytes_write.sum")) / (a.clips * a.steps)
        out = {"dram_bytes_per_clip_step": per, "source": os.path.basename(a.out_csv) + f" ({a.clips} clips x {a.steps} steps)",
               "clips": a.clips, "steps": a.steps}
        with open(os.path.join(os.path.dirname(os.path.abspath(a.out_csv)), "clip_kernel_traffic.json"), "w") as f:
            json.dump(out, f, indent=1)
        print("DRAM bytes per clip-step: %.3f MB" % (per / 1e6))


if __name__ == "__main__":
    main()
