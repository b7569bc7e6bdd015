"""Turn an `ncu --set full` report into the small metric/unit/value extracts kept under profiles/.

    python tools/ncu_extract.py gpurun_out/x.ncu-rep profiles/r02_x        # one CSV per captured launch

Runs `ncu -i <rep> --page raw --csv` (no GPU needed) and keeps the metrics the roofline / stall discussion uses:
duration, DRAM bytes, L2 / L1TEX traffic and hit rates, instruction counts, issue-slot use, occupancy, registers,
shared memory and the per-issue warp stall reasons.
"""
import csv
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes(_read|_write)?\.sum(\.per_second)?|dram__throughput\.avg\.pct_of_peak_sustained_elapsed"
    r"|dram__sectors_(read|write)\.sum|lts__t_sectors(_op_(read|write|red|atom))?\.sum|lts__t_requests_srcunit_tex\.sum"
    r"|lts__t_sector_hit_rate\.pct|lts__t_sector_op_(read|red|write)_hit_rate\.pct|lts__throughput\.avg\.pct_of_peak_sustained_elapsed"
    r"|l1tex__t_sector_hit_rate\.pct|l1tex__throughput\.avg\.pct_of_peak_sustained_(elapsed|active)"
    r"|l1tex__data_pipe_lsu_wavefronts(_mem_shared)?\.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum"
    r"|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__inst_executed\.sum|smsp__inst_executed\.sum|smsp__thread_inst_executed\.sum"
    r"|sm__inst_executed_pipe_[a-z_0-9]+\.sum|smsp__inst_executed_op_[a-z_]+\.sum"
    r"|sm__issue_active\.avg\.pct_of_peak_sustained_elapsed|smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__issue_active\.avg\.per_cycle_active"
    r"|sm__warps_active\.avg\.pct_of_peak_sustained_active|sm__warps_active\.avg\.per_cycle_active|sm__maximum_warps_per_active_cycle_pct"
    r"|smsp__warps_eligible\.avg\.per_cycle_active|smsp__average_warp_latency_per_inst_issued\.ratio"
    r"|smsp__average_warps_issue_stalled_[a-z_]+_per_issue_active\.ratio"
    r"|launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_(static|dynamic)|occupancy_limit_[a-z_]+|waves_per_multiprocessor|occupancy_per_block_size)"
    r"|sm__cycles_elapsed\.(avg|max)|sm__cycles_active\.avg|smsp__cycles_active\.avg"
    r"|smsp__sass_inst_executed_op_(global|shared|local)_(ld|st|red|atom)?[a-z_]*\.sum)$"
)


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    seen = {}
    for r in rows[2:]:
        short = re.sub(r"^void ", "", r[name_col]).split("(")[0]
        short = re.sub(r"[<>, ]+", "_", short).strip("_")
        seen[short] = seen.get(short, 0) + 1
        path = f"{prefix}_{short}" + (f"_{seen[short]}" if seen[short] > 1 else "") + "_ncu_full.csv"
        with open(path, "w") as f:
            f.write("metric,unit,value\n")
            f.write(f"kernel,,\"{r[name_col][:120]}\"\n")
            for h, u, v in zip(hdr, units, r):
                if KEEP.match(h):
                    f.write(f"{h},{u},{v}\n")
        print(path)


if __name__ == "__main__":
    main()
