"""Selected columns of an `ncu --set full` report -> a small CSV for profiles/ (the .ncu-rep files stay in gpurun_out/)."""
import csv, re, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
keep = re.compile(r"^(ID|Kernel Name|Block Size|Grid Size)$|^(dram__bytes_(read|write)\.sum(\.per_second)?|"
                  r"dram__throughput\.avg\.pct_of_peak_sustained_elapsed|gpu__time_duration\.sum|lts__t_sector_hit_rate\.pct|"
                  r"lts__t_bytes\.sum|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|l1tex__t_sector_hit_rate\.pct|"
                  r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__inst_executed\.sum|sm__inst_executed_pipe_tensor.*\.sum|sm__pipe_tensor.*cycles_active.*|"
                  r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"launch__registers_per_thread|launch__occupancy_limit.*|launch__shared_mem_per_block_dynamic|"
                  r"smsp__average_warps?_issue_stalled_(long_scoreboard|short_scoreboard|barrier|membar|wait|math_pipe_throttle|lg_throttle|mio_throttle)_per_issue_active\.ratio|"
                  r"smsp__warp_issue_stalled_(long_scoreboard|short_scoreboard|barrier)_per_warp_active\.pct)$")
idx = [i for i, h in enumerate(hdr) if keep.match(h)]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        if len(r) >= len(hdr) - 5:
            w.writerow([r[i] if i < len(r) else "" for i in idx])
print(out, len(rows) - 2, "kernels,", len(idx), "columns")
