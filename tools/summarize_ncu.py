"""Turns an ncu --set full report into the small text summary committed under profiles/.
   python tools/summarize_ncu.py <report.ncu-rep> <out.md> "<command line that produced it>" """
import csv, re, subprocess, sys
rep, out, cmdline = sys.argv[1], sys.argv[2], sys.argv[3]
WANT = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__pipe_tensor_cycles_active_realtime\.avg\.pct_of_peak_sustained_elapsed|sm__mem_tensor_cycles_active\.avg\.pct_of_peak_sustained_elapsed|"
                  r"TPC\.TriageCompute\.sm__pipe_tensor_cycles_active_realtime\.avg\.pct_of_peak_sustained_elapsed|"
                  r"l1tex__data_pipe_(lsu|tc)_wavefronts_mem_shared\.sum(\.pct_of_peak_sustained_elapsed)?|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|"
                  r"launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic)|sm__cycles_elapsed\.max|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__inst_executed\.avg\.per_cycle_active|sm__inst_executed_pipe_(alu|fma|lsu|xu|uniform)\.avg\.pct_of_peak_sustained_active|"
                  r"lts__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"smsp__average_warps_issue_stalled_(long_scoreboard|short_scoreboard|wait|barrier|sleeping|mio_throttle)_per_issue_active\.ratio)$")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on\n\ncommand: `{cmdline}`\n")
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        f.write(f"\n## `{name[:110]}`\n\n| metric | value | unit |\n|---|---|---|\n")
        for h, u, v in zip(hdr, units, vals):
            if WANT.match(h):
                f.write(f"| {h} | {v} | {u} |\n")
print("wrote", out)
