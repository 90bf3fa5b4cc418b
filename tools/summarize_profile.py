"""Turns gpurun_out ncu artefacts into the small text summaries committed under profiles/.
   python tools/summarize_profile.py <round-tag>"""
import collections, csv, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(root, "profiles"); os.makedirs(out_dir, exist_ok=True)
go = os.path.join(root, "gpurun_out")

launches = os.path.join(go, f"launches_{tag}.csv")
if os.path.exists(launches):
    rows = list(csv.DictReader(l for l in open(launches) if not l.startswith("==")))
    agg = collections.OrderedDict()
    for r in rows:
        k = r["Kernel Name"]; v = float(r["Metric Value"].replace(",", ""))
        agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(out_dir, f"launch_list_{tag}.md"), "w") as f:
        f.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` on `python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sweep --lanes 1`\n\n")
        f.write("Per-launch times are cold-cache and serialised (compare SHARES, not absolutes). Includes the one-time weight packing and the\nsynthetic-input generation of bench.py's setup.\n\n| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:90]}` | {n} | {t/1e3:.1f} | {t/n/1e3:.2f} | {t/tot:.3f} |\n")
    print("wrote launch list", len(rows), "launches")

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second", "lts__t_sectors_srcunit_tex.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.avg.per_cycle_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio"]
for kern, cmdline in (("imu_bf16", "python tools/profile_imu.py 65536 bf16 3 nohead"), ("head_tc", "python tools/profile_imu.py 65536 bf16 3"),
                      ("linear_tc", "python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sweep --lanes 1 (62nd linear_tc launch = the late-fusion layer, 896 -> 128, 2 CTAs, fp32-row staging path, 14 k chunks)"),
                      ("maha_score", "python tools/profile_scoring.py (2 M feature rows; 3rd launch captured)"),
                      ("logit_ring", "python tools/profile_scoring.py (4 M logit rows; 3rd launch captured)"),
                      ("maha_fit", "python tools/profile_fit_pool.py (2 M feature rows; 3rd launch captured)"),
                      ("video_pool", "python tools/profile_fit_pool.py (2 048 clips, 537 MB; 3rd launch captured)")):
    rep = os.path.join(go, f"prof_{kern}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(os.path.join(out_dir, f"{kern}_kernel_{tag}.md"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on, kernel regex `{kern}` ({tag})\n\ncommand: `{cmdline}` \n\n| metric | value | unit |\n|---|---|---|\n")
        for h, u, v in zip(hdr, units, vals):
            if h in WANT:
                f.write(f"| {h} | {v} | {u} |\n")
    print("wrote kernel summary", kern)
