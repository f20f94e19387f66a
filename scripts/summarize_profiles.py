"""Turn gpurun_out/ ncu artefacts into small tracked summaries under profiles/ (round-tagged)."""
import collections, csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
go = os.path.join(ROOT, "gpurun_out")
pr = os.path.join(ROOT, "profiles")
os.makedirs(pr, exist_ok=True)

# 1. launch list of the bench command
src = os.path.join(go, sys.argv[2] if len(sys.argv) > 2 else "launches_bench.csv")
if os.path.exists(src):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        name = re.sub(r"\(.*", "", re.sub(r"<.*", "", row["Kernel Name"])).replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(pr, f"{tag}_launches_bench.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, over `python bench.py --steps 1 --warmup 3 --no-cpu-baseline`\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"{'kernel':58s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>9s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:58]:58s} {v[0]:8d} {v[1]/1e6:10.3f} {v[1]/v[0]/1e3:9.2f} {v[1]/tot*100:6.1f}%\n")
    print(open(os.path.join(pr, f"{tag}_launches_bench.txt")).read())

# 2. full capture of the sweep kernel
rep = os.path.join(go, sys.argv[3] if len(sys.argv) > 3 else "sweep_r1.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units, rows = r[0], r[1], r[2:]
    keep = re.compile(r"^(Kernel Name|gpu__time_duration.sum|dram__bytes_(read|write).sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|"
                      r"sm__throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|"
                      r"launch__(registers_per_thread|grid_size|block_size|waves_per_multiprocessor|occupancy_limit_registers|shared_mem_per_block_static)|"
                      r"smsp__issue_active.avg.pct_of_peak_sustained_active|sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active|"
                      r"sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_(alu|lsu|fma|uniform).avg.pct_of_peak_sustained_active|"
                      r"smsp__inst_executed.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum|lts__t_bytes.sum|sm__cycles_elapsed.max|sm__cycles_active.avg|"
                      r"smsp__average_warps_issue_stalled_[a-z_]+_per_issue_active.ratio|sm__maximum_warps_per_active_cycle_pct)$")
    with open(os.path.join(pr, f"{tag}_sweep_ncu_full.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on -k regex:yk_sweep ({os.path.basename(rep)}), {len(rows)} launches\n")
        for i, h in enumerate(hdr):
            if keep.search(h):
                f.write(f"{h:85s} {units[i]:16s} {' | '.join(row[i] for row in rows)}\n")
    dr = [float(row[hdr.index('dram__bytes_read.sum')]) * {"Kbyte": 1e3, "Mbyte": 1e6, "byte": 1.0, "Gbyte": 1e9}[units[hdr.index('dram__bytes_read.sum')]] for row in rows]
    dw = [float(row[hdr.index('dram__bytes_write.sum')]) * {"Kbyte": 1e3, "Mbyte": 1e6, "byte": 1.0, "Gbyte": 1e9}[units[hdr.index('dram__bytes_write.sum')]] for row in rows]
    json.dump({"dram_bytes_per_launch": (sum(dr) + sum(dw)) / len(rows), "source": f"profiles/{tag}_sweep_ncu_full.txt",
               "read": sum(dr) / len(rows), "write": sum(dw) / len(rows)}, open(os.path.join(pr, "sweep_traffic.json"), "w"))
    print(open(os.path.join(pr, f"{tag}_sweep_ncu_full.txt")).read())
