"""gpurun_out/ ncu artefacts of scripts/profile_uniform.sh -> tracked summaries under profiles/:
  r2b_uniform_launches.txt  launch list of two cycles of the reference benchmark shape on the engine's default (auto -> device)
  r2b_uniform_ncu.txt       --set full capture of the uniform-run kernels (one launch each)"""
import collections, csv, os, re, subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
lines = [l for l in open(os.path.join(go, "r2_launches_uniform.csv")) if not l.startswith("==")]
agg = collections.OrderedDict()
for r in csv.DictReader(lines):
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except Exception:
        continue
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r["Metric Unit"], 1.0)
    name = re.sub(r"\(.*", "", re.sub(r"<.*", "", r["Kernel Name"])).replace("void ", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
with open(os.path.join(pr, "r2b_uniform_launches.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none over `python scripts/prof_cycle.py reference auto`:\n"
            "# TWO cycles of the reference benchmark shape (5 000 nodes x 50 000 identical pods), commit on the device (one uniform\n"
            "# run per cycle).  Per-launch times are cold-cache and serialised: compare SHARES.  Radix sorts per cycle: the node order\n"
            "# (yk_key_kernel's keys, 5 000 x 64 bit) and the elements (160 000 x 64 bit).\n")
    f.write(f"{'kernel':58s} {'launches':>8s} {'total_us':>10s} {'avg_us':>9s} {'share':>7s}\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k[:58]:58s} {v[0]:8d} {v[1]/1e3:10.1f} {v[1]/v[0]/1e3:9.2f} {v[1]/tot*100:6.1f}%\n")
    f.write(f"{'total':58s} {sum(v[0] for v in agg.values()):8d} {tot/1e3:10.1f}\n")
print(open(os.path.join(pr, "r2b_uniform_launches.txt")).read())
rep = os.path.join(go, "r2_uniform.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units, rows = r[0], r[1], r[2:]
    keep = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active"]
    with open(os.path.join(pr, "r2b_uniform_ncu.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none, kernels of csrc/yk_uniform.cuh, one launch each ({os.path.basename(rep)})\n")
        for h in keep:
            if h in hdr:
                i = hdr.index(h)
                f.write(f"{h:70s} {units[i]:14s} {' | '.join(row[i][:34] for row in rows)}\n")
    print(open(os.path.join(pr, "r2b_uniform_ncu.txt")).read())
