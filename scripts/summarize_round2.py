"""gpurun_out/ ncu artefacts of scripts/profile_round2.sh -> small tracked summaries under profiles/:
  r2_launches_bench.txt     launch list of the headline arm (kernel shares)
  r2_sweep_ncu_<shape>.txt  --set full capture of yk_sweep_kernel per launch shape
  sweep_metrics.json        what bench.py attaches to `roofline` (ALU pipe, L2 / DRAM bytes, shared wavefronts per launch)
  r2_lattice_ncu.txt        --set full capture of yk_lattice_kernel (when present)"""
import collections, csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}

src = os.path.join(go, "r2_launches_bench.csv")
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
    with open(os.path.join(pr, "r2_launches_bench.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, over `python bench.py --quick --steps 2 --warmup 3`\n")
        f.write("# (headline arm only: config 2, epoch rows, host commit; value + e2e legs).  Per-launch times are cold-cache and\n")
        f.write("# serialised: compare SHARES, not absolutes.  at::vectorized_elementwise_kernel = the bench's own L2 flush.\n")
        f.write(f"{'kernel':58s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>9s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:58]:58s} {v[0]:8d} {v[1]/1e6:10.3f} {v[1]/v[0]/1e3:9.2f} {v[1]/tot*100:6.1f}%\n")
    print(open(os.path.join(pr, "r2_launches_bench.txt")).read())

KEEP = re.compile(r"^(Kernel Name|gpu__time_duration.sum|dram__bytes_(read|write).sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|"
                  r"sm__throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|"
                  r"launch__(registers_per_thread|grid_size|block_size|waves_per_multiprocessor|occupancy_limit_registers|shared_mem_per_block_static|shared_mem_per_block_dynamic)|"
                  r"smsp__issue_active.avg.pct_of_peak_sustained_active|sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active|"
                  r"sm__inst_executed_pipe_(alu|lsu|fma|uniform|fp64).avg.pct_of_peak_sustained_active|"
                  r"smsp__inst_executed.sum|l1tex__m_xbar2l1tex_read_bytes.sum|l1tex__m_l1tex2xbar_write_bytes.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|lts__t_bytes.sum|"
                  r"sm__cycles_elapsed.max|sm__cycles_active.avg|smsp__average_warps_issue_stalled_[a-z_]+_per_issue_active.ratio)$")


def report(rep, out_name, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units, rows = r[0], r[1], r[2:]
    with open(os.path.join(pr, out_name), "w") as f:
        f.write(f"# {title} ({os.path.basename(rep)}), {len(rows)} launch(es)\n")
        for i, h in enumerate(hdr):
            if KEEP.search(h):
                f.write(f"{h:85s} {units[i]:16s} {' | '.join(row[i] for row in rows)}\n")

    def col(name, conv=True):
        i = hdr.index(name)
        return [float(row[i].replace(",", "")) * (UNIT.get(units[i], 1.0) if conv else 1.0) for row in rows]
    return hdr, units, rows, col


shapes = []
for shape, rows_per_launch, what in (("epochrows", 4, "config 2, epoch rows: every distinct signature (4) once per epoch -- the launches of the headline timed region"),
                                     ("fullload", 3846, "config 2 with YK_NO_ROW_SHARING: one row per ask, 3846 rows per launch"),
                                     ("masks", 3846, "config 3: taints + nodeAffinity masks, every ask its own row")):
    rep = os.path.join(go, f"r2_sweep_{shape}.ncu-rep")
    if not os.path.exists(rep):
        continue
    hdr, units, rows, col = report(rep, f"r2_sweep_ncu_{shape}.txt", f"ncu --set full --clock-control none --import-source on -k regex:yk_sweep; {what}")
    n = len(rows)
    shapes.append({"shape": shape, "what": what, "rows_per_launch": rows_per_launch,
                   "alu_pipe_pct": sum(col("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", False)) / n,
                   "issue_active_pct": sum(col("smsp__issue_active.avg.pct_of_peak_sustained_active", False)) / n,
                   # lts__t_bytes is not in this ncu's --set full: the L2 <-> SM crossbar bytes are (reads + writes)
                   "lts_bytes_per_launch": (sum(col("l1tex__m_xbar2l1tex_read_bytes.sum")) + (sum(col("l1tex__m_l1tex2xbar_write_bytes.sum")) if "l1tex__m_l1tex2xbar_write_bytes.sum" in hdr else 0.0)) / n,
                   "dram_bytes_per_launch": (sum(col("dram__bytes_read.sum")) + sum(col("dram__bytes_write.sum"))) / n,
                   "shared_wavefronts_per_launch": sum(col("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", False)) / n,
                   "duration_us_under_ncu": sum(col("gpu__time_duration.sum")) / n,
                   "source": f"profiles/r2_sweep_ncu_{shape}.txt"})
if shapes:
    json.dump({"shapes": shapes}, open(os.path.join(pr, "sweep_metrics.json"), "w"), indent=1)
    print(json.dumps(shapes, indent=1))

for name in ("r2_lattice_cfg2", "r2a_lattice_cfg2"):
    rep = os.path.join(go, name + ".ncu-rep")
    if os.path.exists(rep):
        report(rep, "r2_lattice_ncu.txt", "ncu --set full --clock-control none --import-source on -k regex:yk_lattice_kernel; config 2, whole cycle in one launch")
        print(open(os.path.join(pr, "r2_lattice_ncu.txt")).read()[:3000])
        break
