"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/ (round tag as argv[1])."""
import collections, csv, gzip, json, os, re, shutil, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, dst = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
os.makedirs(dst, exist_ok=True)

# 1. launch list: share of every kernel in the bench command
lines = [l for l in open(os.path.join(src, f"launches_{tag}.csv")) if not l.startswith("==")]
agg = collections.OrderedDict()
started = False       # everything before the first K1 launch is model construction (table initialisation), not the step
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    started = started or "embed_fwd_kernel" in row["Kernel Name"]
    if not started:
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] == "ns" else v * 1e3 if row["Metric Unit"] == "ms" else v
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:90]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(dst, f"{tag}_launches_summary.csv"), "w") as fh:
    fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none : python bench.py --steps 2 --warmup 3 --no-cpu-baseline\n")
    fh.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes; launches before the first K1 (table initialisation) are skipped\n")
    fh.write("kernel,launches,total_us,avg_us,share_pct,ours\n")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        ours = int(any(t in k for t in ("dfm::", "g3::", "tw::", "peer_barrier", "DeviceRadixSort")))     # CUB sort: library, listed with the path
        fh.write(f"\"{k}\",{c},{t:.1f},{t / c:.2f},{100 * t / tot:.2f},{ours}\n")
with gzip.open(os.path.join(dst, f"{tag}_launches_raw.csv.gz"), "wt") as fh:
    fh.writelines(lines)

# 2. ncu --set full capture: the counters quoted in DESIGN.md / bench.py
rep = os.path.join(src, f"prof_{tag}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
with open(os.path.join(dst, f"{tag}_ncu_full_kernels.csv"), "w") as fh:
    fh.write("# ncu --set full --clock-control none --import-source on : python scripts/ncu_target.py (Criteo shape, B=65536)\n")
    fh.write(",".join(f"{w} [{units[i]}]" for w, i in idx) + "\n")
    for r in rows[2:]:
        fh.write(",".join('"' + r[i].replace('"', "'")[:80] + '"' if w == "Kernel Name" else r[i].replace(",", "") for w, i in idx) + "\n")
# 3. the tcgen05 CIN kernels (forward, backward-data, dW)
rep2 = os.path.join(src, f"prof_{tag}_cin.ncu-rep")
if os.path.exists(rep2):
    raw = subprocess.run(["ncu", "-i", rep2, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want2 = want + ["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
                    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
                    "smsp__inst_executed_pipe_tmem.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    idx = [(w, hdr.index(w)) for w in want2 if w in hdr]
    with open(os.path.join(dst, f"{tag}_ncu_full_cin_tcgen05.csv"), "w") as fh:
        fh.write("# ncu --set full --clock-control none --import-source on -k regex:cin_tc : python scripts/ncu_cin.py 8192 (F=39, D=64, CIN [128,128], TF32)\n")
        fh.write(",".join(f"{w} [{units[i]}]" for w, i in idx) + "\n")
        for r in rows[2:]:
            fh.write(",".join('"' + r[i].replace('"', "'")[:70] + '"' if w == "Kernel Name" else r[i].replace(",", "") for w, i in idx) + "\n")
    print(open(os.path.join(dst, f"{tag}_ncu_full_cin_tcgen05.csv")).read())
print(open(os.path.join(dst, f"{tag}_launches_summary.csv")).read()[:3000])
print(open(os.path.join(dst, f"{tag}_ncu_full_kernels.csv")).read())

# 4. the warp-per-sample attention kernels (config 3) and their tall-skinny parameter-gradient GEMMs
rep3 = os.path.join(src, f"prof_{tag}_attn.ncu-rep")
if os.path.exists(rep3):
    raw = subprocess.run(["ncu", "-i", rep3, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want3 = want + ["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
                    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]
    idx = [(w, hdr.index(w)) for w in want3 if w in hdr]
    with open(os.path.join(dst, f"{tag}_ncu_full_attention.csv"), "w") as fh:
        fh.write("# ncu --set full --clock-control none --import-source on -k regex:attn16|tsgemm : python scripts/ncu_attn.py (B=65536, F=16, D=16, 4 heads, A=64)\n")
        fh.write(",".join(f"{w} [{units[i]}]" for w, i in idx) + "\n")
        for r in rows[2:]:
            fh.write(",".join('"' + r[i].replace('"', "'")[:80] + '"' if w == "Kernel Name" else r[i].replace(",", "") for w, i in idx) + "\n")


# 5. dfm_gemm3 (DNN tower GEMMs): the three products of the first layer at the bench shape
rep5 = os.path.join(src, f"prof_{tag}_gemm3.ncu-rep")
if os.path.exists(rep5):
    raw = subprocess.run(["ncu", "-i", rep5, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want5 = want + ["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
                    "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
                    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
                    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
                    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
    idx = [(w, hdr.index(w)) for w in want5 if w in hdr]
    with open(os.path.join(dst, f"{tag}_ncu_gemm3.csv"), "w") as fh:
        fh.write("# ncu --set full --clock-control none --import-source on -k regex:gemm3_kernel : python scripts/ncu_gemm3.py "
                 "(B=65536, 2496->256: launches = forward Y=XW^T, dX=dY W, dW=dY^T X)\n")
        fh.write(",".join(f"{w} [{units[i]}]" for w, i in idx) + "\n")
        for r in rows[2:]:
            fh.write(",".join('"' + r[i].replace('"', "'")[:60] + '"' if w == "Kernel Name" else r[i].replace(",", "") for w, i in idx) + "\n")
    print(open(os.path.join(dst, f"{tag}_ncu_gemm3.csv")).read()[:2500])
