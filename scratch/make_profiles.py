"""gpurun_out/prof_r01_final_*_{details,raw}.csv (ncu --page details/raw --csv exports) -> profiles/r01_final_*.txt summaries."""
import csv, io, re, json
def details_from_csv(text, out, title):
    rows = list(csv.reader(io.StringIO(text)))
    hi = next(i for i, r in enumerate(rows) if "Metric Name" in r)
    h = rows[hi]; ix = {n: i for i, n in enumerate(h)}
    keep = ("GPU Speed Of Light Throughput", "Compute Workload Analysis", "Memory Workload Analysis", "Scheduler Statistics",
            "Warp State Statistics", "Instruction Statistics", "Launch Statistics", "Occupancy", "Source Counters")
    lines = [title]; seen = set()
    for r in rows[hi + 1:]:
        if len(r) <= ix["Metric Value"]: continue
        sec, name, val, unit = r[ix["Section Name"]], r[ix["Metric Name"]], r[ix["Metric Value"]], r[ix["Metric Unit"]]
        if not name or sec not in keep or (r[ix["ID"]], sec, name) in seen: continue
        seen.add((r[ix["ID"]], sec, name))
        lines.append(f"[{r[ix['ID']]}] {r[ix['Kernel Name']][:60]:60s} | {sec} | {name} | {val} {unit}")
    open(out, "w").write("\n".join(lines) + "\n"); print(out, len(lines))
def raw_selected(text, out, title):
    rows = list(csv.reader(io.StringIO(text)))
    hi = next(i for i, r in enumerate(rows) if "ID" in r and "Kernel Name" in r)
    h = rows[hi]
    want = re.compile(r"issue_stalled.*per_issue_active|dram__bytes_(read|write)\.sum$|gpu__time_duration\.sum|smsp__inst_executed\.sum$|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$|sm__cycles_active\.avg$|smsp__issue_active\.avg\.pct|launch__registers_per_thread|launch__block_size|launch__grid_size|sm__pipe_fp64_cycles_active\.avg|lts__t_bytes\.sum$")
    lines = [title]; data = rows[hi + 2:]; vals = {}
    for ci, name in enumerate(h):
        if want.search(name):
            lines.append(f"{name} [{rows[hi+1][ci]}]: " + " ".join(f"launch{j}={r[ci]}" for j, r in enumerate(data) if len(r) > ci))
            vals[name] = [r[ci] for r in data if len(r) > ci]
    open(out, "w").write("\n".join(lines) + "\n"); print(out, len(lines))
    return vals
g = "gpurun_out/prof_r01_final_"
details_from_csv(open(g + "B1024_details.csv").read(), "profiles/r01_final_B1024_R4_details.txt",
                 "# ncu --set full, scratch/prof_case.py 1024 4 (final code of round 1), third solve of the batch: launch 0 = hard queue (one CTA + 3 PCR assistant warps per SM, 224 threads), launch 1 = the other instances, two 128-thread CTAs per SM, launch 2 = instances migrated after 300 iterations (none left: history flags them)")
v = raw_selected(open(g + "B1024_raw.csv").read(), "profiles/r01_final_B1024_R4_raw_selected.txt", "# selected raw metrics of the same capture")
details_from_csv(open(g + "B16384_details.csv").read(), "profiles/r01_final_B16384_R4_details.txt",
                 "# ncu --set full, scratch/prof_case.py 16384 4 (final code of round 1), second solve of the batch (history known): launch 0 = mpcqp_setup_kernel (three 128-thread CTAs per SM: Ruiz scaling, rho vector, warm start; leaves every instance parked at iteration 0), launch 1 = the solve launch, two CTAs per SM, resuming all 16,384 instances, hard list first")
raw_selected(open(g + "B16384_raw.csv").read(), "profiles/r01_final_B16384_R4_raw_selected.txt", "# selected raw metrics of the same capture")
details_from_csv(open(g + "sweep_details.csv").read(), "profiles/r01_final_sweep_wide_details.txt",
                 "# ncu --set full, scratch/prof_sweep.py (final code of round 1): wide CTA kernel (run-time obstacle count <= 32, rows in shared memory) on the (5 m/s, 20 m/s^2) slice of the first 4,096 sweep instances")
raw_selected(open(g + "sweep_raw.csv").read(), "profiles/r01_final_sweep_wide_raw_selected.txt", "# selected raw metrics of the same capture")
rd = [float(x) for x in v["dram__bytes_read.sum"]]; wr = [float(x) for x in v["dram__bytes_write.sum"]]
print("dram MB per launch:", rd, wr)
json.dump({"source": "profiles/r01_final_B1024_R4_raw_selected.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, all solve launches of one step)",
           "workload": "configs[1]: 1024 QPs, 4 obstacles", "bytes_per_step": int((sum(rd) + sum(wr)) * 1e6)}, open("profiles/r01_final_traffic.json", "w"), indent=1)
