"""gpurun_out/*.ncu-rep -> text summaries kept under profiles/ (details page for the sections that matter + selected raw
metrics incl. local-memory and DRAM counters).  Usage: python tools/make_profiles.py <rep> <out_prefix> "<title>" """
import csv, io, re, subprocess, sys


def page(rep, p):
    return subprocess.run(["ncu", "-i", rep, "--page", p, "--csv"], capture_output=True, text=True).stdout


def details(rep, out, title):
    rows = list(csv.reader(io.StringIO(page(rep, "details"))))
    hi = next(i for i, r in enumerate(rows) if "Metric Name" in r)
    ix = {n: i for i, n in enumerate(rows[hi])}
    keep = ("GPU Speed Of Light Throughput", "Compute Workload Analysis", "Memory Workload Analysis", "Scheduler Statistics",
            "Warp State Statistics", "Instruction Statistics", "Launch Statistics", "Occupancy", "Source Counters")
    lines = [title]; seen = set()
    for r in rows[hi + 1:]:
        if len(r) <= ix["Metric Value"]:
            continue
        sec, name, val, unit = r[ix["Section Name"]], r[ix["Metric Name"]], r[ix["Metric Value"]], r[ix["Metric Unit"]]
        if not name or sec not in keep or (r[ix["ID"]], sec, name) in seen:
            continue
        seen.add((r[ix["ID"]], sec, name))
        lines.append(f"[{r[ix['ID']]}] {r[ix['Kernel Name']][:60]:60s} | {sec} | {name} | {val} {unit}")
    open(out, "w").write("\n".join(lines) + "\n")
    print(out, len(lines))


def raw_selected(rep, out, title):
    rows = list(csv.reader(io.StringIO(page(rep, "raw"))))
    hi = next(i for i, r in enumerate(rows) if "ID" in r and "Kernel Name" in r)
    h = rows[hi]
    want = re.compile(r"issue_stalled.*per_issue_active|dram__bytes_(read|write)\.sum$|gpu__time_duration\.sum|smsp__inst_executed\.sum$|"
                      r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$|sm__cycles_active\.avg$|"
                      r"smsp__issue_active\.avg\.pct|launch__registers_per_thread$|launch__block_size|launch__grid_size|sm__pipe_fp64_cycles_active\.avg|"
                      r"lts__t_bytes\.sum$|sass__inst_executed_local_(loads|stores)|l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                      r"launch__occupancy_limit|smsp__sass_thread_inst_executed_op_dfma_pred_on\.sum$|sm__inst_executed_pipe_fp64")
    lines = [title]; data = rows[hi + 2:]; vals = {}
    for ci, name in enumerate(h):
        if want.search(name):
            lines.append(f"{name} [{rows[hi + 1][ci]}]: " + " ".join(f"launch{j}={r[ci]}" for j, r in enumerate(data) if len(r) > ci))
            vals[name] = [r[ci] for r in data if len(r) > ci]
    open(out, "w").write("\n".join(lines) + "\n")
    print(out, len(lines))
    return vals


if __name__ == "__main__":
    rep, prefix, title = sys.argv[1], sys.argv[2], sys.argv[3]
    details(rep, prefix + "_details.txt", "# " + title)
    v = raw_selected(rep, prefix + "_raw_selected.txt", "# selected raw metrics of the same capture: " + title)
    rd = [float(x) for x in v.get("dram__bytes_read.sum", [])]; wr = [float(x) for x in v.get("dram__bytes_write.sum", [])]
    print("dram read per launch:", rd, "write:", wr)
