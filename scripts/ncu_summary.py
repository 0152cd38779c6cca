"""Turn an .ncu-rep (ncu --set full) into the small JSON summary committed under profiles/:
key metrics, stall-reason samples, and the top stalled SASS instructions.  usage: ncu_summary.py in.ncu-rep out.json"""
import csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, dst):
    raw = page(rep, "raw")
    hdr, units, val = raw[0], raw[1], raw[2]
    idx = {h: i for i, h in enumerate(hdr)}
    res = {"kernel": val[idx["Kernel Name"]] if "Kernel Name" in idx else None, "metrics": {}, "stall_samples": {}}
    for k in KEYS:
        if k in idx:
            res["metrics"][k] = [val[idx[k]], units[idx[k]]]
    for h, i in idx.items():
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try:
                res["stall_samples"][h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = int(val[i])
            except ValueError:
                pass
    src = page(rep, "source")
    sh = src[1]
    ci = {h: i for i, h in enumerate(sh)}
    rows = src[2:]
    tot = sum(int(r[ci["# Samples"]]) for r in rows)
    stalls = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
    top = []
    for r in sorted(rows, key=lambda r: -int(r[ci["# Samples"]]))[:25]:
        best = max(stalls, key=lambda h: int(r[ci[h]]))
        top.append({"sass": r[ci["Source"]].strip(), "samples_pct": round(100.0 * int(r[ci["# Samples"]]) / max(1, tot), 2),
                    "executed": int(r[ci["Instructions Executed"]]), "main_stall": best})
    res["top_sass_by_samples"] = top
    res["total_samples"] = tot
    res["warp_instructions_executed"] = sum(int(r[ci["Instructions Executed"]]) for r in rows)
    json.dump(res, open(dst, "w"), indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
