"""Turn an `ncu --set full --import-source on` report into the text summary kept under profiles/.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep [kernel-regex] > profiles/rN_<name>.txt

Per kernel: the headline counters (duration, registers, occupancy, issue / FP64 / DRAM / L1 utilisation, DRAM
bytes, shared-memory wavefronts and bank conflicts, local-memory sectors), the warp-stall breakdown, and the
stall samples per barrier-delimited phase of the kernel (from the source page; needs -lineinfo).
"""
import csv
import io
import subprocess
import sys

COUNTERS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rx = sys.argv[2] if len(sys.argv) > 2 else None
    sel = ["--kernel-name", "regex:" + rx] if rx else []
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"] + sel))))
    hdr, units, data = raw[0], raw[1], raw[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    print("# %s" % " ".join(sys.argv[1:]))
    seen = set()
    for d in data:
        name = d[idx["Kernel Name"]]
        if name in seen:
            continue
        seen.add(name)
        print("\n## %s   grid %s block %s" % (name, d[idx.get("Grid Size", 0)] if "Grid Size" in idx else "", d[idx["Block Size"]] if "Block Size" in idx else ""))
        for c in COUNTERS:
            if c in idx:
                print("%-78s %-14s %s" % (c, units[idx[c]], d[idx[c]]))
        st = sorted(((float(d[idx[s]]), s.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for s in stalls), reverse=True)
        print("warp stall reasons (cycles per issued instruction): " + ", ".join("%s=%.2f" % (n, v) for v, n in st[:9]))
    # source page: samples per barrier-delimited phase
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"] + sel))))
    i = 0
    done = set()
    while i < len(src):
        if src[i] and src[i][0] == "Kernel Name":
            kname = src[i][1]
            h = src[i + 1]
            ix = {x: k for k, x in enumerate(h)}
            j = i + 2
            rows = []
            while j < len(src) and not (src[j] and src[j][0] == "Kernel Name"):
                if len(src[j]) == len(h):
                    rows.append(src[j])
                j += 1
            i = j
            if kname in done or "# Samples" not in ix:
                continue
            done.add(kname)
            scols = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
            tot = sum(int(r[ix["# Samples"]]) for r in rows) or 1
            print("\n## %s: warp-stall samples per phase (phases are delimited by BAR.SYNC; %d samples, %d SASS instructions)" % (kname, tot, len(rows)))
            start, acc, cur, winst = 0, 0, {}, 0
            nph = 0
            for k, r in enumerate(rows + [None]):
                if r is not None:
                    acc += int(r[ix["# Samples"]])
                    winst += int(r[ix["Instructions Executed"]])
                    for s in scols:
                        cur[s] = cur.get(s, 0) + int(r[ix[s]])
                if r is None or "BAR.SYNC" in r[ix["Source"]]:
                    if acc > 0.01 * tot:
                        top = sorted(cur.items(), key=lambda kv: -kv[1])[:3]
                        print("  phase %2d  SASS %5d-%5d  %5.1f%% of samples  %9d warp-instr  top stalls: %s" % (
                            nph, start, k, 100.0 * acc / tot, winst, ", ".join("%s %d" % (a.replace("stall_", ""), b) for a, b in top)))
                    nph += 1
                    start, acc, cur, winst = k + 1, 0, {}, 0
        else:
            i += 1


if __name__ == "__main__":
    main()
