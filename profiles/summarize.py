"""Turn ncu outputs in gpurun_out/ into the text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches.txt
    python profiles/summarize.py full     gpurun_out/prof_small_r1.ncu-rep > profiles/r1_kf_small_full.txt
"""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    per = []
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e6 if r[ui] == "ns" else v / 1e3 if r[ui] == "us" else v
        agg.setdefault(r[ki], [0, 0.0])
        agg[r[ki]][0] += 1
        agg[r[ki]][1] += v
        per.append((r[ki], v))
    tot = sum(v for _, v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised:")
    print("# compare SHARES, not absolutes)")
    print(f"# total {tot:.3f} ms over {len(per)} launches")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v:10.3f} ms  {c:4d}x  {100 * v / tot:5.1f}%  {k[:150]}")
    print("# per launch (ms):")
    for k, v in per:
        print(f"{v:10.3f}  {k[:110]}")


WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "launch__grid_size", "launch__block_size",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "smsp__average_warp_latency_per_inst_issued.ratio",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum", "sm__cycles_active.avg",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct",
    "smsp__sass_average_data_bytes_per_sector_mem_global_op_st.pct",
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    for D in rows[2:]:
        name = D[H.index("Kernel Name")] if "Kernel Name" in H else "?"
        print("## kernel:", name[:160])
        vals = {}
        for i, h in enumerate(H):
            if h in WANT:
                print(f"{h:75s} {D[i]:>18s} {U[i]}")
                vals[h] = D[i]
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                try:
                    if float(D[i]) >= 0.05:
                        print(f"{h:75s} {D[i]:>18s} (warps stalled per issue-active cycle)")
                except ValueError:
                    pass
        try:
            rd = float(vals["dram__bytes_read.sum"].replace(",", ""))
            wr = float(vals["dram__bytes_write.sum"].replace(",", ""))
            print(f"# traffic = dram read + write = {rd + wr:.3f} (unit as above) per launch")
        except Exception:
            pass


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
