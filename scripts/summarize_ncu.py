#!/usr/bin/env python
"""Turns an `ncu --set full` report into the committed summaries (run HERE, no GPU needed):
  python scripts/summarize_ncu.py gpurun_out/r02_full.ncu-rep profiles/r02_ncu_full_summary.json [--traffic profiles/ncu_traffic.json --frames 64 --templates 3000]
Per kernel (first profiled launch of each name): duration, issue/pipe utilisation, L1/L2/DRAM throughput, achieved occupancy,
registers, shared-memory wavefronts, DRAM and L2 bytes.  --traffic also rewrites profiles/ncu_traffic.json (roofline.traffic and
the L2 bytes per launch that bench.py divides by its live CUDA-event time)."""
import csv
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__waves_per_multiprocessor",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def to_bytes(v, unit):
    f = float(v)
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    summary, traffic = {}, {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("lmk::", "")
        if name in summary:
            continue
        d = {}
        for k in KEEP:
            if k in idx and r[idx[k]] not in ("", "n/a"):
                v = r[idx[k]].replace(",", "")
                try:
                    d[k] = to_bytes(v, units[idx[k]]) if "bytes" in k else float(v)
                except ValueError:
                    pass
                if k == "gpu__time_duration.sum":
                    d[k + ".unit"] = units[idx[k]]
        st = sorted(((float(r[idx[h]]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stall
                     if r[idx[h]] not in ("", "n/a")), reverse=True)[:5]
        d["top_stalls_per_issue"] = {n: round(v, 2) for v, n in st}
        summary[name] = d
        if "dram__bytes_read.sum" in d:
            traffic[name] = {"dram_bytes_per_launch": d["dram__bytes_read.sum"] + d.get("dram__bytes_write.sum", 0.0),
                             # L2 -> L1 read bytes (crossbar into the SMs) + L1 -> L2 write sectors: what the kernel moved through L2
                             "l2_bytes_per_launch": d.get("l1tex__m_xbar2l1tex_read_bytes.sum", 32.0 * d.get("lts__t_sectors_srcunit_tex_op_read.sum", 0.0))
                                                    + 32.0 * d.get("lts__t_sectors_srcunit_tex_op_write.sum", 0.0),
                             "l1_sector_bytes_per_launch": 32.0 * d.get("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", 0.0)}
    json.dump(summary, open(out, "w"), indent=1)
    print("wrote", out, "kernels:", list(summary))
    if "--traffic" in sys.argv:
        tp = sys.argv[sys.argv.index("--traffic") + 1]
        frames = int(sys.argv[sys.argv.index("--frames") + 1]); templates = int(sys.argv[sys.argv.index("--templates") + 1])
        for k in traffic:
            traffic[k].update({"frames": frames, "templates": templates, "source": "%s (ncu --set full --clock-control none; one launch)" % out})
        # bench.py keys the similarity kernels by their plain names
        for plain in ("similarity_coarse_kernel", "similarity_local_kernel"):
            for k in list(traffic):
                if k.startswith(plain) and plain not in traffic:
                    traffic[plain] = traffic[k]
        json.dump(traffic, open(tp, "w"), indent=1)
        print("wrote", tp)


if __name__ == "__main__":
    main()
