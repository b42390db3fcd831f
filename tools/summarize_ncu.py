"""Turns ncu outputs under gpurun_out/ into the tracked summaries under profiles/ (markdown + traffic.json)."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name):
    m = re.search(r"eval_kernel<\(?(?:int\))?(\d+), *\(?(?:int\))?(\d+)>", name)
    if m:
        return "eval_kernel<mode=%s,cap=%s>" % (m.group(1), m.group(2))
    return re.sub(r"\(.*", "", name).split("::")[-1].replace("void ", "").strip()


def launch_list(csv_path, title, out_md):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        agg[short(r[ki])].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(out_md, "w") as f:
        f.write("# %s\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: compare SHARES, not absolutes)\n\n" % title)
        f.write("| kernel | launches | total us | avg us | min us | max us | share |\n|---|---|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("| %s | %d | %.1f | %.2f | %.2f | %.2f | %.1f%% |\n" % (k, len(v), sum(v) / 1e3, sum(v) / len(v) / 1e3, min(v) / 1e3, max(v) / 1e3, 100 * sum(v) / tot))
    return agg


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def full_report(rep, title, out_md):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    with open(out_md, "w") as f:
        f.write("# %s\n\n`ncu --set full --clock-control none` raw page, selected metrics per captured launch\n\n" % title)
        for r in rows[2:]:
            name = short(r[hdr.index("Kernel Name")])
            f.write("## %s  grid %s block %s\n\n| metric | value | unit |\n|---|---|---|\n" % (name, r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
            d = {"kernel": name}
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write("| %s | %s | %s |\n" % (w, r[i], units[i]))
                    d[w] = (r[i], units[i])
            f.write("\n")
            res.append(d)
    return res


if __name__ == "__main__":
    # processes whichever captures are present under gpurun_out/ (scratch); summaries of absent ones are left as committed
    g = os.path.join(ROOT, "gpurun_out")
    p = os.path.join(ROOT, "profiles")
    os.makedirs(p, exist_ok=True)
    WANT.extend(["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
                 "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
                 "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "dram__bytes_read.sum.per_second"])

    def have(name):
        return os.path.exists(os.path.join(g, name))

    traffic_path = os.path.join(p, "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    if have("launches_r01d.csv"):
        launch_list(os.path.join(g, "launches_r01d.csv"), "r01 launch list: bench.py --streams 32 --steps 2 --warmup 3 (KITTI 1232x368, flat-grid DMMA eval kernel, "
                    "pyramids on the overlap stream)", os.path.join(p, "r01_launches.md"))
    if have("prof_eval_r01d.ncu-rep"):
        ev = full_report(os.path.join(g, "prof_eval_r01d.ncu-rep"), "r01 eval_kernel (fused residual / Jacobian / DMMA normal equations, flat grid) inside bench.py, "
                         "32 streams / 8 groups", os.path.join(p, "r01_eval_kernel.md"))
        tr = [to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"]) for d in ev if "dram__bytes_read.sum" in d]
        if tr:
            traffic["pose_eval_dram_bytes_per_launch"] = sum(tr) / len(tr)
            traffic["source"] = "profiles/r01_eval_kernel.md (mean over %d captured launches of the eval kernel inside bench.py --streams 32)" % len(tr)
    if have("prof_eval128_r01d.ncu-rep"):
        full_report(os.path.join(g, "prof_eval128_r01d.ncu-rep"), "r01 eval_kernel, 128 pose items of 9.9k points in one launch (tools/one_eval.py kitti 128)",
                    os.path.join(p, "r01_eval_kernel_128items.md"))
    if have("s3_sc_tile2.ncu-rep"):
        sc = full_report(os.path.join(g, "s3_sc_tile2.ncu-rep"), "r01 sc_scan_tile_kernel (register-blocked TMA tile scan, FFMA2, balanced 64-row groups), 100k descriptors x 32 "
                         "queries (tools/sc_one.py 100000 32 tile)", os.path.join(p, "r01_sc_tile_kernel.md"))
        for d in sc:
            if "dram__bytes_read.sum" in d:
                traffic.setdefault("other", {})[d["kernel"]] = [to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"])]
    # ---- round 2 captures (scratch/job_profiles.sh) --------------------------------------------------------------------------------
    def total(d):
        return to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"])

    if have("r02_launches.csv"):
        import shutil

        launch_list(os.path.join(g, "r02_launches.csv"), "r02 launch list: bench.py --streams 512 --steps 1 --warmup 3 (the TIMED configuration: KITTI 1232x368, "
                    "512 streams, 8 LM groups, pyramids on the overlap stream)", os.path.join(p, "r02_launches.md"))
        rows = [r for r in csv.reader(open(os.path.join(g, "r02_launches.csv"))) if len(r) > 10]
        with open(os.path.join(p, "r02_launches.csv"), "w", newline="") as f:  # kernel, grid, block, ns — the raw list without host names / ids
            wr = csv.writer(f)
            h = rows[0]
            wr.writerow(["kernel", "grid", "block", "gpu__time_duration_ns"])
            for r in rows[1:]:
                wr.writerow([short(r[h.index("Kernel Name")]), r[h.index("Grid Size")], r[h.index("Block Size")], r[h.index("Metric Value")]])
    if have("r02_eval_bench.ncu-rep"):
        ev = full_report(os.path.join(g, "r02_eval_bench.ncu-rep"), "r02 eval_kernel inside bench.py at the TIMED configuration (512 streams, 8 LM groups: ~64-77 items "
                         "and ~320k template points per launch, every stream with its own pyramids)", os.path.join(p, "r02_eval_kernel_512streams.md"))
        tr = [total(d) for d in ev if "dram__bytes_read.sum" in d]
        if tr:
            traffic["pose_eval_dram_bytes_per_launch"] = sum(tr) / len(tr)
            traffic["source"] = ("profiles/r02_eval_kernel_512streams.md (mean over %d launches of the eval kernel captured inside bench.py --streams 512, "
                                 "the timed configuration)" % len(tr))
    if have("r02_eval128.ncu-rep"):
        full_report(os.path.join(g, "r02_eval128.ncu-rep"), "r02 eval_kernel, 128 items of 9.9k points in one launch (tools/one_eval.py kitti 128): pose and scale flavour",
                    os.path.join(p, "r02_eval_kernel_128items.md"))
    if have("r02_pyramid.ncu-rep"):
        py = full_report(os.path.join(g, "r02_pyramid.ncu-rep"), "r02 pyramid kernels inside bench.py --streams 512 (64 frames per launch)", os.path.join(p, "r02_pyramid_kernels.md"))
        for d in py:
            if "dram__bytes_read.sum" in d:
                traffic.setdefault("r02", {}).setdefault(d["kernel"], []).append(total(d))
    if have("r02_sc.ncu-rep"):
        sc = full_report(os.path.join(g, "r02_sc.ncu-rep"), "r02 Scan-Context kernels, 100k descriptors: streaming scan (Q = 1), tile scan (Q = 32), re-score + publish",
                         os.path.join(p, "r02_scan_context_kernels.md"))
        for d in sc:
            if "dram__bytes_read.sum" in d:
                traffic.setdefault("r02", {}).setdefault(d["kernel"], []).append(total(d))
    if have("r02_sc_q1.ncu-rep"):
        sc = full_report(os.path.join(g, "r02_sc_q1.ncu-rep"), "r02 Scan-Context kernels, 100k descriptors, ONE query: HBM-streaming scan + merge / re-score / publish",
                         os.path.join(p, "r02_scan_context_q1.md"))
        for d in sc:
            if "dram__bytes_read.sum" in d:
                traffic.setdefault("r02", {}).setdefault(d["kernel"] + " (Q=1)", []).append(total(d))
    json.dump(traffic, open(traffic_path, "w"), indent=1)
    print(json.dumps(traffic, indent=1))
