"""Turn the ncu captures under gpurun_out/ into the tracked summaries under profiles/ (run in the build container).

    python scripts/summarize_profiles.py r01

writes profiles/<round>_launches.md (every kernel launch of one timed step with its share of the step),
profiles/<round>_<kernel>.md (the `--set full` key counters of the top kernels) and profiles/traffic.json
(DRAM bytes per captured launch, read by bench.py for roofline.traffic).
"""
import collections, csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
KEYS = [
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_issue_stalled_barrier.pct",
]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    return [{h: (v, u) for h, u, v in zip(hdr, units, r)} for r in vals]


def launches(rnd):
    path = os.path.join(SRC, "launches.csv")
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    order = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0].replace("void ", "")[:70]
        ns = float(row["Metric Value"].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
        order.append((name, row["Grid Size"], row["Block Size"], ns))
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(OUT, f"{rnd}_launches.md"), "w") as fh:
        fh.write(f"# {rnd}: every kernel launch of ONE hybrid step (cudaProfilerStart/Stop bracket in bench.py --ncu-range)\n\n"
                 "`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none`; durations are cold-cache and\n"
                 "serialised, so compare SHARES with bench.py's CUDA-event numbers, not absolutes.\n\n"
                 f"total {tot / 1e6:.2f} ms over {sum(a[0] for a in agg.values())} launches\n\n| kernel | launches | ms | share |\n|---|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            fh.write(f"| `{k}` | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / tot:.1f}% |\n")
        fh.write("\n## launch list (in order)\n\n| # | kernel | grid | block | ms |\n|---:|---|---|---|---:|\n")
        for i, (n, g, b, ns) in enumerate(order):
            fh.write(f"| {i} | `{n}` | {g} | {b} | {ns / 1e6:.3f} |\n")
    print("wrote launches", tot / 1e6, "ms")


def kernels(rnd):
    traffic = {}
    reps = {"r01": [("sparse_tile_f64", "prof_sparse_f64"), ("sparse_tile_f32", "prof_sparse_f32"),
                    ("dense_filter_gemm", "prof_dense"), ("maxsim", "prof_maxsim")],
            "r02": [("sparse_tile_f64", "prof_r02_sparse_f64"), ("dense_filter_gemm", "prof_r02_dense"), ("maxsim", "prof_r02_maxsim"),
                    ("splade_head_gemm", "prof_r02_head_gemm"), ("splade_tail_codes", "prof_r02_tail_codes"),
                    ("splade_rescore", "prof_r02_rescore")],
            "r02b": [("dense_filter_gemm", "prof_r02b_dense")],
            "r02f": [("splade_head_gemm", "prof_r02f_head_gemm")]}
    for tag, rep in reps.get(rnd, []):
        path = os.path.join(SRC, rep + ".ncu-rep")
        if not os.path.exists(path):
            continue
        for r in raw(path)[:1]:
            name = r.get("Kernel Name", ("?", ""))[0]
            with open(os.path.join(OUT, f"{rnd}_{tag}.md"), "w") as fh:
                fh.write(f"# {rnd}: `ncu --set full --clock-control none` of one launch of {tag}\n\nkernel: `{name[:120]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
                for k in KEYS:
                    if k in r:
                        fh.write(f"| {k} | {r[k][0]} | {r[k][1]} |\n")
            rd = float(r["dram__bytes_read.sum"][0].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}.get(r["dram__bytes_read.sum"][1], 1)
            wr = float(r["dram__bytes_write.sum"][0].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}.get(r["dram__bytes_write.sum"][1], 1)
            dur = float(r["gpu__time_duration.sum"][0].replace(",", "")) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1}.get(r["gpu__time_duration.sum"][1], 1)
            traffic[tag] = {"dram_bytes": rd + wr, "duration_s_under_ncu": dur, "grid": r.get("launch__grid_size", ("", ""))[0],
                            "round": rnd, "launch": f"largest round of one step (ncu -s index in scripts/gpu_profile{'_r02' if rnd != 'r01' else ''}.sh)"}
            print(tag, traffic[tag])
    if traffic:
        tj = os.path.join(OUT, "traffic.json")
        old = json.load(open(tj)) if os.path.exists(tj) else {}
        old.update(traffic)            # a follow-up round re-captures only the kernels that changed
        json.dump(old, open(tj, "w"), indent=1)


if __name__ == "__main__":
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    launches(rnd)
    kernels(rnd)
