"""Turn the ncu CSVs / reports brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarise_profiles.py            # after the measurement pass described in profiles/README.txt
"""
import collections, csv, json, re, shutil, subprocess, sys
ROOT = __file__.rsplit("/", 2)[0]
G, P = ROOT + "/gpurun_out/", ROOT + "/profiles/"


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    data = [r for r in rows[hdr + 1:] if len(r) == len(H) and r[0].isdigit()]
    return data, {h: i for i, h in enumerate(H)}


def launches():
    data, ci = load(G + "launches_final.csv")
    half = data[len(data) // 2:]
    agg = collections.OrderedDict()
    for r in half:
        name = re.sub(r"\(.*", "", r[ci["Kernel Name"]]).replace("void ", "").replace("crdr::", "")
        v, u = float(r[ci["Metric Value"]]), r[ci["Metric Unit"]]
        ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms
    tot = sum(a[1] for a in agg.values())
    conv = sum(a[1] for k, a in agg.items() if k.startswith("conv_tcgen05"))
    lines = ["ncu launch list of ONE encode+decode step (24 x 512x768, q=1.5, beta=3.84), round 1 final kernels",
             "command: ncu --metrics gpu__time_duration.sum --clock-control none -k regex:<all crdr kernels> -c 2000 --csv "
             "python tools/one_step.py 24",
             "(second of the two iterations; per-launch times under ncu are serialised and cold-cache: compare SHARES with "
             "bench.py, not absolutes)",
             f"launches in the step: {len(half)}; sum of durations {tot:.2f} ms; conv_tcgen05_kernel share {100 * conv / tot:.1f} %", ""]
    for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k:60s} {c:4d} launches {ms:9.3f} ms  {100 * ms / tot:5.1f} %")
    open(P + "launches_r01_summary.txt", "w").write("\n".join(lines) + "\n")
    shutil.copy(G + "launches_final.csv", P + "launches_r01.csv")
    print("\n".join(lines[3:8]))


def dram():
    data, ci = load(G + "conv_dram_final.csv")
    by, ids = collections.defaultdict(float), set()
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3,
             "ms": 1, "msecond": 1}
    for r in data:
        by[r[ci["Metric Name"]]] += float(r[ci["Metric Value"]]) * scale.get(r[ci["Metric Unit"]], 1)
        ids.add(r[ci["ID"]])
    L = len(ids)
    per = (by["dram__bytes_read.sum"] + by["dram__bytes_write.sum"]) / L
    txt = ["ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:conv_tcgen05 -s 366 -c 366 "
           "python tools/one_step.py 24",
           f"conv_tcgen05_kernel launches in one encode+decode step (24 x 512x768): {L}",
           f"DRAM read  {by['dram__bytes_read.sum'] / 1e9:.2f} GB, write {by['dram__bytes_write.sum'] / 1e9:.2f} GB per step",
           f"-> traffic per launch (average) {per / 1e6:.1f} MB   (bench.py roofline.traffic)",
           f"sum of launch durations under ncu {by['gpu__time_duration.sum']:.1f} ms"]
    open(P + "conv_dram_traffic_r01.txt", "w").write("\n".join(txt) + "\n")
    shutil.copy(G + "conv_dram_final.csv", P + "conv_dram_r01.csv")
    print("\n".join(txt[1:4]))


def charm():
    raw = subprocess.run(["ncu", "-i", G + "charm_final.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    m = dict(zip(rows[0], zip(rows[1], rows[2])))
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second",
            "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
            "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "sm__inst_executed.avg.per_cycle_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active"]
    out = ["ncu --set full --import-source on --clock-control none -k conv_tcgen05_kernel<4,1,1,0> -s 200 -c 1 python tools/one_step.py 24",
           "(one ChARM 5x5 first-layer launch of the measured step; F16X3, tile_n 112, CTA pairs; round 1 final kernel)", ""]
    for k in keys:
        if k in m:
            out.append(f"{k:100s} {m[k][1]} {m[k][0]}")
    out += ["", "warp stall reasons (per issued instruction):"]
    for h in rows[0]:
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            v = float(m[h][1] or 0)
            if v >= 0.05:
                out.append(f"  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:.2f}")
    open(P + "ncu_prof_charm_r01.txt", "w").write("\n".join(out) + "\n")
    print("\n".join(o for o in out if "utchmma" in o or "time_duration" in o))


if __name__ == "__main__":
    launches(); dram(); charm()
    shutil.copy(G + "bench_final.json", P + "bench_r01.json")
    shutil.copy(G + "bench_ref_final.json", P + "bench_r01_reference_arm.json")
    shutil.copy(G + "layer_profile_cg2.log", P + "layer_profile_r01_batch24_final.log")
    d = json.load(open(P + "bench_r01.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
