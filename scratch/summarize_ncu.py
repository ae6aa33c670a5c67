"""Turn gpurun_out/*.ncu-rep + launch lists into the small text/CSV/JSON summaries committed under profiles/."""
import csv, io, json, re, subprocess, sys, collections, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return {k: (rows[1][i], rows[2][i]) for i, k in enumerate(rows[0])}

def source(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[1]
    ie, ist = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    ops, stall = collections.Counter(), collections.Counter()
    for r in rows[2:]:
        try: n = int(r[ie])
        except Exception: continue
        t = r[1].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] += n; stall[op] += int(r[ist] or 0)
    return rows[0][1], ops, stall

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]

def summarize(rep, name, note):
    d = raw(rep)
    kname, ops, stall = source(rep)
    lines = [f"# ncu --set full --clock-control none --import-source on, one launch; {note}", f"kernel: {kname}"]
    for k in KEYS:
        if k in d: lines.append(f"{k:75s} {d[k][1]:>16s} {d[k][0]}")
    st = sorted(((float(v[1].replace(',', '')), k) for k, v in d.items()
                 if k.startswith("smsp__pcsamp_warps_issue_stalled") and not k.endswith("_not_issued") and v[1] not in ("", "n/a")), reverse=True)
    lines.append("warp stall samples (top): " + ", ".join(f"{k.replace('smsp__pcsamp_warps_issue_stalled_', '')}={int(x)}" for x, k in st[:8]))
    tot = sum(ops.values())
    lines.append(f"executed warp instructions {tot}; top opcodes: " + ", ".join(f"{k}={v}" for k, v in ops.most_common(14)))
    sass = " ".join(ops)
    lines.append("SASS evidence: " + ", ".join(f"{m}={'yes' if any(o.startswith(m) for o in ops) else 'no'}" for m in ("UTCIMMA", "UTMALDG", "UBLKCP", "LDTM", "SYNCS", "UTCBAR")))
    open(os.path.join(ROOT, "profiles", f"ncu_{name}_{TAG}_summary.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:6]))

def launch_list(path, out_csv, out_json):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]; kn, mn, mv, mu, gi, idc = (h.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "Grid Size", "ID"))
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv: continue
        e = per.setdefault(r[idc], {"kernel": r[kn], "grid": r[gi]})
        v = float(r[mv].replace(",", ""))
        u = r[mu]
        if r[mn] == "gpu__time_duration.sum": e["us"] = v / 1000 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000)
        else:
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            e[r[mn]] = v * mult
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    with open(out_csv, "w") as f:
        f.write("id,kernel,grid,us,dram_read_bytes,dram_write_bytes\n")
        for i, e in per.items():
            short = re.sub(r"\(.*", "", e["kernel"]).replace("void ", "")[:90]
            f.write(f"{i},\"{short}\",\"{e['grid']}\",{e.get('us', 0):.2f},{e.get('dram__bytes_read.sum', 0):.0f},{e.get('dram__bytes_write.sum', 0):.0f}\n")
            key = re.sub(r"<.*", "", short)
            a = agg[key]; a[0] += 1; a[1] += e.get("us", 0); a[2] += e.get("dram__bytes_read.sum", 0); a[3] += e.get("dram__bytes_write.sum", 0)
    tot = sum(a[1] for a in agg.values())
    summ = {k: {"launches": a[0], "us": round(a[1], 1), "share": round(a[1] / tot, 4), "dram_read_MB": round(a[2] / 1e6, 1), "dram_write_MB": round(a[3] / 1e6, 1),
                "dram_bytes_per_launch": round((a[2] + a[3]) / a[0])} for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])}
    json.dump({"total_us": round(tot, 1), "kernels": summ}, open(out_json, "w"), indent=1)
    for k, v in list(summ.items())[:8]: print(k, v)

if __name__ == "__main__":
    G = os.path.join(ROOT, "gpurun_out")
    launch_list(os.path.join(G, "launches_church_dram.csv"), os.path.join(ROOT, "profiles", f"launches_{TAG}_church_b100_eager_step_final.csv"),
                os.path.join(ROOT, "profiles", f"launches_{TAG}_church_b100_summary.json"))
    if os.path.exists(os.path.join(G, "launches_imagenet_dram.csv")):
        launch_list(os.path.join(G, "launches_imagenet_dram.csv"), os.path.join(ROOT, "profiles", f"launches_{TAG}_imagenet_b128_eager_step_final.csv"),
                    os.path.join(ROOT, "profiles", f"launches_{TAG}_imagenet_b128_summary.json"))
    summarize(os.path.join(G, "qgemm_c192.ncu-rep"), "qgemm", "church 32x32 conv 3x3, 192->192 channels, batch 100 (M=102400, N=192, K=1728)")
    summarize(os.path.join(G, "qattn_t1024.ncu-rep"), "qattn", "church attention, 800 (batch x heads) x 1024 tokens x 24 channels")
    summarize(os.path.join(G, "actq_tma.ncu-rep"), "actq", "activation producer [128,192,64,64] fp32 -> u8 NHWC codes + halo")
