"""Turn the raw ncu artefacts of a gpurun call (gpurun_out/) into the small text summaries committed under profiles/.

    python scripts/summarize_profiles.py <tag>      e.g. r01_v1
Writes profiles/<tag>_launches.txt (per-kernel share of the launch list) and profiles/<tag>_<kernel>_ncu.txt (key
metrics of every kernel in gpurun_out/prof_*.ncu-rep)."""
import collections, csv, glob, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "rXX"
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)

lp = os.path.join(OUT, "launches.csv")
if os.path.exists(lp):
    lines = [l for l in open(lp) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
        a = agg.setdefault(row["Kernel Name"][:110], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none : per-kernel totals over the captured launches ({tot:.0f} us)\n")
        f.write("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"{t:10.1f} us {n:5d}x  avg {t / n:9.1f} us  {100 * t / tot:5.1f}%  {k}\n")
        # per-update view: the kernels of ONE gradient update are those launched a whole number of times per td_mse_kernel launch (one per
        # update) inside the capture window -- ALL of them, library kernels included, so that their absence from the captured update is
        # visible (the window also holds the eval round, the roofline micro-benchmarks and setup kernels, whose counts are not multiples)
        n_upd = next((n for k, (n, t) in agg.items() if "td_mse_kernel" in k), 0)
        if n_upd:
            step = [(k, n // n_upd, t / n * (n // n_upd)) for k, (n, t) in agg.items() if n >= n_upd and n % n_upd == 0 and "envelope_td" not in k]
            step += [(k, 1, t / n) for k, (n, t) in agg.items() if "envelope_td" in k][:1]
            tot_s = sum(t for _, _, t in step)
            lib = [k for k, _, _ in step if not ("morl::" in k)]
            f.write(f"\n# one gradient update ({n_upd} captured): every kernel launched an exact multiple of {n_upd} times, {tot_s:.0f} us under ncu\n")
            f.write(f"# kernels of the update that are NOT repo kernels (ATen / cuBLAS / cutlass): {lib if lib else 'none'}\n")
            for k, c, t in sorted(step, key=lambda x: -x[2])[:40]:
                f.write(f"{t:10.1f} us {c:5d}x per update  {100 * t / tot_s:5.1f}%  {k}\n")
    print("wrote launches summary")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]
for rep in glob.glob(os.path.join(OUT, "prof_*.ncu-rep")):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    name = os.path.basename(rep).replace("prof_", "").replace(".ncu-rep", "")
    traffic = []
    with open(os.path.join(ROOT, "profiles", f"{tag}_{name}_ncu.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none, {os.path.basename(rep)} (cold-cache single launches)\n")
        for r in rows[2:]:
            f.write(f"--- {r[idx['Kernel Name']][:120]}\n")
            for w in WANT:
                if w in idx:
                    f.write(f"{w} = {r[idx[w]]} {units[idx[w]]}\n")
            stall = sorted(((float(r[idx[h]].replace(',', '') or 0), h.replace('smsp__pcsamp_warps_issue_stalled_', '')) for h in hdr
                            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h), reverse=True)[:6]
            f.write("top stall reasons (pc samples): " + ", ".join(f"{h}={int(v)}" for v, h in stall) + "\n")
            def b(key):
                v = float(r[idx[key]].replace(",", "")); u = units[idx[key]]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            if "dram__bytes_read.sum" in idx:
                traffic.append((r[idx["Kernel Name"]][:60], b("dram__bytes_read.sum") + b("dram__bytes_write.sum")))
    if name == "envelope" and traffic:
        json.dump({"kernel": traffic[0][0], "dram_bytes_per_launch": traffic[0][1], "source": f"profiles/{tag}_{name}_ncu.txt"},
                  open(os.path.join(ROOT, "profiles", "envelope_td_traffic.json"), "w"))
    if name == "gemm" and traffic:
        k = [t for t in traffic if "gemm_planes_kernel" in t[0]] or traffic
        json.dump({"kernel": k[0][0], "dram_bytes_per_launch": k[0][1], "source": f"profiles/{tag}_{name}_ncu.txt"},
                  open(os.path.join(ROOT, "profiles", "gemm_traffic.json"), "w"))
    if name == "chain" and traffic:
        k = [t for t in traffic if "gemm_chain_kernel" in t[0]] or traffic
        json.dump({"kernel": k[0][0], "dram_bytes_per_launch": k[0][1], "source": f"profiles/{tag}_{name}_ncu.txt"},
                  open(os.path.join(ROOT, "profiles", "gemm_chain_traffic.json"), "w"))
    if name == "qhead" and traffic:
        k = [t for t in traffic if "qhead_envelope_kernel" in t[0]] or traffic
        json.dump({"kernel": k[0][0], "dram_bytes_per_launch": k[0][1], "source": f"profiles/{tag}_{name}_ncu.txt"},
                  open(os.path.join(ROOT, "profiles", "qhead_envelope_traffic.json"), "w"))
    print("wrote", name)
