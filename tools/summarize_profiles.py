#!/usr/bin/env python
"""Developer tool: turn the ncu outputs of tools/profile_round.sh (gpurun_out/) into the summaries kept under profiles/.
usage: python tools/summarize_profiles.py <tag> [layers_json]"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
rows = list(csv.DictReader(l for l in open(f"{G}/launches_{tag}.csv") if l.startswith('"')))
L = collections.OrderedDict()
for r in rows:
    d = L.setdefault(int(r["ID"]), {"name": r["Kernel Name"].split("(")[0].replace("void ", ""), "grid": r["Grid Size"], "block": r["Block Size"]})
    v, u, m = float(r["Metric Value"].replace(",", "")), r["Metric Unit"], r["Metric Name"]
    if m == "gpu__time_duration.sum":
        d["us"] = v / 1000 if u in ("nsecond", "ns") else (v if u in ("usecond", "us") else v * 1000)
    else:
        d[m] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
ids = sorted(L)
for d in L.values():
    d["name"] = d["name"].split("::")[-1]  # (kernels in anonymous namespaces print as "unnamed>::name")
starts = [k for k, i in enumerate(ids) if L[i]["name"].startswith("conv0") or L[i]["name"].startswith("conv_stem")]
per_fwd = starts[1] - starts[0]
step = lambda k0: [L[ids[k]] for k in range(k0, k0 + per_fwd)]
seg = step(starts[3]) + step(starts[4])  # bench.py: 3 warm-up forwards, then the two timed device-resident steps
agg = collections.OrderedDict()
for d in seg:
    a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0]); a[0] += 1; a[1] += d["us"]
    a[2] += d.get("dram__bytes_read.sum", 0); a[3] += d.get("dram__bytes_write.sum", 0)
tot = sum(a[1] for a in agg.values())
out = ["# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none : python bench.py --steps 2 --warmup 3 --quick --no-parity",
       f"# the two timed device-resident steps (2 x {per_fwd} kernel launches), full-416-80cls, batch 64, 1 B200; cold-cache serialised times: compare SHARES"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{k:<30} n={a[0]:4d} {a[1]:10.1f} us {100 * a[1] / tot:5.1f}%   dram read {a[2] / 2e6:9.1f} MB/step  write {a[3] / 2e6:9.1f} MB/step")
out.append(f"total {tot:.1f} us over 2 steps = {tot / 2000:.3f} ms per step under ncu (CUDA-event time of the same step, not under ncu: see the bench line)")
conv = [d for d in step(starts[4]) if "conv" in d["name"]]
rd = sum(d.get("dram__bytes_read.sum", 0) for d in conv); wr = sum(d.get("dram__bytes_write.sum", 0) for d in conv)
alg = None
if len(sys.argv) > 2:
    alg = sum(r["bytes"] for r in json.load(open(sys.argv[2]))["rows"])
out.append(f"conv stack ({len(conv)} launches) DRAM traffic per step under ncu (cold caches per launch): read {rd / 1e6:.1f} MB + write {wr / 1e6:.1f} MB = {(rd + wr) / 1e6:.1f} MB"
           + (f"; algorithmic (every tensor once per use): {alg / 1e6:.1f} MB" if alg else ""))
open(f"{P}/{tag}_ncu_launch_summary.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out))
with open(f"{P}/{tag}_ncu_step_launches.txt", "w") as f:
    f.write("# one timed step, launch by launch: kernel, grid, block, us, dram read MB, dram write MB\n")
    for d in step(starts[4]):
        f.write(f"{d['name']:<30} {d['grid']:<14} {d['block']:<14} {d['us']:8.1f} {d.get('dram__bytes_read.sum', 0) / 1e6:8.1f} {d.get('dram__bytes_write.sum', 0) / 1e6:8.1f}\n")
json.dump({"conv_stack_dram_bytes_per_step_bs64": rd + wr, "read": rd, "write": wr, "algorithmic_bytes_per_step": alg,
           "note": "ncu flushes the caches before every profiled launch, so each layer re-reads from DRAM what the previous one left in the 126 MB L2; in the un-profiled step part of that traffic never reaches DRAM",
           "source": f"profiles/{tag}_ncu_step_launches.txt (dram__bytes_read.sum + dram__bytes_write.sum, summed over the conv launches of one timed step)"},
          open(f"{P}/roofline_traffic.json", "w"), indent=1)
# full captures -> raw metric summaries
for rep, outname, what in ((f"{G}/prof_convtc_{tag}.ncu-rep", f"{P}/{tag}_ncu_conv_tc_full.txt", "consecutive conv_tc launches of a timed step: a 52x52 strip 3x3 layer and the 1x1 layer after it"),
                           (f"{G}/prof_stemblock_{tag}.ncu-rep", f"{P}/{tag}_ncu_stem_block_full.txt", "conv_stem_kernel (conv1 + conv2) and conv_block_kernel (conv3 + conv4 + residual) of one pass")):
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units = r[0], r[1]
    want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_requests.sum", "lts__t_sectors.sum",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.max", "sm__cycles_active.avg", "smsp__inst_executed.sum"]
    lines = [f"# ncu --set full --clock-control none --import-source on (tools/profile_round.sh) : python bench.py --steps 2 --warmup 3 --quick --no-parity",
             f"# {what} (full-416-80cls, batch 64)"]
    for j, h in enumerate(hdr):
        if h in want:
            lines.append(f"{h:<80} {units[j]:<14} {[x[j][:60] for x in r[2:]]}")
    open(outname, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[2:]))
