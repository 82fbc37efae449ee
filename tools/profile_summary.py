"""Turn ncu outputs under gpurun_out/ into the small, committed summaries under profiles/.
usage: python tools/profile_summary.py <tag> <launches.csv> <full.ncu-rep> [A0|A1]"""
import collections, csv, json, os, subprocess, sys

tag, launches, rep = sys.argv[1:4]
out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
rows = list(csv.reader(l for l in open(launches) if l.startswith('"')))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    name = r[ki].split("(")[0].replace("void ", "").strip()[:70]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v for _, v in agg.values())
lines = [f"# {tag}: every launch of `ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 3`",
         "# (cold-cache, serialised: compare SHARES, not absolutes)", "kernel,launches,total_ms,share_pct"]
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{k},{n},{v / 1e6:.3f},{100 * v / tot:.2f}")
open(os.path.join(out_dir, f"{tag}_launch_shares.csv"), "w").write("\n".join(lines) + "\n")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
units = rr[1]
summ = []
for r in rr[2:]:
    d = {}
    for w in want:
        if w in h:
            i = h.index(w)
            d[w] = r[i] + (" " + units[i] if units[i] and w != "Kernel Name" else "")
    summ.append(d)
json.dump(summ, open(os.path.join(out_dir, f"{tag}_ncu_full_summary.json"), "w"), indent=1)
# bench.py's roofline.traffic: DRAM bytes of the FIRST captured forward chain launch (inference-mode fine pass)
def gb(x):
    v, u = x.split()
    return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
# bench.py's roofline traffic figures: DRAM bytes of the FIRST captured launch of the forward chain kernel
# (inference-mode fine pass), of the training forward and of the fused backward.  argv[4] = architecture tag (A0 | A1)
if len(sys.argv) > 4:
    path = os.path.join(out_dir, "r2_ncu_summary.json")
    cur = json.load(open(path)) if os.path.exists(path) else {}
    for needle, key in (("mlp_fwd_tc_kernel<0>", f"mlp_fwd_tc_{sys.argv[4]}_dram_bytes"),
                        ("mlp_fwd_tc_kernel<1>", f"mlp_fwd_train_tc_{sys.argv[4]}_dram_bytes"),
                        ("mlp_bwd_tc_kernel", f"mlp_bwd_tc_{sys.argv[4]}_dram_bytes")):
        for d in summ:
            if needle in d.get("Kernel Name", "").replace("(bool)", ""):
                cur[key] = gb(d["dram__bytes_read.sum"]) + gb(d["dram__bytes_write.sum"])
                break
    json.dump(cur, open(path, "w"), indent=1)
print(open(os.path.join(out_dir, f"{tag}_launch_shares.csv")).read()[:1500])
for d in summ:
    print({k.split(".")[0][-28:]: v for k, v in d.items()})
