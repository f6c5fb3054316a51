"""Copy one round's bench line, launch list and ncu summaries from gpurun_out/ into profiles/ and
recompute profiles/traffic.json.  usage: python tools/update_profiles.py TAG   (files gpurun_out/TAG_cfg{3,4,5}.txt,
gpurun_out/bench_TAG.json, gpurun_out/launches_TAG.csv)"""
import csv, json, re, shutil, sys, os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
rows = list(csv.reader(l for l in open(os.path.join(go, f"launches_{tag}.csv")) if l.startswith('"')))
col = {h: i for i, h in enumerate(rows[0])}
out = ["# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --also none",
       "# (cold-cache, serialised launches: compare shares, not absolutes). 5 steps x 2 resize launches (one per group of row bands) + 1 synthetic fill:",
       "# the timed steps consist only of resize_down_kernel launches.", "# id  kernel  grid  duration_us"]
for r in rows[1:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("picha_b200::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    out.append(f"{int(r[col['ID']]):3d}  {name:45s} {r[col['Grid Size']]:22s} {float(r[col['Metric Value']]) / 1e3:10.1f}")
open(os.path.join(pr, "r01_launches_bench_cfg3.txt"), "w").write("\n".join(out) + "\n")
cmds = {"cfg3": ("-s 6 -c 2  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --also none", "both launches of one timed cfg3 step (downscaling kernel, resize_down.cuh)"),
        "cfg5": ("-s 3 -c 1  python bench.py --workload cfg5 --steps 2 --warmup 3 --no-cpu --no-e2e --also none", "the single launch of one timed cfg5 step (downscaling kernel, 8-row groups)"),
        "cfg4": ("-s 15 -c 5  python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu --no-e2e --also none", "the 5 launches of one timed cfg4 step (upscaling kernel, resize_up.cuh)")}
traffic = {}
for w, (cmd, note) in cmds.items():
    txt = open(os.path.join(go, f"{tag}_{w}.txt")).read()
    open(os.path.join(pr, f"r01_ncu_{w}_final.txt"), "w").write(f"# ncu --set full --clock-control none --import-source on -k regex:resize_ {cmd}\n# {note}\n" + txt)
    tot = 0.0
    for m in re.finditer(r"dram__bytes_(?:read|write)\.sum\s+([0-9.]+) (Gbyte|Mbyte|Kbyte|byte)", txt):
        tot += float(m.group(1)) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[m.group(2)]
    traffic[w] = int(tot)
traffic["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum summed over the resize launches of one step, from the ncu --set full captures "
                    "profiles/r01_ncu_cfg{3,5,4}_final.txt. Algorithmic bytes per step: cfg3 9024307200, cfg5 6571425792, cfg4 2684354560 "
                    "(cfg4's writes are still partly in L2 when the launch ends, so its DRAM count is below the algorithmic figure).")
json.dump(traffic, open(os.path.join(pr, "traffic.json"), "w"), indent=1)
shutil.copy(os.path.join(go, f"bench_{tag}.json"), os.path.join(pr, "r01_bench_n1.json"))
d = json.loads(open(os.path.join(go, f"bench_{tag}.json")).read().strip().splitlines()[-1])
print({k: d[k] for k in ["value", "ms_per_step", "gpu_launches"]}, d["roofline"]["frac"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["parity_vs_gpu"])
for k, v in d.get("also", {}).items():
    print(k, {kk: vv for kk, vv in v.items() if kk in ("value", "ms_per_step", "roofline_frac", "achieved_GBs", "median_us", "cpu_reference_median_us")})
print(traffic)
