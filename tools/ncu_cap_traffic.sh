#!/bin/bash
# One `ncu --set full` capture of the cap-forward chain as bench.py runs it (python bench.py --cap-only), summarised into
# profiles/ncu_cap_forward_traffic.json (what bench.py's roofline.traffic reads) and profiles/ncu_cap_forward_<tag>.md.
#   gpurun -- 'tools/ncu_cap_traffic.sh r02'        (run on the GPU box; ~1 minute)
set -u
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
# skip the warm-up + capture launches: profile the LAST graph replays only (4 kernels per chain)
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"cap_route|cap_hop_e1|cap_hop_ev|cap_recon_hop|cap_recon_proj|gproj2_fwd" \
    --launch-skip 24 --launch-count 6 -f -o $O/prof_cap_fwd_$TAG python bench.py --cap-only > $O/ncu_cap_$TAG.log 2>&1
ncu -i $O/prof_cap_fwd_$TAG.ncu-rep --page raw --csv > $O/ncu_cap_$TAG.csv 2>/dev/null
python - "$O/ncu_cap_$TAG.csv" "$TAG" <<'PY'
import csv, json, subprocess, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
ix = {k: hdr.index(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                               "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
                               "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}
units = rows[1]
def to_bytes(v, u):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
ks = []
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0]
    ks.append({"kernel": name, "us": float(r[ix["gpu__time_duration.sum"]]),
               "dram_read": to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]),
               "dram_write": to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]]),
               "tensor_pct": float(r[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]),
               "warps_active_pct": float(r[ix["sm__warps_active.avg.pct_of_peak_sustained_active"]]),
               "dram_pct": float(r[ix["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]])})
names = []
for k in ks:
    if k["kernel"] in names:
        break
    names.append(k["kernel"])
per_chain = len(names)
chains = len(ks) // per_chain
tot = sum(k["dram_read"] + k["dram_write"] for k in ks[:per_chain * chains]) / chains
try:
    commit = subprocess.check_output(["git", "rev-parse", "--short", "HEAD"], text=True).strip()
except Exception:
    commit = "unknown"
out = {"workload": "pems08", "batch": 64, "dram_bytes_per_chain": tot, "kernels_per_chain": names, "chains_profiled": chains,
       "source": f"ncu --set full of `python bench.py --cap-only` (tools/ncu_cap_traffic.sh {sys.argv[2]}), commit {commit}"}
json.dump(out, open("gpurun_out/ncu_cap_forward_traffic.json", "w"), indent=1)
with open(f"gpurun_out/ncu_cap_forward_{sys.argv[2]}.md", "w") as f:
    f.write(f"# ncu --set full of the cap-forward chain (`python bench.py --cap-only`), {sys.argv[2]}, commit {commit}\n\n")
    f.write(f"{chains} chains of {per_chain} kernels profiled; DRAM read+write per chain: {tot/1e6:.1f} MB (algorithmic 72.07 MB)\n\n")
    f.write("| kernel | us | dram read MB | dram write MB | DRAM % | tensor % | warps active % |\n|---|---|---|---|---|---|---|\n")
    for k in ks[:per_chain]:
        f.write(f"| `{k['kernel']}` | {k['us']:.1f} | {k['dram_read']/1e6:.1f} | {k['dram_write']/1e6:.1f} | {k['dram_pct']:.1f} | {k['tensor_pct']:.1f} | {k['warps_active_pct']:.1f} |\n")
print(json.dumps(out))
PY
