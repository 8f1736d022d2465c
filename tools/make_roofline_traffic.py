"""profiles/roofline_traffic.json from an ncu launch list of every GEMM launch of one search:
    ncu --profile-from-start off --clock-control none -k regex:gemm_sm100_2cta \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv \
        --log-file gpurun_out/gemm_traffic_raw.csv python tools/profile_step.py --precision fp16x3
    python tools/make_roofline_traffic.py gpurun_out/gemm_traffic_raw.csv fp16x3 > profiles/roofline_traffic.json
bench.py reads `dram_bytes_per_launch` as roofline.traffic (same launches as roofline.achieved averages over)."""
import csv
import json
import re
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0,
        "ms": 1e3, "msecond": 1e3}
path, prec = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr = next(r for r in rows if "Metric Name" in r)
ix = {h: i for i, h in enumerate(hdr)}
launches = {}
for r in rows:
    if r is hdr or len(r) != len(hdr) or not r[ix["ID"]].isdigit():
        continue
    m = re.search(r"gemm_sm100_2cta_kernel<[^>]*>", r[ix["Kernel Name"]])
    d = launches.setdefault(int(r[ix["ID"]]), {"kernel": re.sub(r"\(int\)|\s", "", m.group(0)) if m else "?"})
    d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * UNIT.get(r[ix["Metric Unit"]], 1)
ls = list(launches.values())
rd = sum(x["dram__bytes_read.sum"] for x in ls)
wr = sum(x["dram__bytes_write.sum"] for x in ls)
by = {}
for x in ls:
    b = by.setdefault(x["kernel"], {"launches": 0, "bytes": 0.0, "us": 0.0})
    b["launches"] += 1
    b["bytes"] += x["dram__bytes_read.sum"] + x["dram__bytes_write.sum"]
    b["us"] += x["gpu__time_duration.sum"]
out = {prec: {
    "dram_bytes_per_launch": (rd + wr) / len(ls), "launches": len(ls), "dram_read_bytes_per_search": rd,
    "dram_write_bytes_per_search": wr,
    "basis": f"the SAME {len(ls)} GEMM launches of one search (BASELINE configs[1], batch 256) that bench.py's "
             "roofline.achieved averages over; the layer norms run inside the o / co / wo epilogues (NormFold), so "
             "those launches also read the old residual rows and write the new ones and the next GEMM's operand planes",
    "note": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
            "-k regex:gemm_sm100_2cta over one whole search (tools/make_roofline_traffic.py); raw list: "
            "profiles/r02_gemm_traffic_all_raw.csv. DRAM bytes ~= algorithmic operand+output bytes (no wasted HBM "
            "re-reads); the re-read happens L2->SM, see profiles/r02_gemm_tail_fold_ncu.json (l2_to_sm)",
    "by_variant": {k: {"launches": v["launches"], "dram_bytes_per_launch": v["bytes"] / v["launches"],
                       "avg_us": v["us"] / v["launches"]} for k, v in by.items()},
    "per_launch_tail_layer": "profiles/r02_gemm_tail_fold_ncu.json (ncu --set full, the six GEMMs of one decoder "
                             "layer of the forced tail, M=71910)"}}
print(json.dumps(out, indent=1))
