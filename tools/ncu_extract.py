"""Summarise an ncu --set full report (read here with `ncu -i ... --page raw --csv`) into a few per-kernel numbers:
duration, DRAM bytes, L2->SM bytes, tensor-pipe activity, issue activity, occupancy."""
import csv
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm",
    # tensor-pipe activity (tcgen05 MMAs are counted on the hmma sub-pipe by this ncu build)
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
    "TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg": "tensor_hmma_cycles",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active": "hmma_inst_pct",
    "lts__t_bytes.sum": "l2_bytes",
    "sm__cycles_elapsed.max": "sm_cycles",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "launch__grid_size": "grid",
    "launch__registers_per_thread": "regs",
}


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        d = {"kernel": r[idx["Kernel Name"]].split("(")[0].split("::")[-1][:60]}
        for k, name in WANT.items():
            if k in idx and r[idx[k]] != "":
                u = units[idx[k]]
                if "byte" in u:
                    d[name] = to_bytes(r[idx[k]], u)
                else:
                    try:
                        v = float(r[idx[k]].replace(",", ""))
                    except ValueError:               # "no data" for a counter this capture did not collect
                        continue
                    if name == "us":
                        v = v / 1e3 if u in ("ns", "nsecond") else v
                    d[name] = v
        out.append(d)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
