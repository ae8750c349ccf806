"""Turn an .ncu-rep (ncu --set full) into a small JSON summary for profiles/."""
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max"]

UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def summarise(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        item = {"kernel": d.get("Kernel Name", "")[:120]}
        for k in KEYS:
            if k in d:
                try:
                    v = float(d[k].replace(",", ""))
                except ValueError:
                    continue
                if u[k] in UNIT:
                    v *= UNIT[u[k]]
                    item[k + " [bytes]"] = v
                else:
                    item[k + f" [{u[k]}]"] = v
        rd, wr = item.get("dram__bytes_read.sum [bytes]"), item.get("dram__bytes_write.sum [bytes]")
        if rd is not None and wr is not None:
            item["dram_traffic_bytes"] = rd + wr
        res.append(item)
    return res


if __name__ == "__main__":
    print(json.dumps({rep: summarise(rep) for rep in sys.argv[1:]}, indent=1))
