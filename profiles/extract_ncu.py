"""Dump the metrics the roofline discussion uses from `ncu --set full` reports
(read here on the CPU box with `ncu -i ... --page raw --csv`)."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__cluster_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum", "sm__cycles_elapsed.max"]


def to_json(out_path, paths):
    """profiles/ncu_traffic.json: per kernel (template arguments stripped) the per-launch DRAM traffic
    (dram__bytes_read.sum + dram__bytes_write.sum) and a few companions of the FIRST captured launch -- what bench.py
    reports as roofline.traffic."""
    import json
    import re
    res = {}
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]

        def val(r, key):
            if key not in hdr:
                return None
            i = hdr.index(key)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                return None
            u = units[i].lower()
            scale = {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "byte": 1.0, "us": 1.0, "usecond": 1.0, "ms": 1e3,
                     "msecond": 1e3, "ns": 1e-3, "nsecond": 1e-3, "s": 1e6, "second": 1e6}.get(u, 1.0)
            return v * scale
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "").strip()
            name = re.sub(r"<[^<>]*>$", "", name).split("::")[-1]          # no template arguments, no namespaces
            if name in res:
                continue
            rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
            res[name] = {"dram_bytes_per_launch": None if rd is None or wr is None else rd + wr,
                         "dram_bytes_read": rd, "dram_bytes_write": wr,
                         "duration_us_under_ncu": val(r, "gpu__time_duration.sum"),
                         "tensor_pipe_pct_of_peak_active":
                             val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                         "grid_size": val(r, "launch__grid_size"), "cluster_size": val(r, "launch__cluster_size"),
                         "source": "ncu --set full --clock-control none, " + path.split("/")[-1]}
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    print("wrote %s: %s" % (out_path, ", ".join(sorted(res))))


def main(paths):
    if paths and paths[0] == "--json":
        return to_json(paths[1], paths[2:])
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        print("== %s" % path.split("/")[-1])
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
            print("  kernel: %s" % name[:110])
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    print("    %-70s %s %s" % (w, r[i], units[i]))


if __name__ == "__main__":
    main(sys.argv[1:])
