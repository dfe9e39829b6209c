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


def main(paths):
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
