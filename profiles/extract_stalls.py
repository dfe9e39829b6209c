"""Warp-stall breakdown (PC sampling) and issue-slot utilisation of the kernels in `ncu --set full` reports."""
import csv
import subprocess
import sys


def main(paths):
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr = rows[0]
        print("== %s" % path.split("/")[-1])
        for r in rows[2:3]:
            name = r[hdr.index("Kernel Name")].replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
            print("  kernel: %s" % name[:100])
            for m in ("smsp__issue_active.avg.pct_of_peak_sustained_active",
                      "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum",
                      "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                      "smsp__warps_eligible.avg.per_cycle_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
                      "lts__t_sector_hit_rate.pct"):
                if m in hdr:
                    print("    %-68s %s" % (m, r[hdr.index(m)]))
            samples = {}
            for i, h in enumerate(hdr):
                if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                    try:
                        samples[h[len("smsp__pcsamp_warps_issue_stalled_"):]] = float(r[i].replace(",", ""))
                    except ValueError:
                        pass
            tot = sum(samples.values()) or 1.0
            print("    warp samples by stall reason (share of %d samples):" % tot)
            for k, v in sorted(samples.items(), key=lambda kv: -kv[1])[:8]:
                print("      %-24s %5.1f %%" % (k, 100.0 * v / tot))


if __name__ == "__main__":
    main(sys.argv[1:])
