"""Summarise an `ncu --set full` report (read here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [> profiles/rNN_x.summary.txt]
One block per profiled launch: duration, DRAM bytes / achieved GB/s / % of peak, tensor-pipe %, L2 %,
occupancy, registers."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3,
         "second": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3}


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print("ncu report %s: %d profiled launches" % (path, len(rows) - 2))
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        print("== %s  grid %s block %s" % (name[:100], r[col.get("Grid Size", 0)], r[col.get("Block Size", 0)]))
        vals = {}
        for k in KEYS:
            if k in col:
                try:
                    v = float(r[col[k]].replace(",", ""))
                except ValueError:
                    continue
                vals[k] = v * SCALE.get(units[col[k]], 1.0)
        t = vals.get("gpu__time_duration.sum")
        rd, wr = vals.get("dram__bytes_read.sum", 0.0), vals.get("dram__bytes_write.sum", 0.0)
        if t:
            print("   duration %.1f us   DRAM read %.2f MB  write %.2f MB  -> %.0f GB/s" % (t * 1e6, rd / 1e6, wr / 1e6, (rd + wr) / t / 1e9))
        for k in KEYS[3:]:
            if k in vals:
                print("   %-66s %12.2f" % (k, vals[k]))


if __name__ == "__main__":
    main(sys.argv[1])
