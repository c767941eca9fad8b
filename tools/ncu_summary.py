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


def main(path, top=None, traffic_json=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print("ncu report %s: %d profiled launches" % (path, len(rows) - 2))
    body = rows[2:]
    if top:       # the `top` longest launches only (a whole step has hundreds of tiny ones)
        di = col["gpu__time_duration.sum"]
        body = sorted(body, key=lambda r: -float(r[di].replace(",", "") or 0) * SCALE.get(units[di], 1.0))[:top]
    traffic = []
    for r in body:
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
            traffic.append(dict(kernel=name[:60], us=t * 1e6, dram_read=rd, dram_write=wr))
        for k in KEYS[3:]:
            if k in vals:
                print("   %-66s %12.2f" % (k, vals[k]))


    if traffic_json and traffic:
        import json
        # tools/profile_round.sh captures launches 7..12 of tools/ncu_conv_case.py (`-s 6 -c 6`): the six 512->256 launches
        # are skipped, what is profiled is (fwd, dgrad, wgrad) of the 128->128 3x3x3 layer on the 200x200x16 grid, twice
        names = ["fwd 128->128", "dgrad 128->128", "wgrad 128->128", "fwd 128->128 (2nd)", "dgrad 128->128 (2nd)",
                 "wgrad 128->128 (2nd)"]
        per = {n: t for n, t in zip(names, traffic)}
        ref = per.get("fwd 128->128", traffic[0])
        note = "ncu --set full, tc_conv_kernel 3x3x3 128->128 on the 200x200x16 grid, bf16 in/out, DRAM read+write MB per launch: " + \
            "; ".join("%s %.1f (%.0f us)" % (n, (t["dram_read"] + t["dram_write"]) / 1e6, t["us"]) for n, t in per.items()) + \
            ".  Algorithmic bytes: 327.7 MB (x read once + y written once, bf16).  `bytes` = fwd 128->128."
        with open(traffic_json, "w") as f:
            json.dump(dict(bytes=ref["dram_read"] + ref["dram_write"], note=note, per_launch=per), f, indent=1)


if __name__ == "__main__":
    a = sys.argv[1:]
    top = int(a[a.index("--top") + 1]) if "--top" in a else None
    tj = a[a.index("--traffic-json") + 1] if "--traffic-json" in a else None
    main(a[0], top, tj)
