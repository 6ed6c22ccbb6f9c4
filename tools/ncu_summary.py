"""Turns the Nsight Compute outputs of tools/profile_trip.sh into the small text / csv / json files kept under
profiles/ (run here, on the files gpurun brought back):

    python tools/ncu_summary.py launches gpurun_out/<tag>_ncu_launches.csv   > profiles/<tag>_ncu_launch_shares.txt
    python tools/ncu_summary.py kernels  gpurun_out/<tag>_ncu_full_*.ncu-rep > profiles/<tag>_ncu_kernels.csv
    python tools/ncu_summary.py traffic  gpurun_out/<tag>_ncu_traffic.csv qconv_cl_fprop_kernel > profiles/fprop_traffic.json
"""
import csv
import json
import subprocess
import sys


def read_ncu_csv(path):
    rows = list(csv.reader(open(path, newline="")))
    for i, r in enumerate(rows):
        if "Kernel Name" in r and "Metric Name" in r:
            return r, [x for x in rows[i + 1:] if len(x) == len(r)]
    raise SystemExit("no ncu csv header in " + path)


def short(name):
    name = name.replace("seldq::", "")
    return name.split("(")[0][-64:]


def launches(path):
    hdr, rows = read_ncu_csv(path)
    kn, mn, mv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    agg = {}
    for r in rows:
        if r[mn] != "gpu__time_duration.sum":
            continue
        a = agg.setdefault(short(r[kn]), [0.0, 0])
        a[0] += float(r[mv].replace(",", "")) / 1e3
        a[1] += 1
    total = sum(a[0] for a in agg.values())
    print("# one eager training step (+ 2 STFT calls) under ncu --metrics gpu__time_duration.sum --clock-control none")
    print("# cold-cache, serialised launches: shares, not absolutes, are comparable with the bench")
    print("total_us=%.1f launches=%d" % (total, sum(a[1] for a in agg.values())))
    for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
        print("%9.1f us %4d x %7.1f us %5.1f%%  %s" % (us, n, us / n, 100 * us / total, name))


KEYS = [("gpu__time_duration.sum", "duration_us", 1e-3),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_pct", 1),
        ("dram__bytes_read.sum", "dram_read_MB", 1), ("dram__bytes_write.sum", "dram_write_MB", 1),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct", 1),
        ("launch__registers_per_thread", "regs", 1), ("launch__grid_size", "grid", 1), ("launch__block_size", "block", 1),
        ("sm__inst_executed.sum", "warp_inst", 1)]


def kernels(paths):
    w = csv.writer(sys.stdout)
    first = True
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        cols = []
        for key, label, _ in KEYS:
            idx = [i for i, h in enumerate(hdr) if h == key] or [i for i, h in enumerate(hdr) if h.endswith(key)] or \
                [i for i, h in enumerate(hdr) if h.startswith(key)]
            cols.append(idx[0] if idx else None)
        tp = []
        if first:
            w.writerow(["kernel"] + [lab + (" [%s]" % units[c] if c is not None and units[c] else "") for (_, lab, _), c in zip(KEYS, cols)] +
                       [])
            first = False
        kn = hdr.index("Kernel Name")
        for r in rows[2:]:
            if len(r) != len(hdr):
                continue
            vals = [r[c] if c is not None else "" for c in cols]
            tpv = [float(r[i].replace(",", "")) for i in tp if r[i] not in ("", "n/a")]
            w.writerow([short(r[kn])] + vals)


def traffic(path, pattern):
    hdr, rows = read_ncu_csv(path)
    kn, mn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    per = {}
    for r in rows:
        if pattern not in r[kn]:
            continue
        key = r[hdr.index("ID")]
        d = per.setdefault(key, {})
        v = float(r[mv].replace(",", ""))
        u = r[mu].lower()
        if "byte" in u:
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        d[r[mn]] = v
    n = len(per)
    rd = sum(d.get("dram__bytes_read.sum", 0) for d in per.values())
    wr = sum(d.get("dram__bytes_write.sum", 0) for d in per.values())
    us = sum(d.get("gpu__time_duration.sum", 0) for d in per.values()) / 1e3
    print(json.dumps(dict(kernel=pattern, launches_per_step=n, dram_bytes_read_per_step=rd, dram_bytes_write_per_step=wr,
                          traffic_bytes_per_launch=(rd + wr) / max(1, n), ncu_duration_us_per_step=us,
                          how="ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
                              "--clock-control none -k regex:%s over one eager training step (tools/ncu_step.py)" % pattern), indent=1))


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "launches":
        launches(sys.argv[2])
    elif cmd == "kernels":
        kernels(sys.argv[2:])
    elif cmd == "traffic":
        traffic(sys.argv[2], sys.argv[3])
