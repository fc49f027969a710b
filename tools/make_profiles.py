"""Turn the ncu artefacts gpurun brought back (gpurun_out/) into the committed, human-readable summaries under
profiles/ (run here, no GPU):

    python tools/make_profiles.py r1 integrate=gpurun_out/prof_integrate_kernel_r1d.ncu-rep ... launches=gpurun_out/launches_r1.csv

Writes profiles/<round>_<name>.txt (selected raw metrics + top stall reasons), profiles/<round>_launches.txt (per
kernel totals and shares of the bench command's launch list) and profiles/traffic.json (DRAM bytes per launch of
the dominant kernels, read by bench.py for roofline.traffic).
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402


def raw_rows(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def main():
    rnd = sys.argv[1]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for arg in sys.argv[2:]:
        name, path = arg.split("=", 1)
        units = 1.0
        if path.count("=") == 1:                      # name=path=units_per_launch (e.g. pairs in one ICP launch)
            path, u_ = path.split("=")
            units = float(u_)
        if name == "launches":
            rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
            hdr = rows[0]
            iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
            agg = collections.defaultdict(lambda: [0, 0.0])
            for r in rows[1:]:
                k = r[iK].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
                v = float(r[iV].replace(",", ""))
                v = v / 1000 if r[iU] == "ns" else (v * 1000 if r[iU] == "ms" else v)
                agg[k][0] += 1
                agg[k][1] += v
            tot = sum(v[1] for v in agg.values())
            with open(os.path.join(ROOT, "profiles", "%s_launches.txt" % rnd), "w") as f:
                f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, launch list of bench.py (%s)\n" % os.path.basename(path))
                f.write("# cold-cache, serialised launches: compare SHARES, not absolutes\n")
                for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                    f.write("%-40s launches=%5d total_us=%12.1f avg_us=%10.1f share=%5.1f%%\n" % (k[:40], n, t, t / n, 100 * t / tot))
            continue
        hdr, unit_row, data = raw_rows(path)
        out = os.path.join(ROOT, "profiles", "%s_%s.txt" % (rnd, name))
        import io
        import contextlib
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ncu_summary.main(path)
        with open(out, "w") as f:
            f.write("# ncu --set full --clock-control none, source: %s\n" % os.path.basename(path))
            f.write(buf.getvalue())
        r = data[0]
        g = lambda m: float(r[hdr.index(m)].replace(",", ""))   # noqa: E731
        u = lambda m: unit_row[hdr.index(m)]                         # noqa: E731

        def to_bytes(m):
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u(m)]
            return g(m) * scale
        def opt(m):
            try:
                return g(m)
            except Exception:
                return None
        issue = {"issue_active_pct": opt("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                 "warp_instructions": opt("smsp__inst_executed.sum"),
                 "sm_throughput_pct": opt("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                 "dram_throughput_pct": opt("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                 "active_lanes_per_instruction": opt("smsp__thread_inst_executed_per_inst_executed.ratio"),
                 "note": "instruction issue is the binding roof of this kernel: issue_active_pct = share of cycles in which a "
                         "scheduler issued (100 = one warp instruction per scheduler per cycle)"}
        traffic[name] = {"issue": issue, "dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
                         "duration_us_under_ncu": g("gpu__time_duration.sum") * ({"us": 1, "ms": 1e3, "ns": 1e-3}[u("gpu__time_duration.sum")]),
                         "units_per_launch": units, "source": os.path.basename(path), "round": rnd}
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
