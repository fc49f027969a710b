"""Print the key figures of a bench.py JSON line (file argument)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d["roofline"]
print("N=%d value %.0f frames/s (%s) e2e %.0f frac %.3f" % (d["n_gpus"], d["value"], d["scaling"], d["e2e"]["value"], r["frac"]))
print("  touch %.2f classify %.2f integrate %.2f serial %.2f pipelined %.2f us/frame" % (
    r["touch_us_per_frame"], r["classify_us_per_frame"], r["integrate_us_per_frame"], r["serial_us_per_frame"], r["pipelined_us_per_frame"]))
print("  e2e feed:", d["e2e"].get("feed"), "rows", d["e2e"].get("result_rows"))
print("  union:", d.get("union_check"))
print("  replica:", d.get("replica"))
if d.get("ctrlgrid"):
    c = d["ctrlgrid"]
    print("  ctrlgrid %.0f frames/s, reproject %.2f us/frame, blocks %d, cpu %s" % (
        c["value"], c["reproject_us_per_frame"], c["allocated_blocks"], (c.get("cpu_baseline") or {}).get("value")))
if d.get("icp"):
    i = d["icp"]
    print("  icp %.0f pair-it/s passes %s e2e %.0f cpu %s" % (i["value"], i["passes_ms"], i["e2e"]["value"], (i.get("cpu_baseline") or {}).get("value")))
print("  cpu_baseline:", (d.get("cpu_baseline") or {}).get("value"), "clocks", d.get("clocks"))
if d.get("fragopt"):
    f = d["fragopt"]
    if "error" in f:
        print("  fragopt ERROR:", f["error"])
    else:
        print("  fragopt %.3g corr/s, steps %s ms, kernel %.2f ms (%.3g rows/s, %.2f TFLOP/s fp64, frac %.3f), rows/flush %.0f, e2e %.3g, cpu %s, pose err %.2e (from %.2e)" % (
            f["value"], f["steps_ms"], f["kernel"]["ms"], f["kernel"]["rows_per_s"], f["roofline"]["achieved"], f["roofline"]["frac"],
            f["kernel"]["rows_per_flush"], f["e2e"]["value"], (f.get("cpu_baseline") or {}).get("value"), f["max_pose_error_vs_truth"],
            f["initial_pose_error"]))
