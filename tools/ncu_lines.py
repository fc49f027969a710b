"""Per-source-line instruction / stall-sample shares from a .ncu-rep captured with --import-source on.
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [min_share_percent]
"""
import csv
import subprocess
import sys


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main(path, min_share=0.4):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(r for r in rows if "Instructions Executed" in r)
    ie, te, ss = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    lines = [r for r in rows if len(r) > te and r[0].isdigit()]
    tot = sum(num(r[ie]) for r in lines)
    tots = sum(num(r[ss]) for r in lines)
    print("total warp instructions %d, samples %d" % (tot, tots))
    for r in lines:
        if num(r[ie]) > tot * min_share / 100 or num(r[ss]) > tots * min_share / 100:
            print("%5s %6.2f%% inst %6.2f%% smp  %4.1f thr | %s" % (r[0], 100 * num(r[ie]) / tot, 100 * num(r[ss]) / max(tots, 1),
                                                                  num(r[te]) / max(1, num(r[ie])), r[1][:120]))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.4)
