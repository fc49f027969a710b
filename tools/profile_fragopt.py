"""Driver for timing / profiling the FragmentOptimizer leg at a chosen size (one process, no CPU leg):

    python tools/profile_fragopt.py [fragments] [points] [iterations] [nonrigid]

Prints fragopt_bench.run()'s result as one JSON line.  Under ncu: -k regex:slac_assemble -s 1 -c 1 captures the second
assembly launch (steady state)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fragopt_bench  # noqa: E402

if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:4]]
    fragments = a[0] if len(a) > 0 else 8
    points = a[1] if len(a) > 1 else 50000
    iterations = a[2] if len(a) > 2 else 2
    if len(sys.argv) > 4 and sys.argv[4] == "nonrigid":
        print(json.dumps(fragopt_bench.nonrigid_probe(0, fragments=fragments, points=points)))
    else:
        r = fragopt_bench.run(0, fragments=fragments, points=points, iterations=iterations, warmup=1, cpu=False)
        print(json.dumps(r))
