"""CPU simulation (no GPU) of how well integrate_kernel's CTAs keep their warps busy inside one step.

For one 32-frame batch of the benchmark trajectory every touched block is classified on the CPU with the product's
own classify_ball (tests/emul/voxel_emul.cpp -> emul_classify_block); a sub-brick's cost for the batch is modelled as
300 x (exact-path frames) + 25 x (free-space frames) + 60 instructions, the ratio the ncu source view shows.  Compared:

  current   a CTA of 4 warps takes half a brick per step, warp w <-> sub-brick w (step time = slowest warp)
  dynamic   a CTA takes a whole brick; its 8 sub-bricks are tasks pulled by the 4 warps in natural order
  LPT       the same, tasks in order of decreasing cost (-DER_INT_BRICK_LPT)
  8 warps   a whole brick on 8 warps, one sub-brick each
and, as list scheduling of the whole launch on 1184 resident CTAs (resp. 4736 warps) treated as independent equal-speed
machines -- which they are not quite: CTAs of one SM share its issue slots -- the kernel's makespan for
  today     half-brick units handed to CTAs, heavy bricks first
  brick-LPT whole-brick units handed to CTAs (-DER_INT_BRICK_LPT)
  tasks     sub-brick tasks handed to warps, 1 / 2 / 4 per fetch (-DER_INT_TASKS), heavy sub-bricks first

    python tools/simulate_integrate_balance.py
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from elasticreconstruction_b200 import capi, synth  # noqa: E402
import pyoracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "emul", "_build", "libvoxel_emul.so")


def load():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-I",
                           os.path.join(ROOT, "elasticreconstruction_b200", "csrc"), os.path.join(ROOT, "tests", "emul", "voxel_emul.cpp"),
                           "-o", OUT])
    L = C.CDLL(OUT)
    fp, dp = C.POINTER(C.c_float), C.POINTER(C.c_double)
    L.emul_classify_block.argtypes = [fp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_double, C.c_double, dp, dp,
                                      C.c_int, C.c_int, C.POINTER(C.c_ubyte)]
    return L, fp, dp


def batch_classes(L, fp, dp, poses, frames, tight):
    full, one = {}, {}
    ov = pyoracle.OracleTSDF()
    for f in range(len(frames)):
        T = np.ascontiguousarray(poses[f], dtype=np.float64)
        Tinv = np.ascontiguousarray(capi.mat4_inverse(T), dtype=np.float64)
        scaled = np.ascontiguousarray(ov.scale_depth(frames[f]), dtype=np.float32)
        ov.integrate(frames[f], T)
        for key in ov.last_touched():
            cls = np.zeros(4096, dtype=np.uint8)
            L.emul_classify_block(scaled.ctypes.data_as(fp), 640, 480, 525.0, 525.0, 319.5, 239.5, 3.0 / 512, 0.03, T.ctypes.data_as(dp),
                                  Tinv.ctypes.data_as(dp), int(key), tight, cls.ctypes.data_as(C.POINTER(C.c_ubyte)))
            k = int(key)
            if k not in full:
                full[k], one[k] = np.zeros(4096, dtype=np.int32), np.zeros(4096, dtype=np.int32)
            full[k] += cls == 2
            one[k] += cls == 1
    return full, one


def simulate(full, one, a=300.0, b=25.0, c=60.0):
    tot = cur = dyn = lpt = w8 = 0.0
    nbr = 0
    for k in full:
        F, O = full[k].reshape(512, 8), one[k].reshape(512, 8)
        act = (F + O) > 0
        cost = np.where(act, a * F + b * O + c, 0.0)
        for br in np.where(act.any(axis=1))[0]:
            cs = cost[br]
            tot += cs.sum()
            nbr += 1
            cur += cs[:4].max() + cs[4:].max()
            for order, acc in ((list(cs), "dyn"), (sorted(cs, reverse=True), "lpt")):
                m = np.zeros(4)
                for t in order:
                    m[np.argmin(m)] += t
                if acc == "dyn":
                    dyn += m.max()
                else:
                    lpt += m.max()
            w8 += cs.max()
    return tot, cur, dyn, lpt, w8, nbr


def makespan(order, machines, chunk=1):
    import heapq
    h = [0.0] * machines
    heapq.heapify(h)
    for i in range(0, len(order), chunk):
        heapq.heappush(h, heapq.heappop(h) + sum(order[i:i + chunk]))
    return max(h)


def global_model(full, one, a=300.0, b=25.0, c=60.0, ctas=1184, warps=4):
    rng = np.random.default_rng(0)
    cur = [[], []]
    lpt = [[], []]
    task = [[], []]
    for k in full:
        F, O = full[k].reshape(512, 8), one[k].reshape(512, 8)
        act = (F + O) > 0
        cost = np.where(act, a * F + b * O + c, 0.0)
        for br in np.where(act.any(axis=1))[0]:
            cs = cost[br]
            light = int(F[br].sum() < 32)                       # classify_kernel: heavy bricks to the front
            cur[light] += [cs[:4].max(), cs[4:].max()]
            m = np.zeros(warps)
            for t in sorted(cs, reverse=True):
                m[np.argmin(m)] += t
            lpt[light].append(m.max())
            for q in range(8):
                if act[br, q]:
                    task[int(F[br, q] < 16)].append(cs[q])
    order = lambda two: list(rng.permutation(two[0])) + list(rng.permutation(two[1]))     # noqa: E731  arrival order inside a class
    total = sum(task[0]) + sum(task[1])
    res = {"ideal": total / (ctas * warps), "today": makespan(order(cur), ctas), "brick-LPT": makespan(order(lpt), ctas)}
    for ch in (1, 2, 4):
        res["tasks/%d" % ch] = makespan(order(task), ctas * warps, ch)
    return res


def main():
    pyoracle.build()
    L, fp, dp = load()
    poses = synth.quantize_log(synth.loop_trajectory(1000))[300:333]
    frames = synth.room_depth(poses[:32])
    base = None
    for tight in (0, 1):
        full, one = batch_classes(L, fp, dp, poses, frames, tight)
        tot, cur, dyn, lpt, w8, nbr = simulate(full, one)
        if base is None:
            base = cur
        print("%s classification: %d bricks, modelled work %.3g instr | warp utilisation inside a step: current %.1f %%, "
              "dynamic %.1f %%, LPT %.1f %%, 8 warps %.1f %% | summed step time relative to today's default: current %.3f, LPT %.3f"
              % ("box-based" if tight else "ball-based", nbr, tot, 100 * tot / (4 * cur), 100 * tot / (4 * dyn), 100 * tot / (4 * lpt),
                 100 * tot / (8 * w8), cur / base, lpt / base))
        g = global_model(full, one)
        if tight == 0:
            today = g["today"]
        print("    launch makespan model (relative to today's default): " +
              ", ".join("%s %.2f" % (k, v / today) for k, v in g.items()))


if __name__ == "__main__":
    main()
