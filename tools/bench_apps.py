"""Executable-level throughput of the drop-in binaries, files in -> files out (run on the GPU box):

    python tools/bench_apps.py [--frames 1000] [--fragments 50] [--points 200000] [--devices 0[,1,..]] [--out gpurun_out/r2_apps.json]

  Integrate            BASELINE configs[1]: depth container (614 KB/frame) + .log -> world.pcd, `--bench_json` timings
  BuildCorrespondence  BASELINE configs[2]: cloud_bin_<i>.pcd + pairs.log, `--reg_traj .. --registration --reg_num 0
                       --fixed_iters` -> reg_output.log + corres_<i>_<j>.txt, `--bench_json` timings (PCD load, upload, gate,
                       ICP, correspondence pass, corres file writes)
  FragmentOptimizer    --slac on the files BuildCorrespondence just wrote (reg_output.log, corres_<i>_<j>.txt) and the fragments' PCD files
Inputs are generated once into a scratch directory (not timed); the page cache is warm, as it is for the reference
binary timed the same way.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from elasticreconstruction_b200 import synth  # noqa: E402

BIN = os.path.join(ROOT, "elasticreconstruction_b200", "bin")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--fragments", type=int, default=50)
    ap.add_argument("--points", type=int, default=200000)
    ap.add_argument("--devices", default="0")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_apps.json"))
    ap.add_argument("--skip-icp", action="store_true")
    args = ap.parse_args()
    import torch
    res = {"devices": args.devices}
    with tempfile.TemporaryDirectory() as d:
        # ---------------- Integrate
        poses = bench.make_sequence()[: args.frames + 1] if args.frames < bench.SEQ_FRAMES else np.concatenate([bench.make_sequence(), bench.make_sequence()[:1]])
        frames = bench.room_depth_device(poses[:args.frames], torch.device("cuda", 0)).cpu().numpy().view(np.uint16)
        synth.write_depth_container(os.path.join(d, "depth.raw"), frames)
        synth.write_log(os.path.join(d, "traj.log"), poses)
        del frames
        torch.cuda.empty_cache()
        runs = []
        for rep in range(3):
            t0 = time.perf_counter()
            p = subprocess.run([os.path.join(BIN, "Integrate"), "-oni", os.path.join(d, "depth.raw"), "--ref_traj", os.path.join(d, "traj.log"),
                                "--save_to", os.path.join(d, "world.pcd"), "--devices", args.devices, "--bench_json", os.path.join(d, "int.json")],
                               cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            wall = time.perf_counter() - t0
            assert p.returncode == 0, p.stdout[-2000:]
            j = json.load(open(os.path.join(d, "int.json")))
            j["process_wall_ms"] = wall * 1e3
            j["frames_per_s_integrate_phase"] = j["frames_integrated"] / (j["integrate_ms"] * 1e-3)
            j["world_pcd_bytes"] = os.path.getsize(os.path.join(d, "world.pcd"))
            runs.append(j)
        res["Integrate"] = {"workload": "configs[1]: %d frames 640x480 from a depth container, rigid .log, world.pcd out" % args.frames,
                            "runs": runs, "best_frames_per_s_integrate_phase": max(r["frames_per_s_integrate_phase"] for r in runs)}
        print(json.dumps(res["Integrate"]), flush=True)
        os.remove(os.path.join(d, "depth.raw"))
        # ---------------- BuildCorrespondence
        if not args.skip_icp:
            n = args.fragments
            for i in range(n):
                x, nr = synth.make_local_fragment(i, args.points)
                synth.write_pcd_xyzrgbnormal(os.path.join(d, "cloud_bin_%d.pcd" % i), x, nr)
            pairs = [(i, j) for i in range(n) for j in range(i + 1, n)]
            Tinit = np.array([synth.pair_init(i, j) for (i, j) in pairs])
            synth.write_log(os.path.join(d, "pairs.log"), Tinit, ids=[(i, j, n) for (i, j) in pairs])
            runs = []
            for rep in range(2):
                cwd = os.path.join(d, "run%d" % rep)
                os.makedirs(cwd)
                t0 = time.perf_counter()
                p = subprocess.run([os.path.join(BIN, "BuildCorrespondence"), "--reg_traj", os.path.join(d, "pairs.log"), "--registration",
                                    "--reg_num", "0", "--fixed_iters", "--devices", args.devices, "--bench_json", os.path.join(cwd, "bc.json")],
                                   cwd=cwd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
                wall = time.perf_counter() - t0
                assert p.returncode == 0, p.stderr[-2000:]
                j = json.load(open(os.path.join(cwd, "bc.json")))
                j["process_wall_ms"] = wall * 1e3
                j["pair_iterations_per_s_icp_phase"] = j["icp_iterations"] / (j["icp_ms"] * 1e-3)
                j["pair_iterations_per_s_whole_run"] = j["icp_iterations"] / (j["total_ms"] * 1e-3)
                j["corres_bytes"] = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d) if f.startswith("corres_"))
                runs.append(j)
            res["BuildCorrespondence"] = {"workload": "configs[2]: %d fragments ~%dk points (binary PCD), all %d pairs, --registration --reg_num 0 "
                                                      "--fixed_iters, corres files written" % (n, args.points // 1000, len(pairs)), "runs": runs}
            print(json.dumps(res["BuildCorrespondence"]), flush=True)
            # ---------------- FragmentOptimizer --slac on BuildCorrespondence's own output files (added after the round's GPU
            # minutes were spent: this block has not run yet)
            try:
                cwd = os.path.join(d, "run1")
                traj = [np.linalg.inv(synth.BASEPOSE) @ synth.fragopt_initial_pose(i) @ synth.BASEPOSE for i in range(n)]
                synth.write_log(os.path.join(d, "fo_traj.log"), traj)
                fo = os.path.join(d, "fo")
                os.makedirs(fo)
                t0 = time.perf_counter()
                p = subprocess.run([os.path.join(BIN, "FragmentOptimizer"), "--slac", "--rgbdslam", os.path.join(d, "fo_traj.log"), "--registration",
                                    os.path.join(cwd, "reg_output.log"), "--dir", d + os.sep, "--num", str(n), "--interval", "1", "--blacklistpair", "1000",
                                    "--iteration", "5", "--devices", args.devices, "--bench_json", os.path.join(fo, "fo.json")],
                                   cwd=fo, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
                wall = time.perf_counter() - t0
                assert p.returncode == 0, p.stderr[-2000:]
                j = json.load(open(os.path.join(fo, "fo.json")))
                j["process_wall_ms"] = wall * 1e3
                j["correspondences_per_s_iterations"] = j["rows_assembled"] / max(1e-9, (j["assemble_ms"] + j["regulariser_ms"] + j["solve_ms"] + j["update_ms"]) * 1e-3)
                res["FragmentOptimizer"] = {"workload": "--slac, 5 iterations, on the %d fragments' PCD files and the corres files BuildCorrespondence wrote" % n,
                                            "run": j}
                print(json.dumps(res["FragmentOptimizer"]), flush=True)
            except Exception as e:      # noqa: BLE001
                res["FragmentOptimizer"] = {"error": "%s: %s" % (type(e).__name__, e)}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
