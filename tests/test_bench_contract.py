"""bench.py's control flow and JSON contract on the CPU: the library and torch.cuda are replaced by stand-ins, the
reference arm / CPU baseline run for real.  Guards the one-line stdout rule and the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVER = r'''
import json, os, sys
import numpy as np
sys.path.insert(0, %(root)r)
import torch
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a: None
torch.cuda.empty_cache = lambda: None
torch.Tensor.cuda = lambda self, *a, **k: self
torch.Tensor.pin_memory = lambda self, *a, **k: self
_tensor = torch.tensor
torch.tensor = lambda *a, **k: _tensor(*a, **{kk: vv for kk, vv in k.items() if kk != "device"})
from elasticreconstruction_b200 import capi
import icp_bench


class FakeVol:
    def __init__(self, *a, **k):
        self.l = 0
        self.s = dict(updated_voxels=0, touched_blocks=0, one_pairs=0, full_pairs=0)
    def set_fusion(self, f): pass
    def set_partition(self, r, w): pass
    def set_control_grids(self, g, res, length): pass
    def integrate_device(self, p, poses, n):
        self.l += 4 * ((n + 31) // 32)
        for k in self.s:
            self.s[k] += 1000 * n
    def integrate_ctr_device(self, p, Ti, T0, A0, Ai, chunks): self.l += 6 * ((len(chunks) + 31) // 32)
    def profile_reproject_device(self, p, Ti, T0, A0, Ai, chunks): return 0.002 * len(chunks)
    def integrate_ptr(self, p, poses, n): pass
    def inputs_consumed(self): pass
    def extract_count(self): return 1000
    def extract_into(self, buf, cap): return 1000
    def extract_begin(self, buf, cap): pass
    def extract_wait(self): return 1000
    def extract(self): return np.zeros((10, 4), np.float32)
    def profile_device(self, p, poses, n):
        return (0.0018 * n, 0.0019 * n, 0.005 * n)
    def sync(self): pass
    def launch_count(self): return self.l
    def event_record(self, i): pass
    def event_elapsed_ms(self, a, b): return 33.0
    def set_stats(self, b): pass
    def stats(self): return dict(self.s)
    def num_blocks(self): return 100


class FakeIcp:
    n = 0
    class Cloud:
        def __init__(self, x, n, device=0, cell_size=0.03): pass
    def icp_align_batch(self, clouds, tgt, src, T, d, it, eps, fixed):
        m = len(tgt)
        return np.tile(np.eye(4), (m, 1, 1)), np.full(m, it), np.zeros(m)
    def icp_launch_count(self, dev):
        FakeIcp.n += 20
        return FakeIcp.n
    def icp_event_record(self, dev, i): pass
    def icp_event_elapsed_ms(self, dev, a, b): return 12.5


class FakeHost:
    def __init__(self, n):
        self.nbytes = n
        self._a = np.zeros(n, dtype=np.uint8)
        self.ptr = self._a.ctypes.data
    def array(self, dtype, count=None): return self._a.view(dtype)
    def close(self): pass


capi.TSDFVolume = FakeVol
capi.HostBuffer = FakeHost
icp_bench.capi = FakeIcp()
print("a library chatting on stdout before the benchmark line")        # must not reach the real stdout
import bench
_rd = bench.room_depth_device
bench.room_depth_device = lambda poses, device, **k: _rd(poses, "cpu", **k)
sys.argv = ["bench.py"] + %(argv)r
sys.exit(bench.main())
'''


def test_main_arm_prints_one_json_line_with_the_contract_keys(tmp_path):
    argv = ["--steps", "1", "--warmup", "1", "--frames-per-step", "50", "--cpu-frames", "6", "--icp-fragments", "3", "--icp-points", "2000",
            "--ctr-frames", "100", "--ctr-steps", "1", "--ctr-cpu-frames", "50"]
    script = tmp_path / "drive.py"
    script.write_text(DRIVER % dict(root=ROOT, argv=argv))
    p = subprocess.run([sys.executable, str(script)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    # the stand-in's own print happens before bench.main() parks stdout, so exactly that line plus the JSON line
    assert len(lines) == 2 and lines[0].startswith("a library chatting")
    d = json.loads(lines[1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "icp", "ctrlgrid"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["config"]["frames_per_step"] == 50 and "workload" in d["config"]
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"])
    assert d["roofline"]["traffic"] is not None
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"])
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 8
    assert d["icp"]["warmup_passes"] >= 3 and "roofline" in d["icp"] and d["icp"]["timed_passes"] >= 3
    assert d["e2e"]["d2h_bytes_per_step"] > 1000 and d["ctrlgrid"]["reproject_us_per_frame"] > 0
    assert d["ctrlgrid"]["cpu_baseline"]["kind"] == "reference" and d["ctrlgrid"]["cpu_baseline"]["value"] > 1


def test_reference_arm_line(tmp_path):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "depth frames/sec into 512^3 TSDF" and d["unit"] == "frames/s"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] == "reference"
    # ranks other than 0 of a multi-GPU launch stay silent and succeed
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    q = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120, cwd=ROOT)
    assert q.returncode == 0 and q.stdout.strip() == ""
