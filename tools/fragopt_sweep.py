"""Randomised parity sweep of the FragmentOptimizer executable against the VERBATIM reference binary (CPU, build container):

    python tools/fragopt_sweep.py [seed] [trials]

Every trial draws a mode (slac / rigid / nonrigid), 2-5 synthetic fragments, lattice resolution, length, weight, iteration counts,
input format (.xyzn / .pcd) and correspondence density, runs apps/FragmentOptimizer.cpp linked against the CPU emulation of the
kernels (tests/emul/_build/FragmentOptimizer_emul, built by tests/test_fragopt_cpu.py) and oracle/_ref/FragmentOptimizer_ref on the
same files, and prints the largest difference of output.ctr / pose.log.  Round 2: 2 x 14 trials, worst 1e-10 (one unit of the last
printed digit)."""
import sys, os, subprocess, shutil, random
sys.path.insert(0,"/root/repo"); sys.path.insert(0,"/root/repo/tests")
import numpy as np
from elasticreconstruction_b200 import synth
import helpers
EXE="/root/repo/tests/emul/_build/FragmentOptimizer_emul"; REF="/root/repo/oracle/_ref/FragmentOptimizer_ref"
rng=random.Random(int(sys.argv[1]) if len(sys.argv)>1 else 1)
worst=0
for trial in range(int(sys.argv[2]) if len(sys.argv)>2 else 10):
    num=rng.choice([2,3,4,5]); pts=rng.choice([1500,3000,5000]); mode=rng.choice(["slac","rigid","nonrigid"])
    res=rng.choice([3,4,5,6]) if mode=="nonrigid" else rng.choice([4,6,8,9])
    length=rng.choice([3.0,3.0,3.3]); weight=rng.choice([0.3,1.0,5.0]); it=rng.choice([1,2,3]); inner=rng.choice([1,2,4])
    stride=rng.choice([1,1,3]); pcd=rng.choice([False,True])
    base=os.path.join(os.environ.get("TMPDIR","/tmp"),"fragopt_sweep","t%d"%trial)+"/"
    shutil.rmtree(base, ignore_errors=True)
    data=base+"data/"
    ds=synth.make_fragopt_dataset(data, num=num, n_points=pts, pcd=pcd, stride=stride)
    args=["--rgbdslam",data+"traj.log","--registration",data+"reg_output.log","--dir",data,"--num",str(num),"--interval","1","--blacklistpair","0",
          "--resolution",str(res),"--length",str(length),"--weight",str(weight),"--iteration",str(it),"--inner_iteration",str(inner),"--"+mode]
    outs={}
    for who,exe in (("mine",EXE),("ref",REF)):
        os.makedirs(base+who)
        p=subprocess.run([exe]+args,cwd=base+who,stdout=subprocess.PIPE,stderr=subprocess.STDOUT,text=True)
        outs[who]=p
    try:
        a=np.loadtxt(base+"mine/output.ctr"); b=np.loadtxt(base+"ref/output.ctr")
        dc=np.abs(a-b).max()
        dp=0.0
        if mode!="nonrigid":
            dp=np.abs(helpers.read_log_poses(base+"mine/pose.log")-helpers.read_log_poses(base+"ref/pose.log")).max()
        print(trial, mode, "num",num,"pts",pts,"res",res,"len",length,"w",weight,"it",it,inner,"stride",stride,"pcd",pcd,"-> ctr %.2e pose %.2e"%(dc,dp), flush=True)
        worst=max(worst,dc,dp)
    except Exception as e:
        print(trial, mode, num, pts, res, "FAILED", e, outs["mine"].returncode, outs["mine"].stdout[-300:], flush=True)
print("worst", worst)
