"""Compile-time variants of the library for A/B runs in ONE gpurun call (GPU minutes are the scarce resource).

    python tools/variants.py build tight -DER_CLASSIFY_TIGHT
    python tools/variants.py build pixv2 -DER_VOXEL_PIXEL_V2
    python tools/variants.py build both  -DER_CLASSIFY_TIGHT -DER_VOXEL_PIXEL_V2          (here: nvcc cross-compiles)
    python tools/variants.py run [--tests]                                                (on the GPU box)

`build` writes elasticreconstruction_b200/variants/libelastic_b200_<name>.so (git-ignored, but it travels with gpurun).
`run` times the TSDF and ICP paths with the default library and with every variant (capi loads $ER_B200_LIB), and with
--tests first runs the GPU parity tests of the two kernels files against each variant, so that a variant's number is
only read when its results are right.
"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elasticreconstruction_b200 import build  # noqa: E402

VDIR = os.path.join(build.HERE, "variants")


def build_variant(name, macros):
    os.makedirs(VDIR, exist_ok=True)
    objs = []
    for src in build._sources():
        obj = os.path.join(VDIR, "%s_%s.o" % (name, os.path.basename(src)))
        subprocess.check_call([build._nvcc()] + build.ARCH + build.NVCC_FLAGS + list(macros) + ["-c", src, "-o", obj])
        objs.append(obj)
    lib = os.path.join(VDIR, "libelastic_b200_%s.so" % name)
    subprocess.check_call([build._nvcc()] + build.ARCH + ["-shared", "-Xcompiler", "-fopenmp", "-o", lib] + objs + ["-lcudart"])
    for o in objs:
        os.remove(o)
    print(lib)


def run(with_tests):
    libs = [("default", None)] + [(os.path.basename(p)[len("libelastic_b200_"):-3], p) for p in sorted(glob.glob(os.path.join(VDIR, "*.so")))]
    for name, lib in libs:
        env = dict(os.environ)
        if lib:
            env["ER_B200_LIB"] = lib
        print("=== %s" % name, flush=True)
        # a variant only changes one of the two .cu files: test and time that path alone (GPU minutes)
        paths = ("tsdf", "icp") if not lib else (("icp",) if "colmask" in name or name.startswith("icp") else ("tsdf",))
        if with_tests and lib:
            rc = subprocess.call([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "--timeout", "300"] +
                                 [os.path.join(ROOT, "tests", "test_%s_gpu.py" % p) for p in paths], env=env, cwd=ROOT)
            print("parity tests: %s" % ("PASS" if rc == 0 else "FAIL (rc %d) -- ignore this variant's timings" % rc), flush=True)
        for tool, args, path in (("time_tsdf.py", ["256"], "tsdf"), ("time_icp.py", ["12", "256", "20"], "icp")):
            if path in paths:
                subprocess.call(["timeout", "200", sys.executable, os.path.join(ROOT, "tools", tool)] + args, env=env, cwd=ROOT)


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "build":
        build_variant(sys.argv[2], sys.argv[3:])
    elif len(sys.argv) >= 2 and sys.argv[1] == "run":
        run("--tests" in sys.argv)
    else:
        print(__doc__)
