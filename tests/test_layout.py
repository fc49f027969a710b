"""Structural checks that need no GPU: the C-ABI library loads and exports every symbol include/elastic_b200.h
declares, the product never touches oracle/, and there is no CPU fallback."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "elastic_b200.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(er_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from elasticreconstruction_b200 import build, capi
    build.build_library()
    lib = ctypes.CDLL(capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), "libelastic_b200.so lacks %s" % n
    # the ctypes table and the header agree
    assert sorted(capi.SIGNATURES) == names
    assert capi.lib().er_build_arch() == b"sm_100a"


def test_sass_is_sm100a_only():
    from elasticreconstruction_b200 import capi
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", capi.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_product_never_references_the_oracle():
    bad = []
    for base in ("elasticreconstruction_b200", "apps", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dp.split(os.sep) or "__pycache__" in dp:
                continue
            for f in files:
                if not f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                    continue
                txt = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"pyoracle|liboracle|oracle_tsdf|oracle_icp|tsdf_oracle|icp_oracle", txt):
                    if f in ("er_icp.cu",) and "oracle/icp_oracle.c" in txt and not re.search(r"pyoracle|liboracle", txt):
                        continue      # a comment citing the oracle as the tested definition
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_no_cpu_fallback_without_device():
    """On a box without a GPU (this container) creating a handle must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from elasticreconstruction_b200 import capi
    with pytest.raises(capi.ElasticError, match="no CPU path|not available"):
        capi.TSDFVolume()
    import numpy as np
    with pytest.raises(capi.ElasticError, match="no CPU path|not available"):
        capi.Cloud(np.zeros((4, 3), np.float32), np.zeros((4, 3), np.float32))


def test_committed_traffic_table_is_usable():
    """bench.py reports roofline.traffic from profiles/traffic.json (DRAM bytes of one ncu --set full capture);
    a malformed table would silently turn that into null."""
    import importlib.util
    import json
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for k in ("integrate_kernel", "icp_iter_kernel"):
        assert isinstance(t[k]["dram_bytes_per_launch"], float) and isinstance(t[k]["units_per_launch"], float)
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.load_traffic("integrate_kernel") > 1e6
