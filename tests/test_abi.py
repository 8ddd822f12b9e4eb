"""CPU-side checks of the C-ABI boundary: the library loads without a GPU and exports every symbol that
include/season_nerf_b200.h declares; the ctypes table covers the header.  No compute calls."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "season_nerf_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(snb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    from season_nerf_b200 import build
    lib = build.build()
    out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT\s+(snb_[a-z0-9_]+)", out))
    decl = _declared()
    assert len(decl) >= 15
    missing = [d for d in decl if d not in exported]
    assert not missing, missing


def test_ctypes_table_matches_header():
    from season_nerf_b200 import _lib
    lib = _lib.load()
    assert lib.snb_version() >= 100
    assert sorted(_lib.SIGNATURES) == _declared()
    assert _lib.launch_count() == 0
    assert b"invalid argument" in lib.snb_error_string(-1)


def test_no_cpu_fallback():
    import torch as t
    import season_nerf_b200 as snb
    from season_nerf_b200._lib import SeasonNerfCudaError
    net = snb.T_NeRF(64, 4)
    with pytest.raises(SeasonNerfCudaError):
        net.forward(t.zeros(4, 3), t.zeros(4, 3), t.zeros(4, 4))
    with pytest.raises(SeasonNerfCudaError):
        snb.get_PV(t.zeros(2, 4, 1), t.zeros(2, 4, 1))
    with pytest.raises(SeasonNerfCudaError):
        snb.sample_pt_coarse(t.zeros(2, 3), t.zeros(2, 3), 8, True, device=t.device("cpu"))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "season_nerf_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("#", "\n#").split("\n#")[0] or \
                not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_state_dict_is_the_reference_key_set(params0):
    import season_nerf_b200 as snb
    net = snb.T_NeRF(512, 4)
    assert list(net.state_dict().keys()) == list(params0.keys())
    net.load_state_dict(params0, strict=True)
    assert sum(p.numel() for p in net.parameters()) == 3195820
