"""CPU validation of the fused-kernel program builder: schedule hazards, slot allocation, weight packing and BatchNorm
folding, by interpreting the tables in numpy/torch and comparing with the oracle network."""
import numpy as np
import torch as t

from oracle import season_oracle as so


def _inputs(n, seed=1):
    g = t.Generator().manual_seed(seed)
    X = t.rand(n, 3, generator=g) * 2 - 1
    sun = t.nn.functional.normalize(t.rand(n, 3, generator=g), dim=1)
    Time = t.rand(n, 4, generator=g)
    enc = t.cat([so.pe_encode(X, 10), t.zeros(n, 1)], 1)
    senc = t.cat([so.pe_encode(sun, 4), t.zeros(n, 64 - 27)], 1)
    return X, sun, Time, enc, senc


import pytest
from season_nerf_b200 import packing2


@pytest.mark.parametrize("pk", [packing2])
def test_program_matches_oracle_network(params0, pk):
    packing = pk
    blob, info = packing.build_program(params0)
    assert info["n_mma"] == 193 and blob.nbytes % 128 == 0
    hdr = blob[:packing.HEADER_DT.itemsize].view(packing.HEADER_DT)[0]
    assert int(hdr["magic"]) == packing.MAGIC and int(hdr["w_off"]) % 1024 == 0
    X, sun, Time, enc, senc = _inputs(192)
    outs = packing.interpret(info, enc, senc)
    with t.no_grad():
        rho, col, vis, sky, cls, adj = so._link(params0, X, sun, Time, False)
    assert float((outs[packing.OUT_POS] - t.cat([rho, col], 1)).abs().max()) < 3e-2      # bf16 storage emulated
    assert float((outs[packing.OUT_VIS] - vis).abs().max()) < 3e-2
    assert float((outs[packing.OUT_ADJ] - adj.reshape(-1, 12)).abs().max()) < 3e-2


@pytest.mark.parametrize("pk", [packing2])
def test_sigma_only_program(params0, pk):
    packing = pk
    _, info = packing.build_program(params0, sigma_only=True)
    X, sun, Time, enc, senc = _inputs(64, seed=2)
    outs = packing.interpret(info, enc, senc)
    with t.no_grad():
        ref = so.linear(so.encode_x(params0, X, False), params0, "G_NeRF_net.fc10Sigma")
    assert float((outs[packing.OUT_POS] - ref).abs().max()) < 3e-2
    assert all(int(e["kind"]) != packing.K_ENC_SUN for e in info["epi"])


@pytest.mark.parametrize("pk", [packing2])
def test_schedule_checker_catches_hazards(params0, pk):
    packing = pk
    _, info = packing.build_program(params0)
    mma, epi = info["mma"].copy(), info["epi"].copy()
    assert packing.check_schedule(mma, epi)
    bad = mma.copy()
    i = int(np.nonzero(bad["flags"] & packing.F_WAIT_CHUNK)[0][5])
    bad["flags"][i] &= np.uint8(0xFF ^ packing.F_WAIT_CHUNK)                      # drop a chunk-ready wait
    try:
        packing.check_schedule(bad, epi)
        raise SystemExit("hazard not detected")
    except AssertionError:
        pass
    bad = mma.copy()
    i = int(np.nonzero((bad["flags"] & packing.F_ACC) == 0)[0][7])
    bad["flags"][i] &= np.uint8(0xFF ^ packing.F_WAIT_EMPTY)                      # restart a TMEM region without waiting for its drain
    try:
        packing.check_schedule(bad, epi)
        raise SystemExit("hazard not detected")
    except AssertionError:
        pass
