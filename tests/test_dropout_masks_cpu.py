"""CPU checks of the host restatement of the kernels' dropout-mask generator (tests/dropout_masks.py).  The GPU tests
prove the CUDA kernels draw exactly these masks (element-wise parity of a dropout run against the oracle fed with them);
here the generator itself is checked as a random source: Philox known-answer vectors, keep rate and unbiasedness of every
site of a block, independence between sites / offsets / seeds, and the layouts the oracle expects."""
import numpy as np
import pytest
import torch

import dropout_masks as DM
from oracle import qavit_oracle as O


def test_philox_known_answers_10_rounds():
    """Random123 known-answer vectors for philox4x32-10 (kat_vectors: zero counter / key, all-ones, and the pi digits
    vector) pin the round function and the key schedule; the kernels run the same rounds, 7 of them."""
    saved = DM.PHILOX_ROUNDS
    DM.PHILOX_ROUNDS = 10
    try:
        r = DM.philox4x32(0, 0, 0, 0, 0, 0)
        assert [int(v) for v in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
        r = DM.philox4x32(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
        assert [int(v) for v in r] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
        r = DM.philox4x32(0xa4093822, 0x299f31d0, 0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344)
        assert [int(v) for v in r] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    finally:
        DM.PHILOX_ROUNDS = saved
    assert DM.PHILOX_ROUNDS == 7


@pytest.mark.parametrize("bf16", [False, True])
@pytest.mark.parametrize("nt", [16, 64])
def test_block_masks_rate_unbiasedness_and_layouts(bf16, nt):
    p, pp, B = 0.1, 0.3, 24
    m = DM.block_masks(0x5EED, 3, p, pp, B, nt, bf16=bf16)
    cfg = O.OracleConfig(family="qavit_v2")
    want = O.random_masks(cfg, B, nt, p, pp, torch.Generator().manual_seed(0))
    assert set(m) == set(want)
    thr = int(p * 65536 + 0.5)
    q = thr / 65536.0
    for k, v in m.items():
        assert tuple(v.shape) == tuple(want[k].shape), k          # the layouts quad_block(masks=) consumes
        if k.startswith("path"):
            nz = v[v != 0]
            assert torch.allclose(nz, torch.full_like(nz, 1 / (1 - int(pp * 65536 + 0.5) / 65536.0)), rtol=1e-6)
            continue
        n = v.numel()
        zeros = float((v == 0).float().mean())
        sigma = (q * (1 - q) / n) ** 0.5
        assert abs(zeros - q) < 5 * sigma + 1e-9, (k, zeros, q, sigma)
        assert abs(float(v.mean()) - 1.0) < 5 * sigma / (1 - q) + 1e-6, k      # E[keep scale] = 1: unbiased
        nz = v[v != 0]
        assert torch.allclose(nz, torch.full_like(nz, 1 / (1 - q)), rtol=1e-6), k


def test_sites_offsets_and_seeds_are_independent():
    p, B, nt = 0.25, 16, 64
    a = DM.block_masks(11, 0, p, 0.0, B, nt, bf16=False)
    b = DM.block_masks(11, 1, p, 0.0, B, nt, bf16=False)       # next forward call (offset advanced on the device)
    c = DM.block_masks(12, 0, p, 0.0, B, nt, bf16=False)       # another seed (another data-parallel rank)
    assert all(torch.equal(a[k], DM.block_masks(11, 0, p, 0.0, B, nt, bf16=False)[k]) for k in a)   # pure function

    def agree(x, y):
        return float(((x == 0) == (y == 0)).float().mean())
    indep = (1 - p) ** 2 + p ** 2                              # P[two independent masks agree]
    for k in a:
        n = a[k].numel()
        tol = 5 * (indep * (1 - indep) / n) ** 0.5
        assert abs(agree(a[k], b[k]) - indep) < tol, ("offset", k)
        assert abs(agree(a[k], c[k]) - indep) < tol, ("seed", k)
    # different sites of the same call with the same shape
    for x, y in [("proj_swa", "proj_msda"), ("proj_cga", "b2"), ("b2", "ffn"), ("att_swa", "att_msda")]:
        xa, ya = a[x].flatten(), a[y].flatten()
        n = min(xa.numel(), ya.numel())
        tol = 5 * (indep * (1 - indep) / n) ** 0.5
        assert abs(agree(xa[:n], ya[:n]) - indep) < tol, (x, y)
    # neighbouring elements are uncorrelated (lag-1 along the channel axis)
    z = (a["b2"] == 0).float()
    lag = float((z[..., 1:] * z[..., :-1]).mean())
    assert abs(lag - p * p) < 5 * (p * p * (1 - p * p) / z[..., 1:].numel()) ** 0.5
