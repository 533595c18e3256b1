"""Input preparation (pn2_preprocess.py) against the restatement of provider.py:5-21,166-180."""
import numpy as np
import pytest
import torch

import cloudgen


def _raw(B, N, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(B, N, 3, generator=g) * 3.0 - 1.0 + torch.randn(B, 1, 3, generator=g)).float()     # off-centre, unnormalised


def test_oracle_restatement_properties(oracle):
    x = _raw(3, 200, 1)
    y = oracle.normalization(x.numpy())
    assert y.shape == (3, 200, 3) and np.allclose(y.mean(1), 0.0, atol=1e-6)
    assert np.allclose(np.sqrt((y ** 2).sum(-1)).max(1), 1.0, atol=1e-6)
    cat = torch.tensor([0, 6, 3])
    z = oracle.splice_torch(torch.from_numpy(y).float(), cat, 7)
    assert z.shape == (3, 200, 10) and torch.equal(z[..., 3:].sum(-1), torch.ones(3, 200))
    assert all(bool((z[b, :, 3 + int(cat[b])] == 1).all()) for b in range(3))


def test_splice_without_the_batch_loop_equals_the_reference_form(oracle):
    import pn2_preprocess
    x = _raw(4, 50, 2)
    cat = torch.tensor([1, 0, 6, 2])
    assert torch.equal(pn2_preprocess.splice_torch(x, cat, 7), oracle.splice_torch(x, cat, 7))


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,layout", [(5, 1024, "contig"), (3, 777, "strided"), (1, 1, "contig"), (64, 2048, "contig")])
def test_prepare_pose_input_on_the_gpu(oracle, B, N, layout):
    import pn2_preprocess
    x = _raw(B, N, 10 + N)
    if N == 1:
        x = x + 1.0       # a single point: centroid = the point, norm 0 -> 0/0 = nan in the reference too
    cat = torch.arange(B) % 7
    xin = x.permute(0, 2, 1).contiguous().permute(0, 2, 1) if layout == "strided" else x
    got = pn2_preprocess.prepare_pose_input(xin.cuda(), cat.cuda(), 7).cpu()
    exp = oracle.splice_torch(torch.from_numpy(oracle.normalization(x.numpy())).float(), cat, 7).permute(0, 2, 1)
    assert got.shape == (B, 10, N)
    if N == 1:
        assert torch.isnan(got[:, :3]).all() and torch.isnan(exp[:, :3]).all() and torch.equal(got[:, 3:], exp[:, 3:])
        return
    assert torch.equal(got[:, 3:], exp[:, 3:])                                   # one-hot rows: exact
    assert float((got[:, :3] - exp[:, :3]).abs().max()) <= 1e-5                  # normalised coordinates lie in [-1, 1]: the fp32 tolerance of the path
    plain = pn2_preprocess.normalization(xin.cuda()).cpu()
    assert plain.shape == (B, N, 3) and torch.equal(plain, got[:, :3].permute(0, 2, 1))
