"""fused SIGNNet scoring head (models.py:370-376 + :339-346, eval mode) on tcgen05 tensor
cores against a plain PyTorch reference of the same op (float64).  TF32 inputs: tolerance 2e-3 of max|ref|."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference(joint, W, b, scale, shift):
    h = torch.nn.functional.elu(joint.double() @ W.double().t() + b.double()) * scale.double() + shift.double()
    return (h[0::2] * h[1::2]).float()


@pytest.mark.parametrize('rows,kd', [(256, 64), (128, 32), (1000, 2004), (2, 36), (130, 516), (4096, 2004), (33000, 1028),
                                     (600, 1542), (514, 7)])   # 1542 = (K+1) F' of BASELINE config 4: not a multiple of 4
def test_sign_head_matches_torch(rows, kd):
    from s3grl_b200 import sign_head
    g = torch.Generator(device='cuda').manual_seed(rows + kd)
    joint = torch.rand((rows, kd), device='cuda', generator=g) / 8        # row-normalised-feature-like magnitudes
    W = (torch.rand((256, kd), device='cuda', generator=g) - 0.5) * (2.0 / np.sqrt(kd)) * 4
    b = (torch.rand(256, device='cuda', generator=g) - 0.5)
    scale = torch.rand(256, device='cuda', generator=g) + 0.5
    shift = torch.rand(256, device='cuda', generator=g) - 0.5
    got = sign_head(joint, W, b, scale, shift)
    ref = _reference(joint, W, b, scale, shift)
    assert got.shape == (rows // 2, 256)
    err = float((got - ref).abs().max())
    assert err <= 2e-3 * float(ref.abs().max()) + 1e-6, f"max abs err {err:.3e} vs max|ref| {float(ref.abs().max()):.3e}"


def test_sign_head_on_loader_batches_and_module_parameters():
    """End of the path: precompute -> JointLoader -> fused head, against torch modules in eval mode."""
    from s3grl_b200 import DeviceGraph, JointLoader, PrecomputedList, datasets as ds, fold_batchnorm, precompute, sign_head
    edges, N, _ = ds.load_graph('usair')
    A, splits = ds.split_links(edges, N, seed=1)
    X = ds.synthetic_features(N, 16, 0.5, 3)                      # F' = 17, K + 1 = 4 -> 68 columns
    links = ds.all_links(splits)[:, :300]
    res = precompute(DeviceGraph(A, X), links, 2, 3, 'PoS')
    lst = PrecomputedList(res.xs, res.row_ptr, 1)
    torch.manual_seed(0)
    lin = torch.nn.Linear(68, 256).cuda()
    bn = torch.nn.BatchNorm1d(256).cuda()
    with torch.no_grad():
        bn.running_mean.uniform_(-0.2, 0.2)
        bn.running_var.uniform_(0.5, 1.5)
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.3, 0.3)
    bn.eval()
    scale, shift = fold_batchnorm(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, bn.eps)
    for batch in JointLoader(lst, 64, shuffle=True, seed=1):
        assert batch.joint.stride(0) % 4 == 0                         # padded by the loader: consumed in place
        got = sign_head(batch.joint, lin.weight.detach(), lin.bias.detach(), scale, shift)
        with torch.no_grad():
            h = bn(torch.nn.functional.elu(lin(torch.cat([batch.x] + [batch[f'x{k}'] for k in (1, 2, 3)], -1))))
            center = batch.ptr[:-1]
            ref = h[center] * h[center + 1]                        # models.py:341-346
        err = float((got - ref).abs().max())
        assert err <= 2e-3 * float(ref.abs().max()) + 1e-6


def test_sign_head_without_pooling_for_ccn_flows():
    """pool=False: h itself, for PoS Plus where models.py:347-367 pools a variable number of CCN rows per link."""
    from s3grl_b200 import sign_head
    g = torch.Generator(device='cuda').manual_seed(9)
    joint = torch.rand((777, 100), device='cuda', generator=g) / 4
    W = (torch.rand((256, 100), device='cuda', generator=g) - 0.5)
    b, scale, shift = (torch.rand(256, device='cuda', generator=g) - 0.5 for _ in range(3))
    got = sign_head(joint, W, b, scale + 1.0, shift, pool=False)
    ref = (torch.nn.functional.elu(joint.double() @ W.double().t() + b.double()) * (scale.double() + 1.0) + shift.double()).float()
    assert got.shape == (777, 256)
    assert float((got - ref).abs().max()) <= 2e-3 * float(ref.abs().max()) + 1e-6


def test_sign_head_rejects_bad_shapes():
    from s3grl_b200 import sign_head
    z = torch.zeros
    with pytest.raises(NotImplementedError):
        sign_head(z((4, 8), device='cuda'), z((128, 8), device='cuda'), z(128, device='cuda'), z(128, device='cuda'), z(128, device='cuda'))
    with pytest.raises(ValueError):
        sign_head(z((3, 8), device='cuda'), z((256, 8), device='cuda'), z(256, device='cuda'), z(256, device='cuda'), z(256, device='cuda'))
    with pytest.raises(ValueError):
        sign_head(z((4, 8), device='cuda'), z((256, 8), device='cuda'), z(255, device='cuda'), z(256, device='cuda'), z(256, device='cuda'))
