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


# ---- CCN pooling on the GPU (north_star kernel 3 "center / CCN pooling"; reference models.py:347-362) ----
def _pool_reference(h, row_ptr, mode):
    """_centre_pool_helper with k_heuristic set, restated in float64 torch: [h_src * h_dst | mean or sum of the rest]."""
    rp = row_ptr.tolist()
    out = []
    for a, b in zip(rp[:-1], rp[1:]):
        rest = h[a + 2:b].double()
        pooled = rest.sum(0) if mode == 'sum' else (rest.mean(0) if b - a > 2 else torch.zeros_like(h[a].double()))
        out.append(torch.cat([h[a].double() * h[a + 1].double(), pooled]))
    return torch.stack(out).float()


@pytest.mark.parametrize('cols', [256, 501, 2004])
@pytest.mark.parametrize('mode', ['mean', 'sum'])
def test_segment_pool_center_layout(cols, mode):
    from s3grl_b200 import segment_pool
    g = torch.Generator().manual_seed(cols)
    counts = torch.randint(2, 12, (300,), generator=g)
    counts[::7] = 2                               # links without CCN rows: the pooled half is zero
    counts[5] = 216                               # PubMed's largest union
    row_ptr = torch.zeros(301, dtype=torch.int64)
    torch.cumsum(counts, 0, out=row_ptr[1:])
    h = torch.randn((int(row_ptr[-1]), cols), generator=g).cuda()
    got = segment_pool(h, row_ptr.cuda(), mode, 'center')
    ref = _pool_reference(h.cpu(), row_ptr, mode)
    assert got.shape == (300, 2 * cols)
    assert float((got.cpu() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    assert torch.equal(got[:, :cols], h[row_ptr[:-1].cuda()] * h[row_ptr[:-1].cuda() + 1])      # the center product is exact
    again = segment_pool(h, row_ptr.cuda(), mode, 'center')
    assert torch.equal(got, again)                                                                # fixed summation order
    with pytest.raises(NotImplementedError):
        segment_pool(h, row_ptr.cuda(), 'concat')


def test_pooled_output_mode_reduces_the_oracles_rows():
    """parity of the pooled output mode is "reduce the oracle's rows" — PoS Plus union on Cora, the CUDA
    rows pooled on the GPU against the oracle's rows pooled in NumPy."""
    from oracle import s3grl_oracle as orc
    from s3grl_b200 import DeviceGraph, datasets as ds, pool_rows, precompute
    edges, N, X = ds.load_graph('cora')
    X = ds.normalize_features(X)[:, :96].copy()
    A, splits = ds.split_links(edges, N, seed=1)
    links = ds.all_links(splits)[:, ::97][:, :120]
    res = precompute(DeviceGraph(A, X), links, 2, 3, 'PoS', 'union')
    ref = orc.pos_precompute(links, 2, A, X, 3, 'union')
    rp = ref['row_ptr']
    for mode in ('sum', 'mean'):
        pooled = pool_rows(res, mode)
        assert pooled.row_ptr.tolist() == list(range(0, 3 * links.shape[1] + 1, 3))
        for k in range(4):
            want = np.zeros((links.shape[1], 3, X.shape[1] + 1), np.float64)
            for i in range(links.shape[1]):
                rows = ref['xs'][k][rp[i]:rp[i + 1]].astype(np.float64)
                want[i, 0], want[i, 1] = rows[0], rows[1]
                if rows.shape[0] > 2:
                    want[i, 2] = rows[2:].sum(0) if mode == 'sum' else rows[2:].mean(0)
            got = pooled.xs[k].view(links.shape[1], 3, -1).cpu().numpy()
            assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max() + 1e-7, (mode, k)


def test_sign_head_ccn_matches_torch():
    """models.py:370-376 + :347-362 for PoS Plus: tcgen05 head (pool = 0) followed by the GPU segment pooling."""
    from s3grl_b200 import sign_head_ccn
    g = torch.Generator(device='cuda').manual_seed(3)
    counts = torch.randint(2, 9, (400,))
    row_ptr = torch.zeros(401, dtype=torch.int64)
    torch.cumsum(counts, 0, out=row_ptr[1:])
    rows = int(row_ptr[-1])
    joint = torch.rand((rows, 2004), device='cuda', generator=g) / 8
    W = (torch.rand((256, 2004), device='cuda', generator=g) - 0.5) * 0.2
    b, sc, sh = (torch.rand(256, device='cuda', generator=g) - 0.5 for _ in range(3))
    for mode in ('mean', 'sum'):
        got = sign_head_ccn(joint, row_ptr.cuda(), W, b, sc + 1.0, sh, mode)
        h = (torch.nn.functional.elu(joint.double() @ W.double().t() + b.double()) * (sc.double() + 1.0) + sh.double()).float().cpu()
        ref = _pool_reference(h, row_ptr, mode)
        assert got.shape == (400, 512)
        assert float((got.cpu() - ref).abs().max()) <= 2e-3 * float(ref.abs().max()) + 1e-6      # TF32 head tolerance
