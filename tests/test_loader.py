"""collated on-disk format (reference sgrl_link_pred.py:85, :204) and the
GPU-resident batch loader replacing DataLoader(..., follow_batch=[x1..xK]) (:1253-1269)."""
import numpy as np
import pytest
import torch

from s3grl_b200 import PrecomputedList, load_collated, save_collated


def _dataset(rng, L, F1, K, fixed):
    counts = np.full(L, 2) if fixed else rng.integers(2, 9, L)
    row_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    R = int(row_ptr[-1])
    xs = [torch.from_numpy(rng.random((R, F1), dtype=np.float32)) for _ in range(K + 1)]
    y = torch.from_numpy(rng.integers(0, 2, L).astype(np.int64))
    return PrecomputedList(xs, torch.from_numpy(row_ptr), y)


def _collate_like_pyg(ds, idx):
    """What Batch.from_data_list(follow_batch=...) + torch.cat(xs, -1) (models.py:372) give for links idx."""
    rp = ds.row_ptr.cpu().numpy()
    rows = np.concatenate([np.arange(rp[i], rp[i + 1]) for i in idx]) if len(idx) else np.zeros(0, np.int64)
    joint = np.concatenate([x.cpu().numpy()[rows] for x in ds.xs], axis=1)
    batch = np.concatenate([np.full(rp[i + 1] - rp[i], b) for b, i in enumerate(idx)]) if len(idx) else np.zeros(0, np.int64)
    return joint, batch, ds.y.cpu().numpy()[idx]


def test_collated_file_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    ds = _dataset(rng, 37, 11, 3, fixed=False)
    ds.extras['node_id'] = torch.arange(int(ds.row_ptr[-1]))
    path = tmp_path / 'SEAL_train_data.pt'
    save_collated(ds, path)
    data, slices = torch.load(path)
    # InMemoryDataset.collate layout: every key concatenated on dim 0 + its own slice vector; y is [L]
    assert sorted(data) == ['node_id', 'x', 'x1', 'x2', 'x3', 'y'] and data['y'].shape == (37,)
    assert torch.equal(slices['x'], ds.row_ptr) and torch.equal(slices['x3'], ds.row_ptr)
    assert torch.equal(slices['y'], torch.arange(38))
    back = load_collated(path)
    assert len(back) == 37 and torch.equal(back.row_ptr, ds.row_ptr) and torch.equal(back.y, ds.y)
    for a, b in zip(back.xs, ds.xs):
        assert torch.equal(a, b)
    assert torch.equal(back[5]['x2'], ds[5]['x2']) and torch.equal(back.extras['node_id'], ds.extras['node_id'])


def test_loader_refuses_cpu_datasets():
    from s3grl_b200 import JointLoader, joint_rows
    ds = _dataset(np.random.default_rng(1), 8, 5, 2, fixed=True)
    with pytest.raises(RuntimeError):
        JointLoader(ds, 4)
    with pytest.raises(RuntimeError):
        joint_rows(ds.xs, ds.row_ptr, torch.arange(4))


@pytest.mark.gpu
@pytest.mark.parametrize('fixed,L,F1,K', [(True, 1000, 501, 3), (False, 777, 130, 3), (True, 65, 1, 1), (False, 300, 33, 5),
                                          (False, 50, 9, 8)])
def test_joint_rows_matches_collate(fixed, L, F1, K):
    from s3grl_b200 import joint_rows
    rng = np.random.default_rng(L)
    ds = _dataset(rng, L, F1, K, fixed).to('cuda')
    idx = rng.permutation(L)[:L // 2 + 1]
    joint, batch, ptr = joint_rows(ds.xs, ds.row_ptr, torch.from_numpy(idx).cuda(), 2 if fixed else None)
    ref_joint, ref_batch, _ = _collate_like_pyg(ds, idx)
    assert np.array_equal(joint.cpu().numpy(), ref_joint)          # pure data movement: bit-exact
    assert joint.stride(0) % 4 == 0 and joint.shape[1] == (K + 1) * F1
    assert np.array_equal(batch.cpu().numpy(), ref_batch)
    uq, first = np.unique(ref_batch, return_index=True)            # models.py:341 center_indices
    assert np.array_equal(ptr.cpu().numpy()[:-1], first) and int(ptr[-1]) == ref_joint.shape[0]


@pytest.mark.gpu
@pytest.mark.parametrize('fixed', [True, False])
def test_joint_loader_epoch(fixed):
    from s3grl_b200 import JointLoader
    rng = np.random.default_rng(5)
    ds = _dataset(rng, 203, 21, 3, fixed).to('cuda')
    loader = JointLoader(ds, 32, shuffle=True, seed=3)
    assert len(loader) == 7
    seen = []
    perm = torch.randperm(203, device='cuda', generator=torch.Generator(device='cuda').manual_seed(3)).cpu().numpy()
    for bi, b in enumerate(loader):
        idx = perm[bi * 32:(bi + 1) * 32]
        ref_joint, ref_batch, ref_y = _collate_like_pyg(ds, idx)
        assert b.num_graphs == len(idx)
        assert np.array_equal(b.joint.cpu().numpy(), ref_joint)
        assert np.array_equal(b.batch.cpu().numpy(), ref_batch) and np.array_equal(b['x2_batch'].cpu().numpy(), ref_batch)
        assert np.array_equal(b.y.cpu().numpy(), ref_y)
        assert np.array_equal(b['x3'].cpu().numpy(), ref_joint[:, 63:84])
        assert np.array_equal(torch.cat([b.x, b['x1'], b['x2'], b['x3']], -1).cpu().numpy(), ref_joint)
        seen.append(idx)
    assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(203))
    # a second epoch draws a new permutation; an unshuffled loader walks the dataset in order
    first2 = next(iter(loader)).y.cpu().numpy()
    plain = next(iter(JointLoader(ds, 32)))
    assert np.array_equal(plain.joint.cpu().numpy(), _collate_like_pyg(ds, np.arange(32))[0])
    assert first2.shape == (32,)


@pytest.mark.gpu
def test_batches_stay_valid_across_epochs_and_iterators():
    """PyG's DataLoader hands out independent batches; by default so does JointLoader (a fresh epoch matrix per
    __iter__): a batch kept from epoch 1 — e.g. saved for backward — is not overwritten by epoch 2 or a second iterator."""
    from s3grl_b200 import JointLoader
    rng = np.random.default_rng(6)
    ds = _dataset(rng, 96, 9, 2, True).to('cuda')
    loader = JointLoader(ds, 32, shuffle=True, seed=1)
    kept = next(iter(loader))
    snapshot = kept.joint.clone()
    for a, b in zip(loader, loader):             # two concurrent iterators, then a further epoch
        assert a.joint.data_ptr() != b.joint.data_ptr()
    list(loader)
    assert torch.equal(kept.joint, snapshot)
    reuse = JointLoader(ds, 32, shuffle=True, seed=1, reuse_epoch_buffer=True)     # opt-in: one matrix for every epoch
    p0 = next(iter(reuse)).joint.data_ptr()
    assert next(iter(reuse)).joint.data_ptr() == p0
