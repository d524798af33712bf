"""Row f3: BatchedTensorLoader yields the batches of the reference's DataLoader(TensorDataset(...)) (lib/dataloader.py:92-99),
in the same order for the same seed, while consuming the global generator identically (so everything drawn later -- masks,
dropout-free model init -- stays aligned too)."""
import pytest
import torch

from gptst_b200.data import BatchedTensorLoader


@pytest.mark.parametrize("shuffle,drop_last,n,bs", [(True, True, 103, 8), (True, False, 103, 8), (False, True, 64, 16), (True, True, 16, 16)])
def test_same_batches_and_rng_as_dataloader(shuffle, drop_last, n, bs):
    X = torch.arange(n * 6, dtype=torch.float32).view(n, 2, 3)
    Y = -X[:, :1]
    torch.manual_seed(7)
    ref = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(X, Y), batch_size=bs, shuffle=shuffle, drop_last=drop_last)
    want = [[(a.clone(), b.clone()) for a, b in ref] for _ in range(2)]          # two epochs
    tail_ref = torch.rand(3)
    torch.manual_seed(7)
    mine = BatchedTensorLoader(X, Y, batch_size=bs, shuffle=shuffle, drop_last=drop_last)
    assert len(mine) == len(ref)
    got = [[(a, b) for a, b in mine] for _ in range(2)]
    tail = torch.rand(3)
    for e in range(2):
        assert len(got[e]) == len(want[e])
        for (a, b), (c, d) in zip(got[e], want[e]):
            assert torch.equal(a, c) and torch.equal(b, d)
    assert torch.equal(tail, tail_ref)           # the global generator was consumed identically


def test_rejects_mismatched_tensors():
    with pytest.raises(ValueError):
        BatchedTensorLoader(torch.zeros(3, 2), torch.zeros(4, 2), batch_size=2)
