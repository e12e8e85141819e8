import torch


def test_apex_flatten_unflatten_roundtrip():
    from pytorch_distributed_b200.apex.parallel import flatten, unflatten
    ts = [torch.randn(3, 4), torch.randn(5), torch.randn(2, 2, 2)]
    flat = flatten(ts)
    assert flat.shape == (3 * 4 + 5 + 8,)
    back = unflatten(flat, ts)
    for a, b in zip(ts, back):
        assert torch.equal(a, b)
    back[1].zero_()
    assert float(flat[12:17].abs().sum()) == 0.0     # views, not copies
