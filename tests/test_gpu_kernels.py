"""Single-GPU numerics of every hand-written kernel against a plain PyTorch fp32 reference of the same op."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def C():
    from pytorch_distributed_b200 import _ext
    return _ext.lib()


@pytest.mark.parametrize("gdt,cdt", [(torch.bfloat16, torch.bfloat16), (torch.float32, None), (torch.float16, torch.float16)])
@pytest.mark.parametrize("nesterov", [False, True])
def test_fused_sgd_flat_matches_torch(gdt, cdt, nesterov):
    torch.manual_seed(0)
    n = 8 * 4099
    dev = "cuda"
    p0 = torch.randn(n, device=dev)
    grads = [torch.randn(n, device=dev).to(gdt) for _ in range(3)]
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.SGD([ref], lr=0.1, momentum=0.9, weight_decay=1e-4, nesterov=nesterov)
    master, mom = p0.clone(), torch.zeros(n, device=dev)
    copy = torch.zeros(n, device=dev, dtype=cdt) if cdt is not None else None
    hyper = torch.tensor([0.1, 0.9, 1e-4, 0.0, 1.0, 0, 0, 0], device=dev)
    for i, g in enumerate(grads):
        ref.grad = g.float()
        opt.step()
        C().fused_sgd_flat(g, master, mom, copy, hyper, None, nesterov, i == 0)
    torch.testing.assert_close(master, ref.data, rtol=1e-5, atol=1e-6)
    if copy is not None:
        torch.testing.assert_close(copy.float(), ref.data.to(cdt).float(), rtol=0, atol=0)


def test_fused_sgd_flat_skips_on_found_inf_and_unscales():
    n = 1024
    master = torch.ones(n, device="cuda")
    mom = torch.zeros(n, device="cuda")
    g = torch.full((n,), 128.0, device="cuda")
    hyper = torch.tensor([1.0, 0.0, 0.0, 0.0, 1.0 / 128.0, 0, 0, 0], device="cuda")
    flag = torch.ones(1, dtype=torch.int32, device="cuda")
    C().fused_sgd_flat(g, master, mom, None, hyper, flag, False, True)
    assert torch.all(master == 1.0)
    flag.zero_()
    C().fused_sgd_flat(g, master, mom, None, hyper, flag, False, True)
    torch.testing.assert_close(master, torch.zeros(n, device="cuda"))


def test_fused_sgd_multi_matches_torch():
    torch.manual_seed(1)
    shapes = [(64, 3, 7, 7), (64,), (1000, 512), (17,), (300000,)]
    ps = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    mine = [p.detach().clone() for p in ps]
    moms = [torch.zeros_like(m) for m in mine]
    opt = torch.optim.SGD(ps, lr=0.05, momentum=0.9, weight_decay=5e-4)
    hyper = torch.tensor([0.05, 0.9, 5e-4, 0.0, 1.0, 0, 0, 0], device="cuda")
    for it in range(3):
        gs = [torch.randn_like(p) for p in ps]
        for p, g in zip(ps, gs):
            p.grad = g.clone()
        opt.step()
        C().fused_sgd_multi(gs, mine, moms, [], hyper, None, False, it == 0)
    for a, b in zip(mine, ps):
        torch.testing.assert_close(a, b.data, rtol=1e-5, atol=1e-6)


def test_multi_tensor_scale_and_overflow_flag():
    a = [torch.randn(1000, device="cuda"), torch.randn(33, device="cuda").half()]
    out = [torch.empty_like(t) for t in a]
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    C().multi_tensor_scale(a, out, 0.5, flag)
    assert int(flag.item()) == 0
    torch.testing.assert_close(out[0], a[0] * 0.5)
    a[1][5] = float("inf")
    C().multi_tensor_scale(a, out, 0.5, flag)
    assert int(flag.item()) == 1


def test_amp_update_scale_state_machine():
    scale = torch.tensor([1024.0], device="cuda")
    tr = torch.zeros(1, dtype=torch.int32, device="cuda")
    fi = torch.zeros(1, dtype=torch.int32, device="cuda")
    hyper = torch.zeros(8, device="cuda")
    for _ in range(3):
        C().amp_update_scale(scale, tr, fi, 2.0, 0.5, 3, hyper)
    assert scale.item() == 2048.0 and tr.item() == 0
    fi.fill_(1)
    C().amp_update_scale(scale, tr, fi, 2.0, 0.5, 3, hyper)
    assert scale.item() == 1024.0 and fi.item() == 0
    assert abs(hyper[4].item() - 1.0 / 1024.0) < 1e-9


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32, torch.float16])
@pytest.mark.parametrize("C_", [64, 256, 2048, 24])
@pytest.mark.parametrize("relu,res", [(True, False), (True, True), (False, False), (False, True)])
def test_bn_act_forward_backward(dtype, C_, relu, res):
    from pytorch_distributed_b200.ops.bn_act import bn_act, bn_act_reference, begin_step
    torch.manual_seed(0)
    N, H, W = 4, 9, 7
    x = (torch.randn(N, C_, H, W, device="cuda") * 2 + 0.5).to(dtype).contiguous(memory_format=torch.channels_last)
    r = torch.randn(N, C_, H, W, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last) if res else None
    w = (torch.rand(C_, device="cuda") + 0.5)
    b = torch.randn(C_, device="cuda") * 0.1
    go = torch.randn(N, C_, H, W, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)

    def run(fn, xx, rr, ww, bb):
        xx = xx.detach().clone().requires_grad_(True)
        rr = rr.detach().clone().requires_grad_(True) if rr is not None else None
        ww = ww.detach().clone().requires_grad_(True)
        bb = bb.detach().clone().requires_grad_(True)
        rm, rv = torch.zeros(C_, device="cuda"), torch.ones(C_, device="cuda")
        y = fn(xx, ww, bb, rm, rv, residual=rr, relu=relu, training=True, momentum=0.1, eps=1e-5)
        y.backward(go.to(y.dtype))
        return y, xx.grad, (rr.grad if rr is not None else None), ww.grad, bb.grad, rm, rv

    begin_step(x.device)
    got = run(lambda *a, **k: bn_act(*a, fused=True, **k), x, r, w, b)
    ref = run(bn_act_reference, x.float(), r.float() if r is not None else None, w, b)
    tol = dict(rtol=2e-2, atol=2e-2) if dtype != torch.float32 else dict(rtol=1e-4, atol=1e-4)
    names = ["y", "dx", "dres", "dw", "db", "running_mean", "running_var"]
    for name, g_, r_ in zip(names, got, ref):
        if g_ is None:
            assert r_ is None
            continue
        scale = max(1.0, float(r_.float().abs().max()))
        t = {k: v * (scale if name in ("dw", "db") else 1.0) for k, v in tol.items()}
        torch.testing.assert_close(g_.float(), r_.float(), msg=lambda m, n=name: n + ": " + m, **t)


def test_bn_act_eval_uses_running_stats():
    from pytorch_distributed_b200.ops.bn_act import bn_act, bn_act_reference
    x = torch.randn(2, 64, 5, 5, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    w, b = torch.rand(64, device="cuda") + 0.5, torch.randn(64, device="cuda")
    rm, rv = torch.randn(64, device="cuda"), torch.rand(64, device="cuda") + 0.5
    y = bn_act(x, w, b, rm, rv, relu=True, training=False, fused=True)
    ref = bn_act_reference(x.float(), w, b, rm, rv, relu=True, training=False)
    torch.testing.assert_close(y.float(), ref, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("src_dtype", [torch.uint8, torch.float32])
@pytest.mark.parametrize("out", [(torch.bfloat16, True), (torch.float32, False), (torch.float16, True)])
def test_normalize_kernel(src_dtype, out):
    odt, cl = out
    if src_dtype == torch.uint8:
        src = torch.randint(0, 256, (3, 3, 17, 13), dtype=torch.uint8, device="cuda")
    else:
        src = torch.randn(3, 3, 17, 13, device="cuda")
    a = torch.tensor([0.5, 2.0, 1.5], device="cuda")
    b = torch.tensor([0.1, -0.2, 0.3], device="cuda")
    code = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[odt]
    y = C().normalize_nhwc(src, a, b, code, cl)
    ref = src.float() * a.view(1, 3, 1, 1) + b.view(1, 3, 1, 1)
    assert y.dtype == odt and y.shape == src.shape
    assert y.is_contiguous(memory_format=torch.channels_last if cl else torch.contiguous_format)
    tol = 1e-5 if odt == torch.float32 else 1e-2      # kernel uses one FMA; the oracle rounds the product first
    torch.testing.assert_close(y.float(), ref.to(odt).float(), rtol=tol, atol=tol)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_metrics_kernel_world1_matches_oracle(dtype):
    from pytorch_distributed_b200.parallel.comm import FusedCommunicator
    from pytorch_distributed_b200.utils.meters import accuracy
    comm = FusedCommunicator(device=torch.device("cuda", 0), arena_bytes=8 << 20)
    torch.manual_seed(0)
    logits = torch.randn(77, 1000, device="cuda").to(dtype)
    target = torch.randint(0, 1000, (77,), device="cuda")
    logits[torch.arange(20), target[:20]] += 20      # make some of them correct
    loss = torch.tensor(1.25, device="cuda")
    out = torch.zeros(4, device="cuda")
    comm.metrics(logits, target, loss, out)
    a1, a5 = accuracy(logits, target, (1, 5))
    assert abs(out[0].item() - 1.25) < 1e-6
    assert abs(out[1].item() - a1.item()) < 1e-4 and abs(out[2].item() - a5.item()) < 1e-4
    comm.check()


@pytest.mark.parametrize("wire", ["bf16", "fp32", "fp16"])
def test_fused_allreduce_world1_is_identity_and_fills_arena(wire):
    """World size 1 exercises pack -> (self) reduce -> unpack and the arena layout used by the flat optimizer."""
    from pytorch_distributed_b200.parallel.comm import KIND_TWO_SHOT, FusedCommunicator
    comm = FusedCommunicator(device=torch.device("cuda", 0), arena_bytes=64 << 20)
    torch.manual_seed(0)
    shapes = [(64, 3, 7, 7), (64,), (5,), (1000, 2048), (2048,), (33, 17)]
    ts = [torch.randn(s, device="cuda") for s in shapes]
    ts[3] = ts[3].bfloat16()
    orig = [t.clone() for t in ts]
    plan = comm.make_plan([t.numel() for t in ts], wire)
    comm.run(plan, ts, KIND_TWO_SHOT, comm.misc_channel, scale=0.5, writeback=True)
    torch.cuda.synchronize()
    comm.check()
    wdt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[wire]
    arena = plan.arena_tensor()
    for t, o, off in zip(ts, orig, plan.layout.offsets):
        exp = (o.float() * 0.5).to(wdt)
        torch.testing.assert_close(t.float(), exp.to(t.dtype).float(), rtol=0, atol=0)
        torch.testing.assert_close(arena[off:off + o.numel()].float(), exp.float().reshape(-1), rtol=0, atol=0)


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()
