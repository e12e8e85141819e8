"""Probe: cuDNN 1x1 convolution vs cuBLAS GEMM on the same data (NHWC bf16, batch 256), forward / dgrad / wgrad."""
import torch
import torch.nn.functional as F
torch.backends.cudnn.benchmark = True
B = 256
shapes = [(64, 64, 56), (64, 256, 56), (256, 64, 56), (256, 128, 56), (128, 512, 28), (512, 128, 28), (512, 256, 28), (256, 1024, 14),
          (1024, 256, 14), (1024, 512, 14), (512, 2048, 7), (2048, 512, 7)]


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


tot = [0.0] * 6
print("cin cout hw | conv fwd / dgrad / wgrad us | gemm fwd / dgrad / wgrad us")
for cin, cout, hw in shapes:
    x = torch.randn(B, cin, hw, hw, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.randn(cout, cin, 1, 1, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dy = torch.randn(B, cout, hw, hw, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    x2, w2, dy2 = x.permute(0, 2, 3, 1).reshape(-1, cin), w.reshape(cout, cin), dy.permute(0, 2, 3, 1).reshape(-1, cout)
    cf = t(lambda: F.conv2d(x, w))
    cd = t(lambda: torch.ops.aten.convolution_backward(dy, x, w, None, (1, 1), (0, 0), (1, 1), False, (0, 0), 1, (True, False, False)))
    cw = t(lambda: torch.ops.aten.convolution_backward(dy, x, w, None, (1, 1), (0, 0), (1, 1), False, (0, 0), 1, (False, True, False)))
    gf = t(lambda: x2 @ w2.t())
    gd = t(lambda: dy2 @ w2)
    gw = t(lambda: dy2.t() @ x2)
    for i, v in enumerate((cf, cd, cw, gf, gd, gw)):
        tot[i] += v
    print("%4d %4d %3d | %7.1f %7.1f %7.1f | %7.1f %7.1f %7.1f" % (cin, cout, hw, cf, cd, cw, gf, gd, gw), flush=True)
print("sum          | %7.1f %7.1f %7.1f | %7.1f %7.1f %7.1f" % tuple(tot))
