"""Micro-probe: cuDNN 7x7/2 stem conv fwd+bwd time for C_in = 3 vs 4 vs 8 (NHWC bf16, batch 256)."""
import torch
import torch.nn.functional as F
torch.backends.cudnn.benchmark = True
for cin in (3, 4, 8):
    x = torch.randn(256, cin, 224, 224, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.randn(64, cin, 7, 7, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    for _ in range(3):
        y = F.conv2d(x, w, stride=2, padding=3); y.sum().backward()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    go = torch.randn_like(y)
    e0.record()
    for _ in range(10):
        y = F.conv2d(x, w, stride=2, padding=3)
    e1.record()
    for _ in range(10):
        y = F.conv2d(x, w, stride=2, padding=3); y.backward(go)
    e2.record(); torch.cuda.synchronize()
    print("cin=%d fwd %.3f ms, fwd+bwd(wgrad only) %.3f ms" % (cin, e0.elapsed_time(e1) / 10, e1.elapsed_time(e2) / 10), flush=True)
