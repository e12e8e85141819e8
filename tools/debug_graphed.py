import sys, os, traceback, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_distributed_b200.models import create_model
from pytorch_distributed_b200.parallel.amp import cast_model
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
for arch, dt in (("resnet18", torch.float32), ("resnet18", torch.bfloat16), ("resnet50", torch.bfloat16)):
    try:
        m = create_model(arch, num_classes=10).to(dev).to(memory_format=torch.channels_last)
        if dt != torch.float32:
            cast_model(m, dt)
        m.train()
        x = torch.randn(8, 3, 64, 64, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
        for _ in range(2):
            m(x).float().sum().backward()
        torch.cuda.synchronize()
        g = torch.cuda.make_graphed_callables(m, (x.clone(),))
        out = g(x)
        out.float().sum().backward()
        torch.cuda.synchronize()
        print("OK", arch, dt, float(out.float().abs().mean()))
    except Exception:
        print("FAIL", arch, dt)
        traceback.print_exc()
        break
