import copy, os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from pytorch_distributed_b200.parallel.comm import FusedCommunicator
from pytorch_distributed_b200.models import create_model
from pytorch_distributed_b200.ops.fused_sgd import FusedSGD
from pytorch_distributed_b200.parallel.ddp import DistributedDataParallel
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.deterministic = True
comm = FusedCommunicator(device=dev, arena_bytes=256 << 20)
torch.manual_seed(7)
base = create_model("resnet18", num_classes=10, fused_bn=False).to(dev)
plain = copy.deepcopy(base)
own = DistributedDataParallel(copy.deepcopy(base), device_ids=[local], comm=comm, wire_dtype="fp32")
names = [n for n, _ in plain.named_parameters()]
crit = torch.nn.CrossEntropyLoss()
opt = FusedSGD(own.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
popt = torch.optim.SGD(plain.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
eng = own.engine
torch.manual_seed(100 + rank)
for it in range(4):
    x = torch.randn(8, 3, 64, 64, device=dev); y = torch.randint(0, 10, (8,), device=dev)
    with torch.no_grad():   # same weights (and BN buffers) going in => per-iteration comparison, no chaos accumulation
        for a, b in zip(plain.parameters(), own.module.parameters()): a.copy_(b)
        for a, b in zip(plain.buffers(), own.module.buffers()): a.copy_(b)
    popt.zero_grad(); lp = crit(plain(x), y); lp.backward()
    for p in plain.parameters():
        dist.all_reduce(p.grad); p.grad /= world
    opt.zero_grad(); lo = crit(own(x), y); lo.backward()
    eng.wait_for_gradients(); torch.cuda.synchronize()
    arena = eng.grad_arena()
    ge = sorted([(names[i], (arena[eng.param_elem_off[i]:eng.param_elem_off[i] + p.numel()].view_as(p) - q.grad).abs().max().item(), q.grad.abs().max().item())
                 for i, (p, q) in enumerate(zip(eng.params, plain.parameters()))], key=lambda t: -t[1])[:3]
    popt.step(); opt.step(); torch.cuda.synchronize()
    pe = sorted([(n, (a - b).abs().max().item()) for n, a, b in zip(names, plain.parameters(), own.module.parameters())], key=lambda t: -t[1])[:3]
    me = sorted([(n, (popt.state[a]["momentum_buffer"] - opt.state[b]["momentum_buffer"]).abs().max().item()) for n, a, b in zip(names, plain.parameters(), own.module.parameters())], key=lambda t: -t[1])[:2]
    print("rank", rank, "iter", it, "loss", lp.item(), lo.item(), "\n   grad(arena) err", ge, "\n   param err", pe, "\n   mom err", me, flush=True)
dist.barrier(); dist.destroy_process_group()
