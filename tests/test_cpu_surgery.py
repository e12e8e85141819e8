import copy

import pytest
import torch


@pytest.mark.parametrize("arch,min_pairs", [("vgg11_bn", 8), ("regnet_y_400mf", 10), ("shufflenet_v2_x0_5", 5)])
def test_fuse_bn_relu_preserves_outputs_and_keys(arch, min_pairs):
    import torchvision.models as tvm
    from pytorch_distributed_b200.models.resnet import BNAct
    from pytorch_distributed_b200.models.surgery import fuse_bn_relu
    torch.manual_seed(0)
    ref = tvm.__dict__[arch](num_classes=7)
    new = copy.deepcopy(ref)
    n = fuse_bn_relu(new)
    assert n >= min_pairs and sum(isinstance(m, BNAct) for m in new.modules()) == n
    assert list(new.state_dict().keys()) == list(ref.state_dict().keys())
    x = torch.randn(2, 3, 64, 64)
    for train in (True, False):
        ref.train(train); new.train(train)
        torch.manual_seed(1); ya = new(x)          # same dropout masks (VGG classifier) in both runs
        torch.manual_seed(1); yb = ref(x)
        torch.testing.assert_close(ya, yb, rtol=1e-5, atol=1e-5)
    for (k, a), b in zip(ref.state_dict().items(), new.state_dict().values()):   # running stats advanced identically
        torch.testing.assert_close(a.float(), b.float(), rtol=1e-5, atol=1e-6, msg=lambda m, k=k: k + ": " + m)
    # gradients flow to the same Parameter set
    new.train()
    new(x).sum().backward()
    assert all(p.grad is not None for p in new.parameters())


def test_create_model_fused_flag_on_zoo_model():
    from pytorch_distributed_b200.models import create_model
    from pytorch_distributed_b200.models.resnet import BNAct
    m = create_model("vgg11_bn", num_classes=5, fused_bn=True)
    assert any(isinstance(x, BNAct) for x in m.modules())
    m2 = create_model("vgg11_bn", num_classes=5)
    assert not any(isinstance(x, BNAct) for x in m2.modules())
