"""Stem convolution as im2col + GEMM (ops/stem_conv.py, csrc/stem_conv.cu): layout definition against F.conv2d, a line-by-line
NumPy transcription of the CUDA kernel's index arithmetic against that definition, autograd, and the model switch."""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import pytorch_distributed_b200.models.resnet as R
from pytorch_distributed_b200.models import create_model
from pytorch_distributed_b200.ops import stem_conv as S


@pytest.mark.parametrize("hw", [(32, 32), (37, 45), (64, 20)])
def test_im2col_gemm_equals_conv2d(hw):
    torch.manual_seed(0)
    x = torch.randn(2, 3, *hw).contiguous(memory_format=torch.channels_last)
    wt = torch.randn(64, 3, 7, 7).contiguous(memory_format=torch.channels_last)
    a, wp = S.im2col_reference(x), S.pack_stem_weight(wt)
    assert a.shape[1] == 192 and a.is_contiguous(memory_format=torch.channels_last) and wp.shape == (64, 192)
    y = (a.permute(0, 2, 3, 1).reshape(-1, 192) @ wp.t()).reshape(2, a.size(2), a.size(3), 64).permute(0, 3, 1, 2)
    assert torch.allclose(y, F.conv2d(x, wt, stride=2, padding=3), atol=1e-3, rtol=1e-4)
    assert torch.equal(S.unpack_stem_weight(wp, wt), wt)
    assert (wp.reshape(64, 8, 24)[:, :7, 21:] == 0).all() and (wp[:, 168:] == 0).all()


def _kernel_transcription(x_nhwc: np.ndarray) -> np.ndarray:
    """stem_im2col_kernel, thread by thread: i -> (pixel, granule) -> 8 elements (same variable names as the .cu file)."""
    n_img, H, W, _ = x_nhwc.shape
    OH, OW = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    flat = x_nhwc.reshape(-1)
    row_elems = W * 3
    total = n_img * OH * OW * 24
    a = np.zeros(total * 8, dtype=x_nhwc.dtype)
    for i in range(total):
        gq, pix = i % 24, i // 24
        ow, t = pix % OW, pix // OW
        oh, n = t % OH, t // OH
        r = gq // 3
        q = gq - 3 * r
        ih = 2 * oh - 3 + r
        if r < 7 and 0 <= ih < H:
            row = (n * H + ih) * row_elems
            e0 = (2 * ow - 3) * 3 + q * 8
            for j in range(8):
                sc, e = q * 8 + j, e0 + j
                if sc < 21 and 0 <= e < row_elems:
                    a[i * 8 + j] = flat[row + e]
    return a.reshape(n_img, OH, OW, 192)


def test_cuda_kernel_index_arithmetic_matches_definition():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 11, 14, 3)).astype(np.float32)              # N H W C
    got = _kernel_transcription(x)
    ref = S.im2col_reference(torch.from_numpy(x).permute(0, 3, 1, 2)).permute(0, 2, 3, 1).numpy()
    assert np.array_equal(got, ref)


def test_stem_conv_function_gradients_and_statistics():
    torch.manual_seed(1)
    x = torch.randn(2, 3, 40, 40).contiguous(memory_format=torch.channels_last)
    w1 = torch.randn(64, 3, 7, 7).contiguous(memory_format=torch.channels_last).requires_grad_()
    w2 = w1.detach().clone().requires_grad_()
    stats = torch.zeros(128)
    y1 = S._StemConvFn.apply(x, w1, stats, True)
    y2 = F.conv2d(x, w2, stride=2, padding=3)
    g = torch.randn_like(y2)
    y1.backward(g)
    y2.backward(g)
    assert torch.allclose(y1, y2, atol=1e-3, rtol=1e-4)
    assert torch.allclose(w1.grad, w2.grad, atol=1e-3, rtol=1e-4) and w1.grad.stride() == w1.stride()
    assert torch.allclose(stats[:64], y2.detach().sum((0, 2, 3)), atol=1e-2, rtol=1e-4)
    assert torch.allclose(stats[64:], (y2.detach() ** 2).sum((0, 2, 3)), rtol=1e-4)


def test_model_switch_keeps_results():
    torch.manual_seed(0)
    base = create_model("resnet18", num_classes=10, fused_bn=False).float().to(memory_format=torch.channels_last)
    for m in base.modules():
        if isinstance(m, R.BNAct):
            m.fused = "emulate"
    x = torch.randn(4, 3, 64, 64).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 10, (4,))
    outs = []
    for flag in (False, True):
        m = copy.deepcopy(base).train()
        R.STEM_GEMM = flag
        try:
            o = m(x)
            F.cross_entropy(o, y).backward()
        finally:
            R.STEM_GEMM = False
        outs.append((o.detach(), m.conv1.weight.grad.clone(), m.bn1.running_var.clone(), int(m.bn1.num_batches_tracked)))
    assert torch.allclose(outs[0][0], outs[1][0], atol=1e-3, rtol=1e-3)
    assert (outs[0][1] - outs[1][1]).abs().max() / outs[0][1].abs().max() < 1e-3
    assert torch.allclose(outs[0][2], outs[1][2], rtol=1e-4) and outs[1][3] == 1
