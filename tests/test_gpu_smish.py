"""Smish activation kernel (models/local_stage.py:4-6) against the eager expression of the reference in fp64."""
import pytest
import torch
import torch.nn as nn

import synth

pytestmark = pytest.mark.gpu


def smish_ref(x):      # models/local_stage.py:5-6, verbatim formula
    return x * torch.tanh(torch.log(1 + torch.sigmoid(x)))


@pytest.mark.parametrize('shape', [(4, 64, 21, 21), (3, 5, 7), (1,), (1023,), (0,)])
def test_smish_forward_backward_match_eager_fp64(shape):
    from blurry_edges_b200 import smish
    n = 1
    for d in shape:
        n *= d
    x = (synth.uniform((max(n, 1),), 7, -12.0, 12.0)[:n]).reshape(shape)
    x64 = x.double().requires_grad_(True)
    y64 = smish_ref(x64)
    g = synth.uniform((max(n, 1),), 8, -1.0, 1.0)[:n].reshape(shape)
    y64.backward(g.double())
    xc = x.cuda().requires_grad_(True)
    y = smish(xc)
    y.backward(g.cuda())
    assert y.shape == x.shape and xc.grad.shape == x.shape
    if n:
        # fp32 with MUFU exp2/rcp (<= 2 ulp each): 2e-6 of the value scale
        assert float((y.detach().cpu().double() - y64.detach()).abs().max()) < 2e-6 * max(1.0, float(y64.abs().max()))
        assert float((xc.grad.cpu().double() - x64.grad).abs().max()) < 3e-6 * max(1.0, float(x64.grad.abs().max()))


def test_smish_extremes_and_module_in_a_cnn():
    from blurry_edges_b200 import SmishFused, BlurryEdgesError
    x = torch.tensor([-200.0, -30.0, -1e-8, 0.0, 1e-8, 30.0, 200.0], device='cuda')
    y = SmishFused()(x)
    ref = smish_ref(x.double().cpu())
    assert torch.isfinite(y).all() and float((y.cpu().double() - ref).abs().max()) < 1e-4      # 0.6 x at +200
    with pytest.raises(BlurryEdgesError):
        SmishFused()(torch.zeros(3))

    class Eager(nn.Module):
        def forward(self, t):
            return smish_ref(t)

    def net(act):
        torch.manual_seed(0)
        return nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.BatchNorm2d(8), act(), nn.Conv2d(8, 4, 3, padding=1), act()).cuda()

    a, b = net(SmishFused), net(Eager)
    inp = synth.uniform((16, 3, 21, 21), 9, 0.0, 1.0).cuda()
    ya, yb = a(inp), b(inp)
    ya.square().mean().backward()
    yb.square().mean().backward()
    assert float((ya - yb).abs().max()) < 1e-5
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert float((pa.grad - pb.grad).abs().max()) < 1e-5 * max(1.0, float(pb.grad.abs().max()))
