"""CPU test of the gradient-chaining protocol itself (omni3d_b200.nnfunc.fork / _GradSink / CatChannels): stand-in consumers
written with plain torch ops follow the protocol (first arrival parks its buffer, later arrivals add into it in place and
return None) next to consumers that know nothing of it (plain autograd ops), across nested forks.  The gradients must equal
plain autograd's.  The CUDA kernels behind the real consumers are tested in tests/test_grad_chain_gpu.py."""
import pytest
import torch

from omni3d_b200 import nnfunc as F


class _Aware(torch.autograd.Function):
    """y = k * x; the backward follows the sink protocol the way ConvBNAct / ConvBias / MaxPool2 do."""
    calls = {"parked": 0, "accumulated": 0}

    @staticmethod
    def forward(ctx, x, k):
        ctx.sink, ctx.k = F._sink_of(x), k
        return x * k

    @staticmethod
    def backward(ctx, g):
        s = ctx.sink
        if s is not None and s.buf is not None:
            s.buf.add_(g * ctx.k)                       # "inside the producing kernel"
            _Aware.calls["accumulated"] += 1
            return None, None
        _Aware.calls["parked"] += int(s is not None)
        return F._deliver(s, g * ctx.k), None


def _graph(chain, seed=0):
    F.GRAD_CHAIN = chain
    try:
        torch.manual_seed(seed)
        x = torch.randn(2, 3, 3, 4, requires_grad=True)
        w = torch.randn(2, 3, 3, 4)
        t = x * 1.5
        a, b, c, d = F.fork(t, 4)
        a1, a2 = F.fork(a, 2)                            # nested fork: shares the outer sink
        y1 = _Aware.apply(a1, 2.0)
        y2 = _Aware.apply(a2, 3.0)
        z = F.CatChannels.apply(y1, b)                   # b reaches the sink through the cat's channel slice
        u = c[:, ::2]                                    # a consumer outside the protocol (plain autograd op)
        v = _Aware.apply(d, -1.0)
        loss = (z ** 2).sum() + (y2 * w).sum() + (u * 5).sum() + v.sum()
        loss.backward()
        return x.grad.clone()
    finally:
        F.GRAD_CHAIN = True


def test_chained_gradients_equal_plain_autograd():
    ref = _graph(False)
    _Aware.calls.update(parked=0, accumulated=0)
    got = _graph(True)
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-6)
    # the protocol was really exercised: one buffer parked, the other aware consumers added into it
    assert _Aware.calls["accumulated"] >= 2


def test_fork_is_transparent_without_grad():
    t = torch.randn(3, 4)
    assert all(o is t for o in F.fork(t, 3))                              # no grad needed: the tensor itself
    t.requires_grad_(True)
    with torch.no_grad():
        assert all(o is t for o in F.fork(t, 2))
    outs = F.fork(t, 2)
    assert all(o is not t and o.data_ptr() == t.data_ptr() for o in outs)  # aliases, no copies
    assert F._sink_of(outs[0]) is F._sink_of(outs[1]) is not None


def test_unused_aliases_and_single_consumer():
    torch.manual_seed(1)
    x = torch.randn(4, 5, requires_grad=True)
    a, b, c = F.fork(x * 2.0, 3)                         # b and c are never consumed
    _Aware.apply(a, 4.0).sum().backward()
    assert torch.allclose(x.grad, torch.full_like(x, 8.0))


@pytest.mark.parametrize("I,O,want", [(16, 32, True), (64, 128, True), (128, 256, True), (256, 512, False), (4, 32, False)])
def test_merge_rule_of_the_stride2_data_gradient(I, O, want):
    assert bool(F._merge_phases(I, O)) == want
