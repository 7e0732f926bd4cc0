"""-m gpu: bf16 STORAGE variant (fp32 accumulate) of the two benchmarked ops.  Bound (north_star: "a
stated bf16 bound"): against an fp64 result computed from the SAME bf16-rounded inputs the only
extra error is one bf16 rounding of the output, |a-b| <= 2^-8 * |b| + fp32-accumulation slack."""
import numpy as np
import pytest
import torch

import dgl
from gpu_util import graphs

pytestmark = pytest.mark.gpu


def _close_bf16(got, want64, scale64):
    got = got.double().cpu().numpy()
    err = np.abs(got - want64)
    bound = 2.0 ** -8 * np.abs(want64) + 1e-5 * scale64 + 1e-30
    assert (err <= bound).all(), float((err / bound).max())


@pytest.mark.parametrize("D", [1, 2, 6, 8, 24, 64, 100, 128, 256, 602, 1024])
@pytest.mark.parametrize("kind", ["uniform", "powerlaw"])
def test_copy_u_sum_and_mean_bf16(oracle, cuda, D, kind, small_hub_threshold):
    og, g, src, dst = graphs(oracle, 400, 400, 12000, seed=D, kind=kind)
    X = torch.rand(400, D, device=cuda).to(torch.bfloat16)
    X64 = X.double().cpu().numpy()
    want = np.zeros((400, D))
    np.add.at(want, dst, X64[src])
    out = dgl.ops.gspmm(g, "copy_lhs", "sum", X, None)
    assert out.dtype == torch.bfloat16 and out.shape == (400, D)
    _close_bf16(out, want, want)
    deg = np.maximum(np.bincount(dst, minlength=400), 1)[:, None]
    _close_bf16(dgl.ops.gspmm(g, "copy_lhs", "mean", X, None), want / deg, want / deg)


@pytest.mark.parametrize("D", [1, 4, 8, 16, 64, 128, 602, 1024])
def test_u_dot_v_bf16(oracle, cuda, D, small_hub_threshold):
    og, g, src, dst = graphs(oracle, 300, 300, 9000, seed=D + 1, kind="powerlaw")
    U = torch.randn(300, D, device=cuda).to(torch.bfloat16)
    V = torch.randn(300, D, device=cuda).to(torch.bfloat16)
    U64, V64 = U.double().cpu().numpy(), V.double().cpu().numpy()
    want = (U64[src] * V64[dst]).sum(-1, keepdims=True)
    scale = (np.abs(U64[src]) * np.abs(V64[dst])).sum(-1, keepdims=True)
    out = dgl.ops.gsddmm(g, "dot", U, V)
    assert out.dtype == torch.bfloat16 and out.shape == (9000, 1)
    _close_bf16(out, want, scale)


def test_bf16_multi_head_dot_and_backward(oracle, cuda):
    og, g, src, dst = graphs(oracle, 200, 200, 4000, seed=3)
    U = torch.randn(200, 4, 16, device=cuda).to(torch.bfloat16)
    V = torch.randn(200, 4, 16, device=cuda).to(torch.bfloat16)
    out = dgl.ops.gsddmm(g, "dot", U, V)
    want = (U.double().cpu().numpy()[src] * V.double().cpu().numpy()[dst]).sum(-1, keepdims=True)
    scale = (np.abs(U.double().cpu().numpy()[src]) * np.abs(V.double().cpu().numpy()[dst])).sum(-1, keepdims=True)
    _close_bf16(out, want, scale)
    # SAGE-style aggregation backward in bf16: dX = A^T dZ through the same kernel
    X = torch.rand(200, 32, device=cuda).to(torch.bfloat16).requires_grad_(True)
    y = dgl.ops.gspmm(g, "copy_lhs", "sum", X, None)
    dZ = torch.randn(200, 32, device=cuda).to(torch.bfloat16)
    y.backward(dZ)
    want_dx = np.zeros((200, 32))
    np.add.at(want_dx, src, dZ.double().cpu().numpy()[dst])
    sc = np.zeros((200, 32))
    np.add.at(sc, src, np.abs(dZ.double().cpu().numpy()[dst]))
    _close_bf16(X.grad, want_dx, sc)


def test_unsupported_bf16_ops_fail_loudly(cuda):
    g = dgl.graph((torch.tensor([0, 1]), torch.tensor([1, 0]))).int().to(cuda)
    x = torch.rand(2, 8, device=cuda).to(torch.bfloat16)
    w = torch.rand(2, 8, device=cuda).to(torch.bfloat16)
    with pytest.raises(dgl.DGLError, match="bfloat16"):
        dgl.ops.gspmm(g, "mul", "sum", x, w)
    with pytest.raises(dgl.DGLError, match="bfloat16"):
        dgl.ops.gspmm(g, "copy_lhs", "max", x, None)
    with pytest.raises(dgl.DGLError, match="bfloat16"):
        dgl.ops.gsddmm(g, "add", x, x)
    with pytest.raises(dgl.DGLError):
        dgl.ops.gspmm(g, "copy_lhs", "sum", x.to(torch.float16), None)
