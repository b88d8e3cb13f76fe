"""Stand-in for flash-attention's optional `fused_dense_lib` CUDA extension (csrc/fused_dense_lib), which the image's
flash_attn 2.8 wheel does not ship and which cannot be built offline (no flash-attention source tree).

The real extension is a thin wrapper over cuBLASLt GEMMs with bias / activation epilogues.  This module offers the
same four entry points on top of torch's own cuBLAS(Lt) calls, so the *reference* arm of bench.py
(Galvatron -> flash_attn.models.gpt -> flash_attn.ops.fused_dense) runs its stock code path.  Nothing of hetu_b200
is involved here: every call below is plain PyTorch.
"""
import torch
import torch.nn.functional as F


def linear_bias_forward(x, weight, bias):
    return F.linear(x, weight, bias)


def linear_bias_backward(x, weight, grad_output):
    grad_input = grad_output @ weight
    grad_weight = grad_output.t() @ x
    grad_bias = grad_output.sum(dim=0)
    return grad_input, grad_weight, grad_bias


def linear_bias_wgrad(x, grad_output, has_d_bias):
    grad_weight = grad_output.t() @ x
    grad_bias = grad_output.sum(dim=0) if has_d_bias else None
    return grad_weight, grad_bias


def linear_act_forward(x, weight, bias, is_gelu, save_pre_act, heuristic):
    pre_act = F.linear(x, weight, bias)
    out = F.gelu(pre_act, approximate="tanh") if is_gelu else F.relu(pre_act)
    if save_pre_act:
        # the cuBLASLt relu epilogue stores a bit mask, the python side only hands it back to
        # bias_act_linear_dgrad_bgrad below, so keeping the pre-activation itself is equivalent
        return out, pre_act
    return (out,)


def bias_act_linear_dgrad_bgrad(weight, grad_output, pre_act, is_gelu, heuristic):
    grad_act = grad_output @ weight
    if is_gelu:
        with torch.enable_grad():
            p = pre_act.detach().requires_grad_(True)
            y = F.gelu(p, approximate="tanh")
        (grad_pre_act,) = torch.autograd.grad(y, p, grad_act)
    else:
        grad_pre_act = grad_act * (pre_act > 0).to(grad_act.dtype)
    return grad_pre_act, grad_pre_act.sum(dim=0)
