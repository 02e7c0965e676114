"""Host-side scale / zero-point arithmetic on the tiny [C] / [O] vectors
(reference: utils/quantization_utils/quant_utils.py).  Full-tensor work never
goes through these; it runs in the CUDA kernels."""
import torch
from torch.autograd import Function


def asymmetric_linear_quantization_params(num_bits, saturation_min, saturation_max, integral_zero_point=True,
                                          signed=True):
    """quant_utils.py:109-133."""
    n = 2 ** num_bits - 1
    scale = n / (saturation_max - saturation_min)
    zero_point = scale * saturation_min
    if integral_zero_point:
        if isinstance(zero_point, torch.Tensor):
            zero_point = zero_point.round()
        else:
            zero_point = float(round(zero_point))
    if signed:
        zero_point = zero_point + 2 ** (num_bits - 1)
    return scale, zero_point


def _bcast(v, x):
    if x.dim() == 4:
        return v.view(-1, 1, 1, 1)
    if x.dim() == 2:
        return v.view(-1, 1)
    return v


def linear_quantize(input, scale, zero_point, inplace=False):
    """quant_utils.py:62-83."""
    return torch.round(_bcast(scale, input) * input - _bcast(zero_point, input))


def linear_dequantize(input, scale, zero_point, inplace=False):
    """quant_utils.py:86-106."""
    return (input + _bcast(zero_point, input)) / _bcast(scale, input)


class AsymmetricQuantFunction(Function):
    """quant_utils.py:136-167 (used once per layer to put weights on the w_bit grid)."""

    @staticmethod
    def forward(ctx, x, k, x_min=None, x_max=None):
        if x_min is None or x_max is None or (x_min.numel() == 1 and bool((x_min == x_max).all())):
            x_min, x_max = x.min(), x.max()
        scale, zero_point = asymmetric_linear_quantization_params(k, x_min, x_max)
        q = linear_quantize(x, scale, zero_point)
        n = 2 ** (k - 1)
        q = torch.clamp(q, -n, n - 1)
        return linear_dequantize(q, scale, zero_point)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output, None, None, None
