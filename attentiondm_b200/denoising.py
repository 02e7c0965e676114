"""Drop-in for functions/denoising.py::generalized_steps (the DDIM sample loop).

Same signature and return value as the reference (:16-42): (xs, x0_preds), lists
of CPU tensors (xs[0] is the caller's x).  The loop body runs on the device:
the UNet forward through the fused kernels, one attndm_ddim_step per step, the
alpha-bar table computed once instead of a cumprod per step (:8-11), and the
per-step `.to('cpu')` copies (:34,40) replaced by async copies into pinned
buffers with one synchronisation at the end.

When `model` is an attentiondm_b200 Model in inference mode the whole step
(table staging -> UNet -> DDIM update) is captured once in a CUDA graph and
replayed for every step (engine.SamplerEngine); otherwise (calibration mode, or
any other callable) the loop runs eagerly with the same kernels.
"""
import torch

from . import ops


def compute_alpha(beta, t):
    """:8-11."""
    beta = torch.cat([torch.zeros(1).to(beta.device), beta], dim=0)
    return (1 - beta).cumprod(dim=0).index_select(0, t + 1).view(-1, 1, 1, 1)


def ddim_coefficients(seq, b, eta=0.0):
    """[T,8] fp32 table (CPU), row k = step k of the reversed sequence:
    {sqrt(1-at), sqrt(at), sqrt(at_next), c1, c2, t, 0, 0} with the reference's fp32
    op order (:26-27,33,35-38)."""
    seq = list(seq)
    seq_next = [-1] + seq[:-1]
    bc = b.detach().float().cpu()
    rows = []
    for i, j in zip(reversed(seq), reversed(seq_next)):
        at = compute_alpha(bc, torch.tensor([i]).long()).view(())
        an = compute_alpha(bc, torch.tensor([j]).long()).view(())
        c1 = eta * ((1 - at / an) * (1 - an) / (1 - at)).sqrt()
        c2 = ((1 - an) - c1 ** 2).sqrt()
        rows.append(torch.stack([(1 - at).sqrt(), at.sqrt(), an.sqrt(), torch.as_tensor(c1, dtype=torch.float32), c2,
                                 torch.tensor(float(i)), torch.tensor(0.0), torch.tensor(0.0)]))
    return torch.stack(rows).float()


def _is_engine_model(model):
    from .diffusion import Model
    return isinstance(model, Model) and not any(m._calibrate for _, m in model.qconvs())


def generalized_steps(x, seq, model, b, **kwargs):
    """kwargs: eta (reference), plus keep='all'|'last' (default 'all', the reference behaviour),
    use_graph (default True), noise_fn (tests: callable(step, like) -> noise)."""
    eta = kwargs.get("eta", 0)
    keep = kwargs.get("keep", "all")
    noise_fn = kwargs.get("noise_fn", None)
    if not x.is_cuda:
        raise RuntimeError("attentiondm_b200.generalized_steps: x must be a CUDA tensor (no CPU fallback)")
    with torch.no_grad():
        if _is_engine_model(model) and kwargs.get("use_graph", True):
            from .engine import SamplerEngine
            eng = SamplerEngine.for_model(model, seq, b, eta, tuple(x.shape))
            return eng.run(x, keep=keep, noise_fn=noise_fn)
        seq = list(seq)
        coef = ddim_coefficients(seq, b, eta).to(x.device)
        n = x.size(0)
        xt = ops.to_nhwc(x).clone()
        host_x, host_x0 = [], []
        for k in range(len(seq)):
            t = torch.full((n,), float(seq[len(seq) - 1 - k]), device=x.device)
            et = ops.to_nhwc(model(ops.to_nchw(xt), t))
            noise = None
            if eta != 0 or noise_fn is not None:
                nz = noise_fn(k, ops.to_nchw(xt)) if noise_fn is not None else torch.randn_like(ops.to_nchw(xt))
                noise = ops.to_nhwc(nz)
            xt_next, x0 = ops.ddim_step(xt, et, coef[k], noise, want_x0=True)
            if keep == "all" or k == len(seq) - 1:
                host_x.append(_to_host_async(xt_next))
                host_x0.append(_to_host_async(x0))
            xt = xt_next
        torch.cuda.current_stream().synchronize()
        xs = [x] + [ops.to_nchw(h) for h in host_x]
        x0s = [ops.to_nchw(h) for h in host_x0]
        return xs, x0s


def _to_host_async(t):
    h = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
    h.copy_(t, non_blocking=True)
    return h
