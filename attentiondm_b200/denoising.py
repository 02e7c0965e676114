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
    qs = model.qconvs() if isinstance(model, Model) else []
    return len(qs) > 0 and not any(m._calibrate for _, m in qs)      # the FP model (no QConv2d) runs eagerly


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


# ---------------------------------------------------------------------------------------------
# The callers either side of the sampler: functions/denoising.py:13-14,45-151
# ---------------------------------------------------------------------------------------------
def cal_entropy(attn):
    """:13-14 (host-side twin of attndm_alpha_entropy_grad's value; tiny [G,C] tensors)."""
    return -1 * torch.sum((attn * torch.log(attn)), dim=-1).mean()


def _scalar_f32(v):
    return torch.as_tensor(v, dtype=torch.float32).reshape(())


def noise_estimation_loss(model, x0, t, e, b, keepdim=False):
    """:45-60.  x0, e: logical NCHW CUDA tensors; t: [B] (all samples of a call share the timestep on this path, as
    in generalized_steps_loss).  Returns (loss, output) -- loss as a device tensor WITHOUT an autograd graph: its
    gradient w.r.t. every quantizer alpha except conv_out's own is exactly zero (torch.round, utils/quant_util.py:271),
    see include/attndm_b200.h::attndm_alpha_entropy_grad."""
    tl = t.long()
    if not bool((tl == tl[0]).all()):
        raise NotImplementedError("noise_estimation_loss: one timestep per call (functions/denoising.py:70-74)")
    bc = b.detach().float().cpu()
    a = (1 - bc).cumprod(dim=0)[int(tl[0])]                               # (:51) -- no +1 shift, unlike compute_alpha
    coef = torch.stack([a.sqrt(), (1.0 - a).sqrt()]).to(x0.device)
    with torch.no_grad():
        x = ops.noise_mix(ops.to_nhwc(x0).contiguous(), ops.to_nhwc(e).contiguous(), coef)
        output = model(ops.to_nchw(x), t.float())
        per = ops.sq_err(ops.to_nhwc(e).contiguous(), ops.to_nhwc(output).contiguous())   # [B] double
    if keepdim:
        return per.float(), output
    return per.mean(dim=0).float(), output


def _entropy_layers(model, attention_focus):
    from .quant_util import QConv2d
    from .self_attention import EnhancedQSelfAttention
    if attention_focus:
        out = []
        for layer in model.modules():
            if isinstance(layer, EnhancedQSelfAttention):
                out += [s for s in layer.modules() if type(s) is QConv2d]
        return out
    return [m for m in model.modules() if type(m) is QConv2d]


def generalized_steps_loss(x, seq, model, b, optimizer, t_mode, **kwargs):
    """:62-116: the DDIM trajectory with one optimiser step on the quantizer alphas per timestep.

    Per step the reference (a) evaluates the model on the forward-noised x_t and forms the noise-estimation loss,
    (b) adds diff_loss_weight * the entropy of softmax(alpha_activ)[step] of every (attention) QConv2d, (c) makes the
    DDIM update with the SAME model output, (d) back-propagates and steps the optimiser.  The loss of (a) reaches no
    alpha but conv_out's own (zero-derivative rounding in every quantizer), so the gradient written to `.grad` here is
    the closed-form entropy gradient (attndm_alpha_entropy_grad); the optimiser is the caller's (the reference passes
    torch.optim.AdamW, runners/diffusion.py:289).  kwargs: eta, args (diff_loss_weight), attention_focus,
    noise_fn(step, which in {'e', 'z'}, like) for tests, losses (a list that receives the per-step loss values)."""
    args = kwargs.get("args", None)
    attention_focus = kwargs.get("attention_focus", False)
    noise_fn = kwargs.get("noise_fn", None)
    loss_log = kwargs.get("losses", None)
    eta = kwargs.get("eta", 0)
    if not x.is_cuda:
        raise RuntimeError("attentiondm_b200.generalized_steps_loss: x must be a CUDA tensor (no CPU fallback)")
    model.eval()
    seq = list(seq)
    n = x.size(0)
    dev = x.device
    coef = ddim_coefficients(seq, b, eta).to(dev)
    layers = _entropy_layers(model, attention_focus)
    opt_params = {id(p) for grp in optimizer.param_groups for p in grp["params"]}
    co = getattr(model, "conv_out", None)
    if co is not None and hasattr(co, "alpha_activ") and id(co.alpha_activ) in opt_params:
        import warnings
        warnings.warn("generalized_steps_loss: conv_out.alpha_activ is being optimised; its (tiny) noise-estimation-loss "
                      "gradient is not produced on this path -- only the entropy regulariser is", RuntimeWarning)
    weight = float(args.diff_loss_weight) if args is not None else 0.0
    xs, x0_preds = [x], []
    xt = ops.to_nhwc(x).contiguous().clone()
    for count_1 in range(len(seq)):
        i = seq[len(seq) - 1 - count_1]
        t = torch.full((n,), float(i), device=dev)
        like = ops.to_nchw(xt)
        e = noise_fn(count_1, "e", like) if noise_fn is not None else torch.randn_like(like)
        total_loss, et = noise_estimation_loss(model, like, t, e, b)
        if loss_log is not None:
            loss_log.append(total_loss)
        z = noise_fn(count_1, "z", like) if noise_fn is not None else torch.randn_like(like)
        with torch.no_grad():
            xt_next, x0_t = ops.ddim_step(xt, ops.to_nhwc(et).contiguous(), coef[count_1],
                                          ops.to_nhwc(z).contiguous() if eta != 0 else None, want_x0=True)
        x0_preds.append(_to_host_async(x0_t))
        xs.append(_to_host_async(xt_next))
        # ---- optimizer.zero_grad(); total_loss.backward(); optimizer.step() (:111-113) ----
        optimizer.zero_grad()
        if args is not None:
            for q in layers:
                p = q.alpha_activ
                if not p.requires_grad:
                    continue
                g = torch.zeros_like(p.data)
                if weight != 0.0:
                    ops.alpha_entropy_grad(p.data[count_1].contiguous(), weight, g[count_1])
                p.grad = g
        optimizer.step()
        xt = xt_next
    torch.cuda.current_stream().synchronize()
    xs = [xs[0]] + [ops.to_nchw(h) for h in xs[1:]]
    return xs, [ops.to_nchw(h) for h in x0_preds]


def ddpm_coefficients(seq, b):
    """[T,8] fp32 table (CPU), row k = step k of the reversed sequence, in the reference's fp32 op order (:126-148):
    {(1/at).sqrt(), (1/at - 1).sqrt(), atm1.sqrt()*beta_t, (1-beta_t).sqrt()*(1-atm1), 1-at, mask*exp(0.5*log(beta_t)), t, 0}."""
    seq = list(seq)
    seq_next = [-1] + seq[:-1]
    bc = b.detach().float().cpu()
    rows = []
    for i, j in zip(reversed(seq), reversed(seq_next)):
        at = compute_alpha(bc, torch.tensor([i]).long()).view(())
        atm1 = compute_alpha(bc, torch.tensor([j]).long()).view(())
        beta_t = 1 - at / atm1
        mask = 1 - float(i == 0)
        logvar = beta_t.log()
        rows.append(torch.stack([(1.0 / at).sqrt(), (1.0 / at - 1).sqrt(), atm1.sqrt() * beta_t,
                                 (1 - beta_t).sqrt() * (1 - atm1), 1.0 - at, mask * torch.exp(0.5 * logvar),
                                 _scalar_f32(float(i)), _scalar_f32(0.0)]))
    return torch.stack(rows).float()


def ddpm_steps(x, seq, model, b, **kwargs):
    """:119-151: ancestral sampling (the ablation driver's loop).  kwargs: noise_fn(step, like) for tests."""
    noise_fn = kwargs.get("noise_fn", None)
    if not x.is_cuda:
        raise RuntimeError("attentiondm_b200.ddpm_steps: x must be a CUDA tensor (no CPU fallback)")
    with torch.no_grad():
        seq = list(seq)
        coef = ddpm_coefficients(seq, b).to(x.device)
        n = x.size(0)
        xt = ops.to_nhwc(x).contiguous().clone()
        host_x, host_x0 = [], []
        for k in range(len(seq)):
            t = torch.full((n,), float(seq[len(seq) - 1 - k]), device=x.device)
            e = ops.to_nhwc(model(ops.to_nchw(xt), t)).contiguous()
            like = ops.to_nchw(xt)
            noise = noise_fn(k, like) if noise_fn is not None else torch.randn_like(like)
            xt, x0 = ops.ddpm_step(xt, e, coef[k], ops.to_nhwc(noise).contiguous())
            host_x.append(_to_host_async(xt))
            host_x0.append(_to_host_async(x0))
        torch.cuda.current_stream().synchronize()
        return [x] + [ops.to_nchw(h) for h in host_x], [ops.to_nchw(h) for h in host_x0]
