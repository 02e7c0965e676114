"""Drop-in for the sampling half of runners/diffusion.py: `Diffusion(args, config,
device).sample()` (:67-98, :308-459) and `get_beta_schedule` (:34-64).

Differences from the shipped reference, all forced by its defects (SURVEY.md
section 0.3): the model is calibrated before sampling (the reference never calls any
calibration, so groups_range stays zero and the output is NaN), weight ranges
are initialised (they are never assigned in the reference), and images are
returned (and saved only if args.image_folder is set).
"""
import os

import numpy as np
import torch

from .denoising import generalized_steps, generalized_steps_loss
from .diffusion import Model


def get_beta_schedule(beta_schedule, *, beta_start, beta_end, num_diffusion_timesteps):
    """runners/diffusion.py:34-64."""
    def sigmoid(x):
        return 1 / (np.exp(-x) + 1)

    if beta_schedule == "quad":
        betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, num_diffusion_timesteps, dtype=np.float64) ** 2
    elif beta_schedule == "linear":
        betas = np.linspace(beta_start, beta_end, num_diffusion_timesteps, dtype=np.float64)
    elif beta_schedule == "const":
        betas = beta_end * np.ones(num_diffusion_timesteps, dtype=np.float64)
    elif beta_schedule == "jsd":
        betas = 1.0 / np.linspace(num_diffusion_timesteps, 1, num_diffusion_timesteps, dtype=np.float64)
    elif beta_schedule == "sigmoid":
        betas = np.linspace(-6, 6, num_diffusion_timesteps)
        betas = sigmoid(betas) * (beta_end - beta_start) + beta_start
    else:
        raise NotImplementedError(beta_schedule)
    assert betas.shape == (num_diffusion_timesteps,)
    return betas


def inverse_data_transform(config, X):
    """datasets/__init__.py:206-215 (the only dataset helper sample() touches)."""
    if hasattr(config, "image_mean"):
        X = X + config.image_mean.to(X.device)[None, ...]
    if getattr(config.data, "logit_transform", False):
        X = torch.sigmoid(X)
    elif getattr(config.data, "rescaled", False):
        X = (X + 1.0) / 2.0
    return torch.clamp(X, 0.0, 1.0)


def make_seq(args, num_timesteps):
    """runners/diffusion.py:319-329."""
    if args.skip_type == "uniform":
        skip = num_timesteps // args.timesteps
        return range(0, num_timesteps, skip)
    if args.skip_type == "quad":
        seq = np.linspace(0, np.sqrt(num_timesteps * 0.8), args.timesteps) ** 2
        return [int(s) for s in list(seq)]
    raise NotImplementedError(args.skip_type)


def load_by_shape_match(model, states):
    """Positional shape-match checkpoint copy of runners/diffusion.py:376-400."""
    state_dict = model.state_dict()
    keys = list(states.keys())
    i = 0
    skip = ("activation_range_min", "activation_range_max", "x_min", "x_max", "groups_range", "alpha_activ",
            "mix_activ_mark1")
    for k, v in state_dict.items():
        if any(s in k for s in skip):
            continue
        if i < len(keys) and v.shape == states[keys[i]].shape:
            state_dict[k] = states[keys[i]]
            i += 1
    model.load_state_dict(state_dict, strict=False)
    return i


# ---------------------------------------------------------------------------------------------
# Checkpoint I/O for the quantizer state.  The reference can only persist what
# `state_dict()` holds ({weight, bias, groups_range, alpha_activ} per QConv2d); what calibration also produces --
# weight_range_min/max, init_range_min/max (first-calibrate search), the per-layer bit widths and the lazily
# created channel_proj convs -- lives in plain attributes and is lost on reload.  These two helpers keep it.
# ---------------------------------------------------------------------------------------------
QUANT_STATE_VERSION = 1


def quant_state(model):
    """Everything needed to resume sampling without re-calibrating: the state_dict plus the non-state_dict
    attributes of every QConv2d (utils/quant_util.py:88-118)."""
    layers = {}
    for name, q in model.qconvs():
        layers[name] = dict(
            weight_range_min=q.weight_range_min.detach().cpu().clone(),
            weight_range_max=q.weight_range_max.detach().cpu().clone(),
            init_range_min=q.init_range_min.detach().cpu().clone(),
            init_range_max=q.init_range_max.detach().cpu().clone(),
            a_bit=int(q._a_bit), w_bit=int(q._w_bit), group_num=int(q.group_num), index_seq=int(q.index_seq))
    return dict(version=QUANT_STATE_VERSION,
                state_dict={k: v.detach().cpu().clone() for k, v in model.state_dict().items()}, layers=layers)


def save_quant_state(model, path):
    torch.save(quant_state(model), path)


def load_quant_state(model, state):
    """Inverse of quant_state(); `state` may be a path.  The model must have been built from the same config."""
    if isinstance(state, (str, os.PathLike)):
        state = torch.load(state, map_location="cpu")
    if state.get("version") != QUANT_STATE_VERSION:
        raise RuntimeError(f"quant state version {state.get('version')} not understood")
    model.materialize_lazy_layers()
    mods = dict(model.qconvs())
    missing = set(state["layers"]) ^ set(mods)
    if missing:
        raise RuntimeError(f"quant state does not match the model's QConv2d layers: {sorted(missing)[:4]} ...")
    for name, d in state["layers"].items():
        q = mods[name]
        if d["group_num"] != q.group_num:
            raise RuntimeError(f"{name}: group_num {d['group_num']} != {q.group_num}")
    model.load_state_dict(state["state_dict"], strict=True)
    for name, d in state["layers"].items():
        q = mods[name]
        dev = q.weight.device
        q.weight_range_min = d["weight_range_min"].to(dev)
        q.weight_range_max = d["weight_range_max"].to(dev)
        q.init_range_min = d["init_range_min"].clone()
        q.init_range_max = d["init_range_max"].clone()
        q._a_bit, q._w_bit = d["a_bit"], d["w_bit"]
        q.index_seq = d["index_seq"]
        q.invalidate_cache()
    return model


class Diffusion(object):
    def __init__(self, args, config, device=None):
        self.args = args
        self.config = config
        if device is None:
            device = torch.device("cuda")
        self.device = device
        self.model_var_type = config.model.var_type
        betas = get_beta_schedule(
            beta_schedule=config.diffusion.beta_schedule, beta_start=config.diffusion.beta_start,
            beta_end=config.diffusion.beta_end, num_diffusion_timesteps=config.diffusion.num_diffusion_timesteps)
        self.betas = torch.from_numpy(betas).float().to(self.device)
        self.num_timesteps = self.betas.shape[0]
        self.seq = None
        self.model = None

    def build_model(self, states=None, snap_weights=True):
        self.seq = make_seq(self.args, self.num_timesteps)
        model = Model(self.config, quantization=True, sequence=self.seq, args=self.args).to(self.device)
        if states is not None:
            load_by_shape_match(model, states)
        if snap_weights:
            model.snap_weights_()
        else:
            model.init_weight_ranges()
        model.eval()
        self.model = model
        return model

    def build_fp_model(self, states=None):
        """The un-quantized twin (Model(quantization=False), models/diffusion.py:281-345) whose trajectories
        generate_calibrate_set samples; same checkpoint loader as the quantized model."""
        if self.seq is None:
            self.seq = make_seq(self.args, self.num_timesteps)
        fp = Model(self.config, quantization=False, sequence=self.seq, args=self.args).to(self.device)
        if states is not None:
            load_by_shape_match(fp, states)
        fp.eval()
        self.fpmodel = fp
        return fp

    # ---- active timestep selection (runners/diffusion.py:195-264) ----
    def cal_entropy(self, attn):
        """:195-196."""
        return -1 * torch.sum((attn * torch.log(attn)), dim=-1).mean()

    def timestep_uncertainty(self, model):
        """The `diff` branch's score (:232-241): sum over QConv2d of the entropy of softmax(alpha_activ, dim=1)[t] / C_in,
        minus sample_weight * sample_count; one [T] vector (tiny host-side math on [T,G,C] tables)."""
        Tn = self.args.timesteps
        unc = torch.zeros(Tn, device=self.device)
        for _, layer in model.qconvs():
            alpha = torch.softmax(layer.alpha_activ.detach().float(), dim=1)            # [T,G,C]
            dim = alpha.shape[2]
            ent = (-1 * torch.sum(alpha * torch.log(alpha), dim=-1)).mean(dim=-1)        # cal_entropy(alpha[t]) for every t
            unc += ent.to(self.device) / dim
        if not hasattr(self, "sample_count"):
            self.sample_count = torch.zeros(Tn)
        unc -= float(self.args.sample_weight) * self.sample_count.to(self.device)
        return unc

    def generate_calibrate_set(self, fpmodel, model, t_mode, num_calibrate_set, x=None, t_random=None):
        """:198-264.  Samples n <= 16 latents through the FP model and picks, per t_mode, which point of each
        trajectory calibration sees.  `x` / `t_random` (tests) replace the two random draws (:202-208, :225-227).
        Returns the calibration images on the device."""
        with torch.no_grad():
            n = min(num_calibrate_set, 16)
            if x is None:
                x = torch.randn(n, self.config.data.channels, self.config.data.image_size, self.config.data.image_size,
                                device=self.device)
            x = x.to(self.device).clone()
            xs = generalized_steps(x, self.seq, fpmodel, self.betas, eta=self.args.eta)[0]
            if t_mode == "real":
                x = xs[-1].to(self.device)
            elif t_mode == "range":
                for s_ in range(n):
                    x[s_] = (xs[-1][s_] if s_ >= 100 else xs[s_][s_]).to(self.device)
            elif t_mode == "random":
                if t_random is None:
                    normal_val = torch.nn.init.normal_(torch.Tensor(n), mean=0.4, std=0.4) * self.args.timesteps
                    t_random = normal_val.clone().type(torch.int).clamp(0, self.args.timesteps - 1)
                for s_ in range(n):
                    x[s_] = xs[int(t_random[s_])][s_].to(self.device)
            elif t_mode == "diff":
                unc = self.timestep_uncertainty(model)
                mark = torch.arange(0, self.args.timesteps, device=self.device)
                unc, mark = unc[30:], mark[30:]                       # (:242-243)
                t_sel = int(mark[unc == torch.max(unc)][-1])
                self.sample_count[t_sel] += 1
                x = xs[t_sel].to(self.device)
                self.timestep_select = t_sel
            else:
                raise NotImplementedError(t_mode)
            return inverse_data_transform(self.config, x)

    # ---- attention calibration with the entropy regulariser (runners/diffusion.py:266-306) ----
    def attention_qconvs(self, model):
        from .quant_util import QConv2d
        from .self_attention import EnhancedQSelfAttention
        out = []
        for _, module in model.named_modules():
            if isinstance(module, EnhancedQSelfAttention):
                out += [sub for _, sub in module.named_modules() if isinstance(sub, QConv2d)]
        return out

    def calibrate_attention(self, model, image, device=None, batchsize=None, **kw):
        """:266-306: the attention projections' QConv2d go to calibration mode, their alpha_activ is optimised with AdamW
        (lr 0.05, weight decay 0.05) along one generalized_steps_loss trajectory, then they return to inference mode."""
        device = self.device if device is None else device
        convs = self.attention_qconvs(model)
        for q in convs:
            q.set_calibrate(calibrate=True)
            q.first_calibrate(calibrate=getattr(self, "first_flag", False))
        image = image.to(device)
        attention_params = []
        for name, param in model.named_parameters():
            if "alpha_activ" in name and any(a in name for a in ["query_conv", "key_conv", "value_conv", "output_conv"]):
                param.requires_grad = True
                attention_params += [param]
        try:
            if attention_params:
                optimizer = torch.optim.AdamW(attention_params, 0.05, weight_decay=0.05)
                self.last_calibration = generalized_steps_loss(
                    image, self.seq, model, self.betas, optimizer, eta=getattr(self.args, "eta", 0.0),
                    t_mode=getattr(self, "t_mode", None), timestep_select=getattr(self, "timestep_select", None),
                    args=self.args, attention_focus=True, **kw)
        finally:
            for q in convs:
                q.set_calibrate(calibrate=False)
        return model

    def calibrate_general(self, model, data, device=None, batchsize=None, first=False):
        """Stage 1 of calibrate_model (:461-467); the reference calls it but never defines it.  One calibration pass of
        every QConv2d along the sample trajectory of `data`."""
        model.reset_index_seq()
        model.set_calibrate(True, first=first)
        try:
            generalized_steps(data.to(self.device), self.seq, model, self.betas, eta=getattr(self.args, "eta", 0.0), keep="last")
        finally:
            model.set_calibrate(False)
        model.reset_index_seq()
        return model

    def calibrate_pipeline(self, model, data, device=None):
        """calibrate_model of the reference (:461-478): general calibration, attention calibration with the entropy
        regulariser, and -- with args.mixed_precision_attention -- the attention-internal quantizers."""
        self.calibrate_general(model, data, device, getattr(self.args, "batchsize", None))
        self.calibrate_attention(model, data, device, getattr(self.args, "batchsize", None))
        model.reset_index_seq()
        if getattr(self.args, "mixed_precision_attention", False):
            self.calibrate_mixed_precision_attention(model, data, device)
        return model

    def calibrate_mixed_precision_attention(self, model, image, device=None):
        """:480-513."""
        from .attention_quant_utils import AttentionCalibrator
        mods = [m for _, m in model.named_modules()
                if getattr(m, "mixed_precision", False) and getattr(m, "quantization", False)]
        if not mods:
            return
        AttentionCalibrator(model, self.device if device is None else device).calibrate(image.to(self.device), [0, 250, 500, 750, 999])

    def calibrate_model(self, x, first=False):
        """One calibration pass over the sample trajectory (what calibrate_general was meant to
        drive, runners/diffusion.py:461-478): every QConv2d collects its per-step group ranges."""
        m = self.model
        m.reset_index_seq()
        m.set_calibrate(True, first=first)
        try:
            generalized_steps(x, self.seq, m, self.betas, eta=getattr(self.args, "eta", 0.0), keep="last")
        finally:
            m.set_calibrate(False)
        m.reset_index_seq()

    def sample(self, x=None, calibrate=True):
        """runners/diffusion.py:308-459."""
        if self.model is None:
            states = None
            ckpt = getattr(self.args, "ckpt_path", None)
            if ckpt:
                states = torch.load(ckpt, map_location="cpu")
                if isinstance(states, (list, tuple)):
                    states = states[-1] if self.config.data.dataset == "CELEBA" else states[0]
            self.build_model(states)
        n = getattr(self.args, "num_samples", 50)
        if x is None:
            x = torch.randn(n, self.config.data.channels, self.config.data.image_size,
                            self.config.data.image_size, device=self.device)
        if calibrate:
            self.calibrate_model(x)
        xs, _ = generalized_steps(x, self.seq, self.model, self.betas, eta=getattr(self.args, "eta", 0.0),
                                  keep="last")
        imgs = inverse_data_transform(self.config, xs[-1])
        folder = getattr(self.args, "image_folder", None)
        if folder:
            try:
                import torchvision.utils as tvu
                os.makedirs(folder, exist_ok=True)
                for i in range(imgs.shape[0]):
                    tvu.save_image(imgs[i], os.path.join(folder, f"sample_{i}.png"))
            except ImportError:
                torch.save(imgs, os.path.join(folder, "samples.pt"))
        return imgs
