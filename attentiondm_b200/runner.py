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

from .denoising import generalized_steps
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
