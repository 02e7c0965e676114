"""Drop-in for the evaluation caller of the hot path: ablation_study_attention_quantization.py
(AttentionQuantizationAblation, :35-372) -- four bit-width variants of the quantized UNet

    A  uniform 4-bit            B  4-bit convs, 8-bit attention projections
    C  8-bit convs, 4-bit attention        D  uniform 8-bit

built with the same `Model` / `EnhancedQSelfAttention` surface and mutated through the `w_bit` / `a_bit` setters exactly
as the reference does (:159-207), calibrated, and sampled with functions/denoising.py's ddpm_steps / generalized_steps.

Forced deviations (the shipped driver cannot run): its mock `args` have no `timesteps`
(utils/quant_util.py:228 reads it), its "calibration" runs ten forwards with every QConv2d in INFERENCE mode (so
groups_range stays zero), and its inline DDPM loop uses prod(alpha) of a scalar for alpha-bar.  Here `args.timesteps`
is set, calibration is a calibration pass along the sampling trajectory, and sampling goes through the reference's own
samplers.  FID / CLIP need Inception / CLIP weights that cannot be fetched here; `compute_fid` reports NaN like the
reference does when pytorch-fid is missing (:386-388), and `evaluate` adds metrics that need no external network:
eps and final-image deviation from the FP model on the same latents.
"""
import argparse
import os

import torch

from .denoising import ddpm_steps, generalized_steps
from .diffusion import Model
from .runner import get_beta_schedule, inverse_data_transform
from .self_attention import EnhancedQSelfAttention

VARIANTS = {                      # name: (conv bits, attention projection bits)   (:120-155)
    "A": (4, 4),
    "B": (4, 8),
    "C": (8, 4),
    "D": (8, 8),
}


class AttentionQuantizationAblation:
    def __init__(self, config, device=None, timesteps=100, output_dir=None, logger=None):
        self.config = config
        self.device = torch.device("cuda") if device is None else torch.device(device)
        self.timesteps = int(timesteps)
        n = config.diffusion.num_diffusion_timesteps
        self.sequence = list(range(0, n, max(1, n // self.timesteps)))[:self.timesteps]
        betas = get_beta_schedule(beta_schedule=config.diffusion.beta_schedule, beta_start=config.diffusion.beta_start,
                                  beta_end=config.diffusion.beta_end, num_diffusion_timesteps=n)
        self.betas = torch.from_numpy(betas).float().to(self.device)
        self.output_dir = output_dir
        self.log = (logger.info if logger is not None else (lambda *a, **k: None))

    # ---- :111-157 ----
    def prepare_model_variants(self, state_dict=None):
        variants = {}
        for name, (conv_bits, attn_bits) in VARIANTS.items():
            args = argparse.Namespace(bitwidth=conv_bits, calibrate_attention=(conv_bits != attn_bits),
                                      timesteps=len(self.sequence))
            model = Model(self.config, quantization=True, sequence=self.sequence, args=args)
            self._set_attention_precision(model, attn_bits, attn_bits, attn_bits, attn_bits)
            model = model.to(self.device).eval()
            model.materialize_lazy_layers()
            if state_dict is not None:
                model.load_state_dict(state_dict, strict=False)            # (:228-230)
            model.snap_weights_()          # weights onto each layer's own w_bit grid (H1): the integer path's precondition
            variants[name] = model
            self.log(f"Variant {name}: {conv_bits}-bit convs, {attn_bits}-bit attention projections")
        return variants

    def build_fp_model(self, state_dict=None):
        args = argparse.Namespace(bitwidth=8, timesteps=len(self.sequence))
        fp = Model(self.config, quantization=False, sequence=self.sequence, args=args).to(self.device).eval()
        fp.materialize_lazy_layers()
        if state_dict is not None:
            fp.load_state_dict({k: v for k, v in state_dict.items()
                                if not (k.endswith(".groups_range") or k.endswith(".alpha_activ"))}, strict=False)
        return fp

    # ---- :159-207 ----
    def _set_attention_precision(self, model, query_bits, key_bits, value_bits, output_bits):
        for module in model.modules():
            if isinstance(module, EnhancedQSelfAttention):
                module.bit_config = {"query": query_bits, "key": key_bits, "value": value_bits, "output": output_bits}
                if module.quantization and hasattr(module, "query_conv"):
                    for conv, bits in ((module.query_conv, query_bits), (module.key_conv, key_bits),
                                       (module.value_conv, value_bits), (module.output_conv, output_bits)):
                        conv.w_bit = bits
                        conv.a_bit = bits

    def bit_widths(self, model):
        """{layer name: (w_bit, a_bit)} -- what the setters left behind."""
        return {n: (q.w_bit, q.a_bit) for n, q in model.qconvs()}

    # ---- :232-272 ----
    def calibrate_models(self, model_variants, x, first=False):
        """One calibration pass of every variant along the sampling trajectory of `x` (device latents)."""
        for name, model in model_variants.items():
            model.reset_index_seq()
            model.set_calibrate(True, first=first)
            try:
                generalized_steps(x, self.sequence, model, self.betas, eta=0.0, keep="last")
            finally:
                model.set_calibrate(False)
            model.reset_index_seq()
            self.log(f"Finished calibrating Variant {name}")

    # ---- :274-372 ----
    def generate_samples(self, model_variants, num_samples=100, batch_size=10, sampler="ddpm", latents=None, save=False):
        """{variant: images [num_samples, C, H, W] in [0, 1] (CPU)}.  `latents` (optional, [num_samples, C, H, W]) fixes the
        starting noise so that variants can be compared sample by sample."""
        c = self.config.data
        out = {}
        for name, model in model_variants.items():
            imgs = []
            for i in range(0, num_samples, batch_size):
                n = min(batch_size, num_samples - i)
                x = (latents[i:i + n].to(self.device) if latents is not None
                     else torch.randn(n, c.channels, c.image_size, c.image_size, device=self.device))
                model.reset_index_seq()
                if sampler == "ddpm":
                    xs, _ = ddpm_steps(x, self.sequence, model, self.betas)
                else:
                    xs, _ = generalized_steps(x, self.sequence, model, self.betas, eta=0.0, keep="last")
                imgs.append(inverse_data_transform(self.config, xs[-1]).cpu())
            out[name] = torch.cat(imgs)
            if save and self.output_dir:
                d = os.path.join(self.output_dir, f"variant_{name}")
                os.makedirs(d, exist_ok=True)
                torch.save(out[name], os.path.join(d, "samples.pt"))
        return out

    def compute_fid(self, sample_paths, real_images_path=None):
        """:374-407 with pytorch-fid unavailable (its Inception weights cannot be downloaded here)."""
        return {variant: float("nan") for variant in sample_paths}

    def evaluate(self, model_variants, fp_model, x, probe_steps=(0, -1)):
        """Metrics that need no external network: for each variant the relative L2 error of eps against the FP model on
        the same (x, t) at the probed sequence positions, and of the final DDIM image from the same latents."""
        res = {}
        with torch.no_grad():
            ref_img = generalized_steps(x, self.sequence, fp_model, self.betas, eta=0.0, keep="last")[0][-1]
            rseq = list(reversed(self.sequence))
            for name, model in model_variants.items():
                errs = []
                for k in probe_steps:
                    k = k % len(rseq)
                    t = torch.full((x.shape[0],), float(rseq[k]), device=self.device)
                    model.reset_index_seq(k)
                    e_q, e_f = model(x, t), fp_model(x, t)
                    errs.append(float((e_q - e_f).norm() / e_f.norm()))
                model.reset_index_seq()
                img = generalized_steps(x, self.sequence, model, self.betas, eta=0.0, keep="last")[0][-1]
                model.reset_index_seq()
                res[name] = dict(eps_rel_l2=errs, image_rel_l2=float((img - ref_img).norm() / ref_img.norm()),
                                 int8_layers=sum(1 for _, q in model.qconvs() if q.int8_ok_all_steps()),
                                 layers=len(model.qconvs()))
        return res

    def run_ablation(self, state_dict=None, num_calibration_samples=16, num_samples=16, batch_size=16, sampler="ddim"):
        c = self.config.data
        variants = self.prepare_model_variants(state_dict)
        fp = self.build_fp_model(state_dict if state_dict is not None else variants["D"].state_dict())
        g = torch.Generator(device="cpu").manual_seed(1234)
        xc = torch.randn(num_calibration_samples, c.channels, c.image_size, c.image_size, generator=g).to(self.device)
        self.calibrate_models(variants, xc)
        metrics = self.evaluate(variants, fp, xc)
        lat = torch.randn(num_samples, c.channels, c.image_size, c.image_size, generator=g)
        samples = self.generate_samples(variants, num_samples, batch_size, sampler=sampler, latents=lat, save=True)
        fid = self.compute_fid({k: None for k in samples})
        return dict(metrics=metrics, fid=fid, samples={k: tuple(v.shape) for k, v in samples.items()})


def main(argv=None):
    import json
    from types import SimpleNamespace as ns
    ap = argparse.ArgumentParser(description="bit-width ablation of the quantized UNet (random-init weights unless --ckpt)")
    ap.add_argument("--image_size", type=int, default=32)
    ap.add_argument("--ch_mult", default="1,2,2,2")
    ap.add_argument("--timesteps", type=int, default=20)
    ap.add_argument("--samples", type=int, default=16)
    ap.add_argument("--sampler", default="ddim", choices=["ddim", "ddpm"])
    ap.add_argument("--ckpt", default=None)
    ap.add_argument("--out", default=None)
    a = ap.parse_args(argv)
    cfg = ns(data=ns(channels=3, image_size=a.image_size, dataset="CIFAR10", rescaled=True, logit_transform=False),
             model=ns(ch=128, ch_mult=[int(v) for v in a.ch_mult.split(",")], num_res_blocks=2, dropout=0.1, var_type="fixedlarge"),
             diffusion=ns(beta_schedule="linear", beta_start=0.0001, beta_end=0.02, num_diffusion_timesteps=1000))
    sd = None
    if a.ckpt:
        sd = torch.load(a.ckpt, map_location="cpu")
        sd = sd[0] if isinstance(sd, (list, tuple)) else sd
    torch.manual_seed(0)
    res = AttentionQuantizationAblation(cfg, timesteps=a.timesteps, output_dir=a.out).run_ablation(sd, num_samples=a.samples, sampler=a.sampler)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
