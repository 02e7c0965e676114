"""attentiondm_b200 -- B200-native hot path of PTQ-AttnDM (aqilmarwan/attentionDM).

Host-side mirror of the reference's operator surface; all compute goes through
the C-ABI in include/attndm_b200.h (libattndm_b200.so, hand-written sm_100a CUDA).
"""
from .quant_util import (FConv2d, GroupWise_Quantizaion, QConv2d, QModule, Quant, find_scale_by_percentile_max,
                         find_scale_by_percentile_min, lp_loss)
from .quantization_utils import AsymmetricQuantFunction, asymmetric_linear_quantization_params
from .attention_quant_utils import AttentionCalibrator, MixedPrecisionAttention
from .self_attention import EnhancedQSelfAttention, create_enhanced_attention
from .diffusion import DownBlock, Model, ResidualBlock, UpBlock, get_timestep_embedding
from .denoising import (cal_entropy, compute_alpha, ddpm_steps, generalized_steps, generalized_steps_loss,
                        noise_estimation_loss)
from .runner import Diffusion, get_beta_schedule

__all__ = [
    "QModule", "QConv2d", "Quant", "GroupWise_Quantizaion", "lp_loss", "find_scale_by_percentile_min",
    "find_scale_by_percentile_max", "AsymmetricQuantFunction", "asymmetric_linear_quantization_params",
    "MixedPrecisionAttention", "AttentionCalibrator", "EnhancedQSelfAttention", "create_enhanced_attention",
    "Model", "ResidualBlock", "DownBlock", "UpBlock", "get_timestep_embedding", "generalized_steps",
    "compute_alpha", "Diffusion", "get_beta_schedule", "FConv2d", "generalized_steps_loss", "noise_estimation_loss",
    "ddpm_steps", "cal_entropy",
]
