"""Public surface of the package."""
from . import _lib
from ._lib import GbnfError
from .boosted_flow import BoostedFlow
from .extract import extract_model
from .flow_modules import ActNorm1d, BatchNorm, CouplingMLP, Glow, GlowStep, Permute1d, RealNVPFlow, ResidualBlock, ResidualNet
from .losses import compute_kl_pq_loss, evaluate, toy_compute_kl_pq_loss
from .build import build

__all__ = ["BoostedFlow", "Glow", "RealNVPFlow", "GlowStep", "CouplingMLP", "ResidualNet", "ResidualBlock", "ActNorm1d", "BatchNorm", "Permute1d", "GbnfError",
           "compute_kl_pq_loss", "evaluate", "toy_compute_kl_pq_loss", "extract_model", "build", "_lib"]
