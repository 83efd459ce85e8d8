"""Drives the UNMODIFIED reference (staged under baseline/_ref by vendor_reference.py) for bench.py's reference arm and
its `gpu_torch_baseline` leg.  Benchmark infrastructure only: nothing in the product imports this.

The metric is the boosted-mixture log-density + boosting weights of one batch.  The reference computes exactly that in the
first half of `compute_kl_pq_loss` (density_experiment.py:606-641) and in `evaluate` (:561-571); neither function can be
called for that quantity alone (the former goes on to resample and run the new component with autograd, the latter adds a
separate "c" pass and returns batch means), so `density_step` below replays those driver lines -- the loop over
`model(x=x, components=c)`, `log_normal_standard`, the two-term `torch.logsumexp` recursion, `utils.utilities.softmax`, the
clamp and the renormalisation -- calling the reference's own modules and helpers for every operation.
"""
import argparse
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF_DIR, "models", "boosted_flow.py"))


def load():
    """Import the staged reference (logging / IO-only third-party modules it imports at module scope are stubbed)."""
    import torch  # noqa: F401
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    sys.dont_write_bytecode = True
    if "tensorboardX" not in sys.modules:
        sys.modules["tensorboardX"] = types.SimpleNamespace(SummaryWriter=object)
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from models.boosted_flow import BoostedFlow
        from utils.distributions import log_normal_standard
        from utils.utilities import softmax
    return types.SimpleNamespace(BoostedFlow=BoostedFlow, log_normal_standard=log_normal_standard, softmax=softmax)


def make_args(cfg, device):
    """The reference's args Namespace for one BASELINE configuration."""
    import torch
    device = torch.device(device)
    return argparse.Namespace(
        flow="boosted", boosted=True, density_evaluation=True, device=device, cuda=device.type == "cuda",
        component_type=cfg["kind"], num_components=cfg["C"], num_flows=cfg["K"], z_size=cfg["D"], input_size=[cfg["D"]],
        h_size=cfg["h"], rho_init=cfg["rho"], coupling_network="tanh", coupling_network_depth=1, batch_norm=cfg["bn"],
        flow_permutation="shuffle", flow_coupling="affine", actnorm_scale=1.0, LU_decomposed=True, num_blocks=1,
        num_dequant_blocks=0, learn_top=False, y_classes=1, y_condition=False, sample_size=16, save_results=False,
        batch_size=100, rho_iters=0)


def build_model(ref, cfg, device, x_init):
    """Reference BoostedFlow under torch.manual_seed(1) (the driver's default --manual_seed), ActNorm initialised by one
    train-mode forward per component (density_experiment.py:346-356), then eval mode with all C components fixed."""
    import torch
    torch.manual_seed(1)
    args = make_args(cfg, device)
    model = ref.BoostedFlow(args).to(args.device)
    model.train()
    with torch.no_grad():
        for c in range(cfg["C"]):
            model(x=x_init, components=c)
    model.eval()
    model.component, model.all_trained = cfg["C"] - 1, False
    return model, args


def density_step(ref, model, x, args, toy=False):
    """G_ll over all C components (density_experiment.py:561-571) and the boosting weights (:627-641); toy configuration:
    base density model.base_dist and the toy clamp (toy_experiment.py:424-432,440,453-459)."""
    import torch
    with torch.no_grad():
        G_ll = torch.zeros(x.size(0))
        for c in range(model.component + 1):
            z_G, _, _, ldj_G, _ = model(x=x, components=c)
            if toy:
                ll = model.base_dist.log_prob(z_G).sum(1) + ldj_G
            else:
                ll = ref.log_normal_standard(z_G, reduce=True, dim=-1, device=args.device) + ldj_G
            if c == 0:
                G_ll = ll
            else:
                rho_simplex = model.rho[0:(c + 1)] / torch.sum(model.rho[0:(c + 1)])
                last_ll = torch.log(1 - rho_simplex[c]) + G_ll
                next_ll = torch.log(rho_simplex[c]) + ll
                uG_ll = torch.cat([last_ll.view(x.size(0), 1), next_ll.view(x.size(0), 1)], dim=1)
                G_ll = torch.logsumexp(uG_ll, dim=1)
        G_nll = -1.0 * G_ll
        weights = ref.softmax(G_nll)
        weights = torch.pow(weights, 1.0)          # heuristic == "unity", density_experiment.py:628-636
        if toy:
            weights = weights / torch.sum(weights)
            if weights.max() > 0.1:
                weights = torch.max(torch.min(weights, torch.tensor([0.1], device=x.device)),
                                    torch.tensor([0.1 / args.batch_size], device=x.device))
                weights = weights / torch.sum(weights)
        else:
            if weights.max() > 0.1:
                weights = torch.max(torch.min(weights, torch.tensor([0.1], device=x.device)),
                                    torch.tensor([0.01], device=x.device))
            if weights.sum() != 1.0:
                weights = weights / torch.sum(weights)
    return G_ll, weights
