"""The boosted objective and the full-mixture evaluation, served by the CUDA library.

Mirrors the two driver functions of the reference that ARE the hot path (SURVEY correction note):
    compute_kl_pq_loss(model, x, args)                    density_experiment.py:606-674
    evaluate(model, data_loader, args, results_type=None)  density_experiment.py:544-603
    toy_compute_kl_pq_loss(model, data_or_sampler, beta, args)   toy_experiment.py:397-503

Same arguments, same returned dicts, same exceptions.  Differences, all deliberate and documented in DESIGN.md:
  * the fixed mixture G^(c-1), its boosting weights and the resampling run as fused CUDA kernels under no_grad
    (the reference builds and discards autograd graphs for them);
  * resampling draws float64 uniforms from `generator` and applies the inverse-CDF contract that reproduces
    torch.multinomial(w, B, replacement=True) on CPU;
  * fixed RealNVP components always use eval-mode BatchNorm statistics.
"""
import torch

from .flow_modules import LOG_2PI


def _std_normal_logprob(z):
    return (-0.5 * LOG_2PI - 0.5 * z.pow(2)).sum(dim=-1)      # utils/distributions.py:44-60


def _new_component_nll(model, x, base="std"):
    """g_nll of the component being trained, with autograd when it has trainable parameters
    (density_experiment.py:647-649)."""
    z, _, _, ldj, _ = model(x=x, components="c")
    if base == "std":
        return -1.0 * (_std_normal_logprob(z) + ldj)
    return -1.0 * (model.base_dist.log_prob(z).sum(1) + ldj)


def _uniforms(n, device, generator):
    if generator is not None and generator.device.type != "cpu":
        return torch.rand(n, dtype=torch.float64, device=device, generator=generator)
    return torch.rand(n, dtype=torch.float64, generator=generator).to(device)


def compute_kl_pq_loss(model, x, args, generator=None, return_aux=False):
    if args.flow != "boosted":
        raise NotImplementedError("only the boosted density path is accelerated (density_experiment.py:606-661)")
    aux = {}
    if model.all_trained or model.component > 0:
        # 1. fixed mixture over the first `component` components            density_experiment.py:612-624
        G_ll = model.mixture_log_density(x, model.component)
        # 2. boosting weights + resampling                                   density_experiment.py:627-644
        weights = model.boosting_weights(G_ll, mode="density")
        u = _uniforms(x.size(0), x.device, generator)
        idx = model.resample(weights, u)
        x_resampled = model.gather_rows(x, idx)
        # 3. new component on the resampled batch                            density_experiment.py:647-653
        g_nll = _new_component_nll(model, x_resampled)
        losses = {"nll": torch.mean(g_nll), "G_nll": torch.mean(-G_ll), "g_nll": torch.mean(g_nll)}
        aux = {"G_ll": G_ll, "weights": weights, "idx": idx}
    else:
        g_nll = _new_component_nll(model, x)
        losses = {"nll": torch.mean(g_nll), "g_nll": torch.mean(g_nll)}
        losses["G_nll"] = torch.zeros_like(losses["g_nll"])
    if torch.isnan(losses["nll"]).any():
        model.check_status()      # a value that fp16 could not hold raises GbnfError(GBNF_ERR_NUMERIC) with the cause
        raise ValueError(f"Nan Encountered. nll={losses['nll']}, x={x}, losses={losses}")
    return (losses, aux) if return_aux else losses


@torch.no_grad()
def evaluate(model, data_loader, args, results_type=None):
    model.eval()
    if not args.boosted:
        raise NotImplementedError("only the boosted density path is accelerated (density_experiment.py:547-591)")
    G_nll, g_nll = [], []
    track_new = model.component > 0 or model.all_trained
    cnew = min(model.component, model.num_components - 1)
    for (x, _) in data_loader:
        x = x.to(args.device)
        n = model.component + 1
        if track_new and cnew < n:
            G_ll, logq = model.mixture_log_density(x, n, return_logq=True)   # one pass serves both quantities
            g_nll.append(-logq[:, cnew])
        else:
            G_ll = model.mixture_log_density(x, n)
            if track_new:
                g_nll.append(-model.component_log_density(x, cnew, cnew + 1)[:, 0])
        G_nll.append(-G_ll)
    G_nll = torch.cat(G_nll, dim=0)
    mean_G = G_nll.mean().item()
    losses = {"nll": mean_G}
    if track_new:
        g_nll = torch.cat(g_nll, dim=0)
        losses["g_nll"] = g_nll.mean().item()
        losses["ratio"] = torch.mean(g_nll - G_nll).item()
    else:
        losses["g_nll"] = mean_G
        losses["ratio"] = 0.0
    if getattr(args, "save_results", False) and results_type is not None:
        with open(args.exp_log, "a") as ff:
            print(f'{results_type} set loss: {losses["nll"]:.6f}', file=ff)
    return losses


def toy_compute_kl_pq_loss(model, data_or_sampler, beta, args, generator=None, return_aux=False):
    """toy_experiment.py:397-503 (boosted branch; base density model.base_dist => build the model with
    args.toy_base=True).  The 10 %-probability debug dump to counts.txt (:464-472) is not reproduced."""
    x = data_or_sampler(args.batch_size).to(args.device) if callable(data_or_sampler) else data_or_sampler
    if not args.boosted:
        raise NotImplementedError("only the boosted path is accelerated")
    aux = {}
    if model.component > 0:
        n = args.num_components if model.all_trained else model.component
        skip = model.component if model.all_trained else -1
        G_ll = model.mixture_log_density(x, n, skip_c=skip)
        weights = model.boosting_weights(G_ll, mode="toy", batch_size=args.batch_size)
        u = _uniforms(x.size(0), x.device, generator)
        idx = model.resample(weights, u)
        x_resampled = model.gather_rows(x, idx)
        g_nll = _new_component_nll(model, x_resampled, base="diag")
        losses = {"nll": torch.mean(g_nll), "G_nll": torch.mean(-G_ll), "g_nll": torch.mean(g_nll)}
        aux = {"G_ll": G_ll, "weights": weights, "idx": idx}
    else:
        g_nll = _new_component_nll(model, x, base="diag")
        losses = {"nll": torch.mean(g_nll), "g_nll": torch.mean(g_nll)}
        losses["G_nll"] = torch.zeros_like(losses["g_nll"])
    return (losses, aux) if return_aux else losses
