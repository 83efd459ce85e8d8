"""Host mirror of the reference's grid-density caller of the hot path (utils/density_plotting.py:185-232).

The numbers come from the C-ABI library (per-component log q on the grid in one launch, geometric mixture in a second);
drawing is optional and only happens when Matplotlib axes are passed in, exactly where the reference draws."""
import numpy as np
import torch


def setup_grid(range_lim, n_pts, device="cpu"):
    """utils/density_plotting.py:115-119 (meshgrid in 'ij' order, as torch.meshgrid defaulted to when it was written)."""
    x = torch.linspace(-range_lim, range_lim, n_pts)
    xx, yy = torch.meshgrid((x, x), indexing="ij")
    zz = torch.stack((xx.flatten(), yy.flatten()), dim=1)
    return xx, yy, zz.to(device)


@torch.no_grad()
def boosted_fwd_flow_density(model, test_grid, n_pts, batch_size, args, axs=None, cmap=None):
    """plot_boosted_fwd_flow_density (utils/density_plotting.py:185-226): per-component densities exp(log q_c) on the grid and
    the geometric mixture exp(sum_c rho_c log q_c / sum_c rho_c).  Returns (total_prob [n_pts, n_pts], {c: prob_c}); the
    reference returns total_prob only.  `batch_size` is accepted for signature parity: the grid is evaluated in one launch."""
    xx, yy, zz = test_grid
    n = max(1, args.num_components if model.all_trained else model.component + 1)
    logq = model.component_log_density(zz, 0, n)                      # [n_pts^2, n], models/boosted_flow.py:220-228 per component
    total_log = model.mixture_from_logq(logq, n, geometric=True)
    plt_width = max(2, int(np.ceil(np.sqrt(args.num_components))))
    probs = {}
    rho = model.rho.detach().cpu()
    for c in range(n):
        if float(rho[c]) == 0.0:                                      # :200-201
            continue
        probs[c] = logq[:, c].exp().cpu().view(n_pts, n_pts)
        if axs is not None:
            ax = axs[int(1 + np.floor(c / plt_width)), int(c % plt_width)]
            ax.pcolormesh(xx, yy, probs[c], cmap=cmap)
            ax.set_title(f"c={c}", fontdict={"fontsize": 20})
    total_prob = total_log.exp().cpu().view(n_pts, n_pts)
    if axs is not None:
        axs[0, 1].pcolormesh(xx, yy, total_prob, cmap=cmap)
        axs[0, 1].set_title("GBF - All Components", fontdict={"fontsize": 20})
    return total_prob, probs
