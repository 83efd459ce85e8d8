"""Shared test helpers: golden fixture -> oracle model dict / gbnf_b200.BoostedFlow."""
import argparse

import numpy as np
import torch

from oracle import gbnf_oracle as orc


def golden_model(g):
    flat = {k[len("model."):]: v for k, v in g.items() if k.startswith("model.")}
    return orc.unflatten_model(flat)


def args_for(md, device="cpu", **kw):
    a = argparse.Namespace(
        flow="boosted", boosted=True, density_evaluation=True, device=torch.device(device), cuda=(device != "cpu"),
        component_type=md["kind"], num_components=md["C"], num_flows=md["K"], z_size=md["D"], input_size=[md["D"]],
        h_size=md["h"], rho_init="decreasing", coupling_network=md["act"], coupling_network_depth=md["depth"],
        batch_norm=any(st.get("bn") is not None for c in md["components"] for st in c["steps"]) if md["kind"] == "realnvp" else False,
        flow_permutation="shuffle", flow_coupling=md.get("coupling", "affine"), actnorm_scale=1.0, LU_decomposed=True,
        num_blocks=1, num_dequant_blocks=0, learn_top=False, y_classes=1, y_condition=False, sample_size=16,
        save_results=False, batch_size=64, rho_iters=0, toy_base=md.get("base_mean") is not None)
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def load_into(model, md):
    """Copy a model dict's arrays into a gbnf_b200.BoostedFlow (host modules)."""
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    with torch.no_grad():
        model.rho.copy_(t(md["rho"]))
        if md.get("base_mean") is not None:
            model.base_dist_mean.copy_(t(md["base_mean"]))
            model.base_dist_var.copy_(t(md["base_scale"]))
        for c, comp in enumerate(md["components"]):
            flow = model.flows[c]
            for k, st in enumerate(comp["steps"]):
                step = flow.steps()[k]
                if md["kind"] == "glow":
                    step.actnorm.bias.copy_(t(st["an_bias"]).view(1, -1))
                    step.actnorm.logs.copy_(t(st["an_logs"]).view(1, -1))
                    step.actnorm.inited = True
                    if st.get("ic") is not None:
                        for n_, v in st["ic"].items():
                            getattr(step.invconv, n_).copy_(t(v))
                    else:
                        step.permutation.set_indices(t(st["perm"]))
                    nets = [(step.block, st["net"])]
                else:
                    nets = [(step[0], st["t"]), (step[1], st["s"])]
                    if st.get("bn") is not None:
                        bn = step[2]
                        bn.log_gamma.copy_(t(st["bn"]["log_gamma"])); bn.beta.copy_(t(st["bn"]["beta"]))
                        bn.running_mean.copy_(t(st["bn"]["mean"])); bn.running_var.copy_(t(st["bn"]["var"]))
                for net, layers in nets:
                    for lin, (W, b) in zip(net.linears(), layers):
                        lin.weight.copy_(t(W)); lin.bias.copy_(t(b))
    return model


def build_model(md, device="cpu", gemm_mode="fp32", **kw):
    import gbnf_b200
    ic = md["components"][0]["steps"][0].get("ic") if md["kind"] == "glow" else None
    perm_kind = "invconv" if ic is not None else "shuffle"
    if ic is not None:
        kw.setdefault("LU_decomposed", "weight" not in ic)
    a = args_for(md, device, flow_permutation=perm_kind, **kw)
    torch.manual_seed(0)
    m = gbnf_b200.BoostedFlow(a, gemm_mode=gemm_mode)
    load_into(m, md)
    m = m.to(device)
    m.eval()
    return m


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30)))
