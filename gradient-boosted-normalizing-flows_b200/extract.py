"""BoostedFlow module tree -> plain "model dict" of numpy arrays (the container the CPU oracle and the golden
fixtures use; format documented in oracle/gbnf_oracle.py).  Works on this package's modules AND on the reference's
(same attribute layout), which is how tests/golden/make_golden.py feeds reference models to the oracle."""
import numpy as np
import torch


def _np(t):
    return t.detach().cpu().numpy().copy()


def _linears(net):
    if hasattr(net, "initial_layer"):      # ResidualNet (models/layers.py:277-301): initial, (block first, block second)*, final
        mods = [net.initial_layer] + [lin for b in net.blocks for lin in b.linear_layers] + [net.final_layer]
        return [(_np(m.weight), _np(m.bias)) for m in mods]
    return [(_np(m.weight), _np(m.bias)) for m in net.network if isinstance(m, torch.nn.Linear)]


def extract_model(model, toy_base=False):
    args = model.args
    kind = args.component_type
    md = {"kind": kind, "D": int(args.z_size), "h": int(args.h_size), "K": int(args.num_flows),
          "C": int(args.num_components), "depth": int(args.coupling_network_depth),
          "act": args.coupling_network, "coupling": getattr(args, "flow_coupling", "affine") if kind == "glow" else "affine",
          "rho": _np(model.rho).astype(np.float32), "base_mean": None, "base_scale": None, "components": []}
    if toy_base:
        md["base_mean"], md["base_scale"] = _np(model.base_dist_mean), _np(model.base_dist_var)
    for c, flow in enumerate(model.flows):
        steps = []
        if kind == "glow":
            for st in flow.flow.layers:
                d = {"an_bias": _np(st.actnorm.bias).reshape(-1), "an_logs": _np(st.actnorm.logs).reshape(-1), "net": _linears(st.block)}
                if hasattr(st, "invconv"):        # InvertibleConv1x1 factors as the module holds them (models/layers.py:722-749)
                    ic = st.invconv
                    d["perm"] = None
                    d["ic"] = ({"p": _np(ic.p), "sign_s": _np(ic.sign_s), "lower": _np(ic.lower), "log_s": _np(ic.log_s), "upper": _np(ic.upper)}
                               if ic.LU_decomposed else {"weight": _np(ic.weight)})
                else:
                    perm = st.shuffle if hasattr(st, "shuffle") else st.reverse
                    d["perm"] = _np(perm.indices).astype(np.int64)
                steps.append(d)
        else:
            for (t_net, s_net, bn) in flow.flow_param:
                b = None
                if bn is not None:
                    b = {"log_gamma": _np(bn.log_gamma), "beta": _np(bn.beta), "mean": _np(bn.running_mean),
                         "var": _np(bn.running_var)}
                steps.append({"bn": b, "t": _linears(t_net), "s": _linears(s_net)})
        md["components"].append({"flip_init": int(getattr(flow, "flip_init", 0)), "steps": steps})
    return md
