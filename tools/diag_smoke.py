"""Diagnostic: the smoke() configuration stage by stage against the oracle (which stage carries the error)."""
import sys, os
ROOT = os.environ.get("KG_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import gcn_vae_b200 as K
from oracle import kgvae_oracle as O

dev = torch.device("cuda:0")
torch.manual_seed(0)
n_ent, n_rel, h, bases, k, n_flows = 300, 6, 40, 8, 4, int(os.environ.get("FLOWS", "1"))
data = K.datasets.synthetic_kg("toy", seed=3)
data.train[:, [0, 2]] %= n_ent
data.train[:, 1] %= n_rel
model = K.LinkPredict(K.KGVAE, n_ent, h, n_rel, num_bases=bases, dropout=0.2, use_cuda=True,
                      reg_param=0.01, kl_param=1e-3, k=k, n_flows=n_flows).to(dev)
np.random.seed(0)
g, node_id, edge_type, node_norm, samples, labels = K.utils.generate_sampled_graph_and_labels(
    data.train, 600, 0.5, n_rel, None, None, 4, "uniform")
n = len(node_id)
eps = torch.randn(n, h)
m1 = (torch.rand(n, h) < 0.8).float() / 0.8
m2 = (torch.rand(n, 2 * h) < 0.8).float() / 0.8
enc = model.encoder
enc.preset_eps, enc.rconv_layer_1.dropout_mask, enc.rconv_layer_2.dropout_mask = eps.to(dev), m1.to(dev), m2.to(dev)
edge_norm = K.node_norm_to_edge_norm(g, torch.from_numpy(node_norm).view(-1, 1)).to(dev)
ids, et = torch.from_numpy(node_id).view(-1, 1).to(dev), torch.from_numpy(edge_type).to(dev)
params = {key: val.detach().cpu().clone() for key, val in model.state_dict().items() if not key.endswith(("mask", "pi"))}
graph = O.build_graph_from_triplets(n, n_rel, *[np.zeros(0, dtype=np.int64)] * 3)
graph.update(src=g._src, dst=g._dst, etype=edge_type, norm=node_norm, edge_norm=node_norm[g._dst].reshape(-1, 1))
with torch.no_grad():
    ref = O.kgvae_encode(params, graph, node_id, eps, bases, n_flows, (m1, m2))
    x0 = enc.input_layer(g, ids, et, edge_norm)
    h1 = enc.rconv_layer_1(g, x0, et, edge_norm)
    h2 = enc.rconv_layer_2(g, h1, et, edge_norm)
    z = model(g, ids, et, edge_norm)
rel = lambda a, b: float((a.detach().cpu().double() - b.double()).abs().max() / b.double().abs().max())
got = {"h1": h1, "h2": h2, "z_mean": enc.z_mean, "z_sigma": enc.z_sigma, "z": z}
print(os.path.basename(ROOT), "n =", n, " ".join(f"{k_}:{rel(v, ref[k_]):.2e}" for k_, v in got.items() if k_ in ref),
      "| oracle keys:", sorted(ref.keys()))
if "z0" in ref and n_flows > 0:
    with torch.no_grad():
        zm, zs, z0 = K.ops.ReparamFn.apply(h2, eps.to(dev))
        print("  z0:", f"{rel(z0, ref['z0']):.2e}")
        made = enc.nf[0]
        ws = [l.masked_weight() for l in made._linears()]
        bs = [l.bias for l in made._linears()]
        for rows in (n, 1):
            xin = torch.randn(rows, h, device=dev)
            cur, cur64 = xin, xin.double()
            for i, (w, b) in enumerate(zip(ws, bs)):
                relu = i + 1 < len(ws)
                cur = K.ops.LinearFn.apply(cur, w, b, relu)
                cur64 = cur64 @ w.double().t() + b.double()
                cur64 = torch.relu(cur64) if relu else cur64
                print(f"  rows={rows} linear {i} {tuple(w.shape)}: {rel(cur, cur64.cpu()):.2e}")
        xo, ld = made.forward(z0)
        print("  made.forward vs oracle z:", f"{rel(xo, ref['z']):.2e}", " log_det:", f"{rel(ld.reshape(-1), ref['log_det_sum'].reshape(-1)):.2e}" if 'log_det_sum' in ref else "")
with torch.no_grad():
    zc = (enc.z_mean + eps.to(dev) * torch.sqrt(enc.z_sigma)).cpu()
    zr = ref["z_mean"] + eps * torch.sqrt(ref["z_sigma"])
    z0_ref = ref.get("z0", ref["z"])
    d = (z.cpu() - ref["z"]).abs() if n_flows == 0 else (zc - z0_ref).abs()
    i = int(d.argmax())
    r, c = i // h, i % h
    print(f"  z(kernel) vs m+eps*sqrt(v) from GPU outputs: {rel(z if n_flows == 0 else zc, zc):.2e};  oracle z0 vs its own m+eps*sqrt(v): {rel(z0_ref, zr):.2e}")
    print(f"  worst element ({r},{c}): got {float(z.cpu()[r, c]) if n_flows == 0 else float(zc[r, c]):.8f} want {float(z0_ref[r, c]):.8f}  m {float(enc.z_mean[r, c]):.8f}/{float(ref['z_mean'][r, c]):.8f}"
          f"  v {float(enc.z_sigma[r, c]):.8e}/{float(ref['z_sigma'][r, c]):.8e} eps {float(eps[r, c]):.6f}")
    print(f"  elements with |diff| > 1e-5*max: {int((d > 1e-5 * float(z0_ref.abs().max())).sum())} of {d.numel()}")
with torch.no_grad():
    e_bad = int((enc.preset_eps.cpu() != eps).sum())
    m_bad = int((enc.rconv_layer_1.dropout_mask.cpu() != m1).sum()) + int((enc.rconv_layer_2.dropout_mask.cpu() != m2).sum())
    zk = z if n_flows == 0 else K.ops.ReparamFn.apply(h2, enc.preset_eps)[2]
    z_again = K.ops.ReparamFn.apply(h2, enc.preset_eps)[2]
    zerr = rel(z, ref["z"])
    print(f"CHECK zerr={zerr:.2e} device-eps-mismatches={e_bad} mask-mismatches={m_bad} reparam-repeat-maxdiff={float((zk - z_again).abs().max()):.2e} "
          f"{'EXCURSION' if zerr > 5e-6 else 'ok'}")
