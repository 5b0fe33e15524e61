"""Diagnostic: repeat the FB15k-step-shaped encoder forward and compare stages with the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gcn_vae_b200 as K
from oracle import kgvae_oracle as O

DEV = "cuda:0"
n_ent, n_rel, h, bases, k = 14541, 237, 500, 100, 10
data = K.datasets.synthetic_kg("FB15k-237", seed=0, scale=0.05)
torch.manual_seed(0)
model = K.LinkPredict(K.KGVAE, n_ent, h, n_rel, num_bases=bases, dropout=0.2, reg_param=0.01, kl_param=1e-5, k=k, n_flows=0).to(DEV)
np.random.seed(0)
g, node_id, edge_type, node_norm, samples, labels = K.utils.generate_sampled_graph_and_labels(data.train, 2000, 0.5, n_rel, None, None, 10, "uniform")
n = len(node_id)
eps = torch.randn(n, h)
m1 = (torch.rand(n, h) < 0.8).float() / 0.8
m2 = (torch.rand(n, 2 * h) < 0.8).float() / 0.8
enc = model.encoder
enc.preset_eps, enc.rconv_layer_1.dropout_mask, enc.rconv_layer_2.dropout_mask = eps.to(DEV), m1.to(DEV), m2.to(DEV)
edge_norm = K.node_norm_to_edge_norm(g, torch.from_numpy(node_norm).view(-1, 1)).to(DEV)
ids = torch.from_numpy(node_id).view(-1, 1).to(DEV)
et = torch.from_numpy(edge_type).to(DEV)
params = {key: val.detach().cpu().clone() for key, val in model.state_dict().items() if not key.endswith(("mask", "pi"))}
graph = {"num_nodes": n, "src": g._src, "dst": g._dst, "etype": edge_type, "norm": node_norm,
         "edge_norm": node_norm[g._dst].reshape(-1, 1).astype(np.float32)}
with torch.no_grad():
    ref = O.kgvae_encode(params, graph, node_id, eps, bases, 0, (m1, m2))
print("oracle keys", list(ref.keys()))
rel = lambda a, b: float((a.detach().cpu().double() - b.double()).abs().max() / b.double().abs().max())
first = None
for it in range(8):
    with torch.no_grad():
        x0 = enc.input_layer(g, ids, et, edge_norm)
        h1 = enc.rconv_layer_1(g, x0, et, edge_norm)
        h2 = enc.rconv_layer_2(g, h1, et, edge_norm)
        z = model(g, ids, et, edge_norm)
    cur = (h1.cpu(), h2.cpu(), z.cpu())
    msg = f"it {it}: z vs oracle {rel(z, ref['z']):.3e}"
    for name, t in (("h1", h1), ("h2", h2)):
        if name in ref:
            msg += f"  {name} vs oracle {rel(t, ref[name]):.3e}"
    if first is not None:
        msg += "  vs first run: " + " ".join(f"{float((a - b).abs().max()):.2e}" for a, b in zip(cur, first))
    else:
        first = cur
    print(msg, flush=True)
