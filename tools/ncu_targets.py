"""Small driver for `ncu --set full` captures of the kernels that changed in round 2 (one launch of each)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gcn_vae_b200 as K
from gcn_vae_b200 import ops, _lib as L

dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
# 1. one FB15k-237 full-graph train step (DistMult pass, fused KL backward, in-place epilogue, message passing)
data = K.datasets.synthetic_kg("FB15k-237", seed=0)
torch.manual_seed(0)
model = K.LinkPredict(K.KGVAE, data.num_nodes, 500, data.num_rels, num_bases=100, dropout=0.2, use_cuda=True,
                      reg_param=0.01, kl_param=1e-5, k=10, n_flows=0).to(dev)
np.random.seed(0)
g, node_id, etype, node_norm, samples, labels = K.utils.generate_sampled_graph_and_labels(
    data.train, len(data.train), 0.5, data.num_rels, None, None, 10, "uniform")
norm = K.node_norm_to_edge_norm(g, torch.from_numpy(node_norm).view(-1, 1)).to(dev)
for _ in range(2):
    model.zero_grad(set_to_none=True)
    z = model(g, torch.from_numpy(node_id).view(-1, 1).to(dev), torch.from_numpy(etype).to(dev), norm)
    loss, _, _, _ = model.get_loss(g, z, torch.from_numpy(samples).to(dev), torch.from_numpy(labels).to(dev))
    loss.backward()
# 2. top-k generation over the test queries
test = torch.from_numpy(data.test).to(dev)
K.utils.generate(z.detach(), model.w_relation, test, topk=10)
# 3. column-chunk message passing at a streaming-size graph (scaled wikikg2: 625 k nodes, 8 M edges)
n, E, R, B, si, so = 625_000, 8_000_000, 1070, 100, 5, 10
src = torch.randint(0, n, (E,), device=dev, generator=gen, dtype=torch.int32)
dst = torch.randint(0, n, (E,), device=dev, generator=gen, dtype=torch.int32)
et = torch.randint(0, R, (E,), device=dev, generator=gen, dtype=torch.int32)
nm = torch.rand(E, device=dev, generator=gen)
gi = ops.graph_index(src, dst, et, nm, n, R)
x = torch.randn(n, B * si, device=dev, generator=gen)
w = torch.randn(R, B * si * so, device=dev, generator=gen) * 0.05
agg = torch.zeros(n, B * so, device=dev)
pack = ops._rel_order(gi, 0, n, 4 * B * so)
for b0, b1 in ops._block_chunks(B, si, so, 2):
    xc = x[:, b0 * si:b1 * si].contiguous()
    L.call("kg_bdd_rel_fwd_cols", L.f32(xc), L.i32(pack), E, L.f32(w), b0, b1 - b0, B, si, so, L.f32(agg),
           ops.HINT_STREAM_X | ops.HINT_TILE_RESIDENT, L.stream())
# 4. L2 probe (the roofline denominator)
a = torch.randn(14541, 500, device=dev); b = torch.zeros(14541, 500, device=dev); sink = torch.zeros(4, device=dev)
for mode in (1, 2, 3):
    L.call("kg_probe_l2", L.f32(a), L.f32(b), 14541, 500, 3_000_000, mode, L.f32(sink), L.stream())
torch.cuda.synchronize()
print("done")
