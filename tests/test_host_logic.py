"""Host-side logic of the package against the oracle and the golden vectors (CPU only)."""
import numpy as np
import torch

from helpers import O

import gcn_vae_b200 as K


def _train_triples(seed, n_ent, n_rel, n):
    rng = np.random.default_rng(seed)
    s = rng.integers(0, n_ent, size=n)
    o = rng.integers(0, n_ent, size=n)
    r = rng.integers(0, n_rel, size=n)
    return np.stack([s, r, o], axis=1).astype(np.int64)


def test_sampler_matches_reference_golden(golden):
    gv = golden("sampling_seed0")
    n_ent, n_rel, n_train, batch, neg, seed, _ = (int(x) for x in gv["cfg"])
    train = _train_triples(seed, n_ent, n_rel, n_train)
    np.random.seed(0)
    g, uniq_v, rel, norm, samples, labels = K.utils.generate_sampled_graph_and_labels(
        train, batch, 0.5, n_rel, None, None, neg, "uniform")
    assert np.array_equal(g._src, gv["g_src"])
    assert np.array_equal(g._dst, gv["g_dst"])
    assert np.array_equal(rel, gv["edge_type"])
    assert np.array_equal(norm.astype(np.float32), gv["node_norm"])
    assert np.array_equal(uniq_v, gv["node_id"])
    assert np.array_equal(samples, gv["samples"])
    assert float(labels.sum()) == float(gv["labels_sum"])
    en = K.node_norm_to_edge_norm(g, torch.from_numpy(norm).view(-1, 1))
    assert np.array_equal(en.numpy().reshape(-1), norm[g._dst])


def test_graph_build_matches_oracle():
    t = _train_triples(5, 50, 4, 300)
    g, rel, norm = K.utils.build_graph_from_triplets(50, 4, (t[:, 0], t[:, 1], t[:, 2]))
    want = O.build_graph_from_triplets(50, 4, t[:, 0], t[:, 1], t[:, 2])
    assert np.array_equal(g._src, want["src"]) and np.array_equal(g._dst, want["dst"])
    assert np.array_equal(rel, want["etype"])
    assert np.array_equal(norm.astype(np.float32), want["norm"])
    assert len(g) == 50 and g.number_of_nodes() == 50 and g.number_of_edges() == 600
    assert np.array_equal(g.in_degrees(range(50)).numpy(), np.bincount(want["dst"], minlength=50))


def test_adjacency_lists_follow_reference_order():
    t = _train_triples(6, 20, 3, 60)
    t[3] = (4, 1, 4)                      # self loop: appears twice in its list
    adj, deg = K.utils.get_adj_and_degrees(20, t)
    ref = [[] for _ in range(20)]
    for i, (s, _, o) in enumerate(t):     # kgvae/utils.py:24-26
        ref[s].append([i, o])
        ref[o].append([i, s])
    for v in range(20):
        assert deg[v] == len(ref[v])
        assert np.array_equal(adj[v].reshape(-1, 2), np.array(ref[v]).reshape(-1, 2))


def test_neighbor_sampler_is_reproducible_and_valid():
    t = _train_triples(8, 40, 3, 200)
    adj, deg = K.utils.get_adj_and_degrees(40, t)
    np.random.seed(1)
    a = K.utils.sample_edge_neighborhood(adj, deg, len(t), 50)
    np.random.seed(1)
    b = K.utils.sample_edge_neighborhood(adj, deg, len(t), 50)
    assert np.array_equal(a, b) and len(set(a.tolist())) == 50


def test_state_dict_keys_match_reference(golden):
    gv = golden("kgvae_tiny_flow3")
    n_ent, n_rel, h, bases, k, n_flows, _ = (int(x) for x in gv["cfg"])
    model = K.LinkPredict(K.KGVAE, n_ent, h, n_rel, num_bases=bases, k=k, n_flows=n_flows)
    want = {key[len("param/"):]: val.shape for key, val in gv.items() if key.startswith("param/")}
    got = {key: tuple(val.shape) for key, val in model.state_dict().items()}
    assert set(got) == set(want)
    for key in want:
        assert got[key] == tuple(want[key]), key
    model.load_state_dict({key: torch.from_numpy(gv["param/" + key]) for key in want})


def test_made_masks_and_pass_plan(golden):
    gv = golden("made_block")
    D, nh, _ = (int(x) for x in gv["cfg"])
    made = K.MADE(D, D, nh)
    for l in range(nh + 2):
        assert np.array_equal(made.net[2 * l].mask.numpy(), gv[f"param/net.{2 * l}.mask"])
    for i, deg in enumerate(made.m):
        assert np.array_equal(deg.numpy(), gv[f"deg/{i}"])
    mid = [2] + [1] * (D - 2) + [0]
    assert [m.tolist() for m in made._col_mult] == [[1] * D] + [mid] * (nh + 1) + [[1] * D]


def test_synthetic_shapes():
    d = K.datasets.synthetic_kg("FB15k-237", seed=0, scale=0.01)
    assert d.num_nodes == 14541 and d.num_rels == 237 and d.train.shape == (2721, 3)
    assert d.train[:, 1].max() < 237 and d.train[:, [0, 2]].max() < 14541
    z = K.datasets.synthetic_kg("toy", seed=0, skew=1.0)
    assert np.bincount(z.train[:, 0]).max() > 5 * len(z.train) / z.num_nodes


def test_filter_csr():
    allt = np.array([[0, 0, 1], [0, 0, 2], [0, 0, 2], [3, 1, 0], [4, 0, 1]])
    ptr, idx = K.utils.build_filter(allt, [0, 3, 9], [0, 1, 0], 2, "object", "cpu")
    assert ptr.tolist() == [0, 2, 3, 3] and idx.tolist() == [1, 2, 0]
    ptr, idx = K.utils.build_filter(allt, [1, 2], [0, 0], 2, "subject", "cpu")
    assert ptr.tolist() == [0, 2, 3] and idx.tolist() == [0, 4, 0]


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU oracle port on the host cores) prints ONE JSON line with the keys the
    driver reads; runs without a GPU.  The non-zero ranks of a torchrun launch print nothing."""
    import json
    import os
    import subprocess
    import sys
    from conftest import ROOT
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["config"]["workload"] == "fb15k237-full"
    other = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                           capture_output=True, text=True, timeout=60, env=dict(env, RANK="1", WORLD_SIZE="2"))
    assert other.returncode == 0 and other.stdout.strip() == ""


def test_neighbor_sampler_matches_reference_golden(golden):
    """utils.sample_edge_neighborhood and the "neighbor" mode of generate_sampled_graph_and_labels reproduce
    the reference's picks integer for integer (kgvae/utils.py:33-124 on the legacy global numpy stream;
    fixture produced by the unmodified reference)."""
    gv = golden("neighbor_sampler_seed5")
    n_ent, n_rel, n_train, sample = (int(v) for v in gv["cfg"])
    train = gv["train_triples"]
    adj, deg = K.utils.get_adj_and_degrees(n_ent, train)
    np.random.seed(5)
    edges = K.utils.sample_edge_neighborhood(adj, deg, n_train, sample)
    assert np.array_equal(np.asarray(edges, dtype=np.int64), gv["edges"])
    np.random.seed(6)
    g, node_id, edge_type, node_norm, data, labels = K.utils.generate_sampled_graph_and_labels(
        train, sample, 0.5, n_rel, adj, deg, 3, "neighbor")
    assert np.array_equal(node_id, gv["node_id"]) and np.array_equal(edge_type, gv["edge_type"])
    assert np.array_equal(data, gv["samples"])
    assert np.array_equal(g._src, gv["g_src"]) and np.array_equal(g._dst, gv["g_dst"])


def test_entity_classify_directory_loader(tmp_path):
    """N4: the on-disk typed-graph loader that stands in for dgl.contrib.data.load_data (kgvae/entity_classify.py:47):
    inverse relation types, self-loops, (dst, src, type) order, per-(dst, type) norms, bfs pruning, relabelling."""
    from gcn_vae_b200 import entity_classify as EC
    (tmp_path / "edges.tsv").write_text("a\tknows\tb\nb\tknows\tc\nc\tlikes\ta\nd\tlikes\te\nx\tknows\ty\n")
    (tmp_path / "trainingSet.tsv").write_text("a\tperson\nc\trobot\n")
    (tmp_path / "testSet.tsv").write_text("b\tperson\n")
    full = EC.load_data(str(tmp_path), bfs_level=0)
    n, E = full.num_nodes, len(full.edge_src)
    assert (n, full.num_rels, full.num_classes) == (7, 5, 2) and E == 2 * 5 + 7
    order = np.lexsort((full.edge_type, full.edge_src, full.edge_dst))
    assert np.array_equal(order, np.arange(E))                       # sorted by (dst, src, type)
    assert np.all(full.edge_type[full.edge_src == full.edge_dst] == 4)   # self-loop relation = 2 P
    for d, t_, w_ in zip(full.edge_dst, full.edge_type, full.edge_norm):
        assert abs(w_ - 1.0 / np.sum((full.edge_dst == d) & (full.edge_type == t_))) < 1e-7
    assert list(full.labels[full.train_idx]) == [0, 1] and list(full.labels[full.test_idx]) == [0]
    # one round from the labelled nodes {a, b, c}: only edges INTO them survive; d, e, x, y are never reached
    pruned = EC.load_data(str(tmp_path), bfs_level=1)
    assert set(pruned.edge_dst.tolist()) <= {0, 1, 2} and len(pruned.edge_src) == 3 + 6
    small = EC.load_data(str(tmp_path), bfs_level=2, relabel=True)
    assert small.num_nodes == 3 and small.edge_src.max() < 3 and small.edge_dst.max() < 3
    toy = EC.load_data("toy:3")
    assert toy.num_nodes == 300 and len(toy.edge_src) == 2500


def test_block_chunks_for_the_column_pipeline():
    """ops._block_chunks: contiguous block ranges the column-chunk kernels accept (5x5 / 5x10 blocks, cuts at
    multiples of 4 blocks so that every chunk's columns start 16-byte aligned, at most 64 blocks per chunk)."""
    from gcn_vae_b200 import ops
    assert ops._block_chunks(100, 5, 5, 2) == [(0, 48), (48, 100)]
    assert ops._block_chunks(100, 5, 10, 2) == [(0, 48), (48, 100)]
    assert ops._block_chunks(8, 5, 5, 2) == [(0, 4), (4, 8)]
    assert ops._block_chunks(100, 5, 10, 1) == []                 # one collective per layer
    assert ops._block_chunks(25, 20, 20, 2) == []                 # other block shapes: no chunked kernels
    assert ops._block_chunks(6, 5, 5, 2) == []                    # not a multiple of 4 blocks
    assert ops._block_chunks(200, 5, 5, 2) == []                  # chunks would exceed 64 blocks
    for B, n in ((100, 2), (100, 4), (64, 2), (128, 4)):
        ch = ops._block_chunks(B, 5, 10, n)
        if ch:
            assert ch[0][0] == 0 and ch[-1][1] == B and all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
            assert all((b1 - b0) % 2 == 0 and b0 % 4 == 0 for b0, b1 in ch)


def test_captured_train_step_rejects_what_it_cannot_capture():
    """link_predict.CapturedTrainStep checks its preconditions before touching the device: a capturable optimizer
    (device-resident step counters) and CUDA inputs (there is no CPU fallback)."""
    import pytest
    model = K.LinkPredict(K.KGVAE, 20, 8, 3, num_bases=2, use_cuda=True, k=2)
    ex = {"node_id": torch.arange(20, dtype=torch.int32).view(-1, 1), "src": torch.zeros(4, dtype=torch.int32),
          "dst": torch.ones(4, dtype=torch.int32), "etype": torch.zeros(4, dtype=torch.int32), "norm": torch.ones(4, 1),
          "samples": torch.zeros(5, 3, dtype=torch.int32), "labels": torch.zeros(5)}
    with pytest.raises(RuntimeError, match="capturable"):
        K.link_predict.CapturedTrainStep(model, torch.optim.Adam(model.parameters()), ex, 20)
    with pytest.raises(RuntimeError, match="CUDA"):
        K.link_predict.CapturedTrainStep(model, torch.optim.Adam(model.parameters(), capturable=True), ex, 20)


def test_graph_from_device_edges_keeps_the_host_surface():
    """Graph.from_device_edges adopts an edge list without copying it; the DGLGraph queries the reference's loop makes
    (g.in_degrees, g.edges, number_of_edges; kgvae/link_predict.py:216) still answer from it."""
    src, dst = torch.tensor([0, 1, 2, 2], dtype=torch.int32), torch.tensor([1, 2, 0, 1], dtype=torch.int32)
    g = K.Graph.from_device_edges(3, src, dst)
    assert g.number_of_nodes() == 3 and g.number_of_edges() == 4
    assert g.in_degrees(range(3)).tolist() == [1, 2, 1]
    s, d = g.edges()
    assert s.tolist() == [0, 1, 2, 2] and d.tolist() == [1, 2, 0, 1]


def test_release_graph_detaches_the_cached_encoder_outputs():
    """LinkPredict.release_graph: the tensors KGVAE keeps between forward and get_loss (kgvae/model.py:113-123) stop
    holding the autograd graph of the last step - which is what lets a step be captured on another stream."""
    model = K.LinkPredict(K.KGVAE, 20, 8, 3, num_bases=2, use_cuda=True, k=2)
    x = torch.randn(20, 8, requires_grad=True)
    model.encoder.z_mean, model.encoder.z_sigma = x * 2, x.exp()
    model.encoder.flow_log_prob = None
    model.release_graph()
    assert model.encoder.z_mean.grad_fn is None and model.encoder.z_sigma.grad_fn is None
    assert torch.equal(model.encoder.z_mean, (x * 2).detach())
