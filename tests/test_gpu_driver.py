"""The training / evaluation driver (kgvae/link_predict.py:103-268) on a synthetic toy graph:
the reference's flags, train loop, periodic validation with checkpoints, --load and --test-mode, with the
host sampler (bit-exact numpy stream) and the opt-in device sampler."""
import os

import numpy as np
import pytest
import torch

import gcn_vae_b200 as K

pytestmark = pytest.mark.gpu


def _args(tmp_path, *extra):
    argv = ["-d", "toy", "--gpu", "0", "--n-hidden", "40", "--n-bases", "8", "--graph-batch-size", "600",
            "--negative-sample", "4", "--n-epochs", "4", "--evaluate-every", "2", "--eval-batch-size", "64",
            "--mog-k", "4", "--n-flows", "1", "--kl-param", "1e-3",
            "--model-state-file", os.path.join(str(tmp_path), "model_state.pth"), *extra]
    return K.link_predict.build_parser().parse_args(argv)


@pytest.mark.parametrize("device_sampler", [False, True])
def test_driver_trains_validates_and_checkpoints(tmp_path, device_sampler, capsys):
    np.random.seed(0)
    torch.manual_seed(0)
    args = _args(tmp_path, *(["--device-sampler"] if device_sampler else []))
    best = K.link_predict.main(args)
    out = capsys.readouterr().out
    assert "Epoch 0004" in out and "start eval" in out and "training done" in out
    assert 0.0 < best <= 1.0
    ckpt = torch.load(args.model_state_file, map_location="cpu")
    assert ckpt["epoch"] in (2, 4)
    # the reference's state-dict keys so that its checkpoints load
    for key in ("encoder.input_layer.embedding.weight", "encoder.rconv_layer_1.weight", "encoder.rconv_layer_1.h_bias",
                "encoder.rconv_layer_1.loop_weight", "encoder.rconv_layer_2.weight", "encoder.z_pre", "encoder.pi",
                "encoder.nf.0.net.0.weight", "encoder.nf.0.net.0.mask", "w_relation"):
        assert key in ckpt["state_dict"], key


def test_driver_test_mode_loads_the_checkpoint(tmp_path, capsys):
    np.random.seed(1)
    torch.manual_seed(1)
    K.link_predict.main(_args(tmp_path, "--n-epochs", "2"))
    capsys.readouterr()
    mrr = K.link_predict.main(_args(tmp_path, "--test-mode", "1"))
    out = capsys.readouterr().out
    assert "start testing" in out and "Using best epoch: 2" in out and "MRR (raw)" in out
    assert 0.0 < mrr <= 1.0


def test_driver_readme_configuration_terms(tmp_path, capsys):
    """The README run's loss terms together (kgvae/README.md:5-6: --n-flows 3 --mmd-param 1 --mog-k 10), at toy size."""
    np.random.seed(2)
    torch.manual_seed(2)
    best = K.link_predict.main(_args(tmp_path, "--n-flows", "3", "--mmd-param", "1", "--mog-k", "10", "--n-epochs", "2"))
    out = capsys.readouterr().out
    assert "Epoch 0002" in out and "mmd" in out and 0.0 < best <= 1.0


def test_driver_generate_writes_result_file(tmp_path, capsys, monkeypatch):
    """--test-mode --generate (kgvae/link_predict.py:181-184): top-1 tail per test query through the tcgen05 score
    pass with a top-k epilogue; the highest-scored tail must be the argmax of the dense score matrix."""
    np.random.seed(3)
    torch.manual_seed(3)
    K.link_predict.main(_args(tmp_path, "--n-epochs", "2"))
    capsys.readouterr()
    monkeypatch.chdir(tmp_path)
    tails = K.link_predict.main(_args(tmp_path, "--test-mode", "1", "--generate", "1"))
    assert tails.shape[1] == 1 and (tmp_path / "result.txt").exists()
    assert "generated tails" in capsys.readouterr().out


def test_entity_classify_driver_on_disk_dataset(tmp_path, capsys):
    """kgvae/entity_classify.py:45-135 end to end: the directory loader (N4), EntityClassify on integer-id features,
    cross-entropy training, test accuracy; the loss must go down on a learnable toy problem."""
    from gcn_vae_b200 import entity_classify as EC
    rng = np.random.default_rng(0)
    n, n_cls = 120, 3
    cls = rng.integers(0, n_cls, n)
    lines = []
    for _ in range(900):                       # same-class nodes are linked by relation "c<k>"
        u = int(rng.integers(0, n))
        same = np.nonzero(cls == cls[u])[0]
        lines.append(f"n{u}\tc{cls[u]}\tn{int(rng.choice(same))}")
    (tmp_path / "edges.tsv").write_text("\n".join(lines) + "\n")
    perm = rng.permutation(n)
    (tmp_path / "trainingSet.tsv").write_text("".join(f"n{i}\tk{cls[i]}\n" for i in perm[:80]))
    (tmp_path / "testSet.tsv").write_text("".join(f"n{i}\tk{cls[i]}\n" for i in perm[80:]))
    torch.manual_seed(0)
    args = EC.build_parser().parse_args(["-d", str(tmp_path), "--gpu", "0", "--n-hidden", "16", "--n-bases", "-1",
                                         "-e", "40", "--testing", "--l2norm", "0"])
    args.bfs_level = args.n_layers + 1
    res = EC.main(args)
    out = capsys.readouterr().out
    assert "Epoch 00039" in out and "Test Accuracy" in out
    assert res["train_loss"] < 1.0 and res["test_acc"] > 0.5


def _toy_step_inputs(dev, n=300, n_rel=6, e=2400, s=3000, seed=0):
    g = torch.Generator().manual_seed(seed)
    src, dst = torch.randint(0, n, (e,), generator=g), torch.randint(0, n, (e,), generator=g)
    deg = torch.bincount(dst, minlength=n).clamp(min=1).float()
    t = {"node_id": torch.arange(n, dtype=torch.int32).view(-1, 1), "src": src.int(), "dst": dst.int(),
         "etype": torch.randint(0, 2 * n_rel, (e,), generator=g).int(), "norm": (1.0 / deg)[dst].view(-1, 1),
         "samples": torch.stack([torch.randint(0, n, (s,), generator=g), torch.randint(0, n_rel, (s,), generator=g),
                                 torch.randint(0, n, (s,), generator=g)], 1).int(),
         "labels": (torch.rand(s, generator=g) < 0.3).float()}
    return {k: v.to(dev) for k, v in t.items()}


@pytest.mark.parametrize("n_flows", [0, 1])
def test_captured_train_step_replays_the_eager_step(n_flows):
    """link_predict.CapturedTrainStep: the CUDA-graph replay of the train step (kgvae/link_predict.py:217-228)
    leaves the model where the same number of eager steps leaves it - same noise (the Philox seed is re-read at
    every replay), same gradients, same Adam updates - on the captured inputs and on new inputs copied into place."""
    import copy
    dev = torch.device("cuda:0")
    n, n_rel = 300, 6
    torch.manual_seed(0)
    base = K.LinkPredict(K.KGVAE, n, 40, n_rel, num_bases=8, dropout=0.2, use_cuda=True, reg_param=0.01,
                         kl_param=1e-3, k=4, n_flows=n_flows).to(dev)
    inputs = [_toy_step_inputs(dev, n, n_rel, seed=i) for i in range(3)]

    def eager_run():
        m = copy.deepcopy(base).train()
        opt = torch.optim.Adam(m.parameters(), lr=1e-3, fused=True, capturable=True)
        bk = m.grad_buckets()
        losses = []
        for i, t in enumerate(inputs):
            torch.manual_seed(100 + i)
            g = K.Graph.from_device_edges(n, t["src"], t["dst"])
            bk.zero()
            emb = m(g, t["node_id"], t["etype"], t["norm"])
            loss = m.get_loss(g, emb, t["samples"], t["labels"])[0]
            loss.backward()
            bk.finish()
            bk.clip_(1.0)
            opt.step()
            losses.append(float(loss))
        return m, losses

    def captured_run():
        m = copy.deepcopy(base).train()
        opt = torch.optim.Adam(m.parameters(), lr=1e-3, fused=True, capturable=True)
        torch.manual_seed(100)
        cap = K.link_predict.CapturedTrainStep(m, opt, inputs[0], n, grad_norm=1.0, warmup=1)   # warm-up = step 0
        assert cap.launches_per_step > 10
        losses = []
        for i in (1, 2):
            torch.manual_seed(100 + i)
            losses.append(float(cap.step(**inputs[i])))
        return m, losses

    m_e, l_e = eager_run()
    m_c, l_c = captured_run()
    assert np.allclose(l_e[1:], l_c, rtol=1e-5), (l_e, l_c)
    for (name, p), q in zip(m_e.named_parameters(), m_c.parameters()):
        # Adam normalises the update: an element whose gradient is fp32 reduction noise moves by up to lr either way
        err = float((p - q).abs().max() / p.abs().max().clamp(min=1e-12))
        assert err < 1e-3, (name, err)


def test_captured_train_step_needs_a_capturable_optimizer():
    dev = torch.device("cuda:0")
    m = K.LinkPredict(K.KGVAE, 300, 40, 6, num_bases=8, use_cuda=True, k=2).to(dev)
    with pytest.raises(RuntimeError, match="capturable"):
        K.link_predict.CapturedTrainStep(m, torch.optim.Adam(m.parameters(), fused=True), _toy_step_inputs(dev), 300)


def test_full_batch_device_sampler_follows_the_reference_procedure():
    """utils.FullBatchDeviceSampler (kgvae/utils.py:79-124,158-171 with sample_size = all training triples): positives
    kept, every negative differs from its positive in exactly the subject OR the object (about half each, replacement
    among the sampled nodes), labels 1 / 0, the graph is a random ``split_size`` of the positives in both directions with
    reverse relation ids and 1 / in-degree norms; two draws differ."""
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    n_ent, n_rel, T, rate = 500, 7, 4000, 5
    train = np.stack([rng.integers(0, n_ent, T), rng.integers(0, n_rel, T), rng.integers(0, n_ent, T)], 1)
    sm = K.utils.FullBatchDeviceSampler(torch.from_numpy(train).to(dev), n_rel, rate, split_size=0.5)
    torch.manual_seed(0)
    a, b = sm.sample(), sm.sample()
    uniq = np.unique(np.concatenate([train[:, 0], train[:, 2]]))
    assert np.array_equal(a["node_id"].cpu().numpy().reshape(-1), uniq) and sm.n == len(uniq)
    relabel = {v: i for i, v in enumerate(uniq)}
    pos = np.array([[relabel[s], r, relabel[o]] for s, r, o in train])
    smp = a["samples"].cpu().numpy()
    assert smp.shape == (T * (rate + 1), 3) and np.array_equal(smp[:T], pos)
    neg = smp[T:].reshape(rate, T, 3)
    same_s, same_o = neg[:, :, 0] == pos[None, :, 0], neg[:, :, 2] == pos[None, :, 2]
    assert np.array_equal(neg[:, :, 1], np.broadcast_to(pos[None, :, 1], (rate, T)))
    assert (same_s | same_o).all()                         # at most one end is replaced
    frac_head = float((~same_s).mean())
    assert 0.45 < frac_head < 0.55 and 0 <= smp.min() and smp[:, [0, 2]].max() < sm.n
    lab = a["labels"].cpu().numpy()
    assert lab[:T].min() == 1 and lab[T:].max() == 0
    src, dst, et = (a[k].cpu().numpy() for k in ("src", "dst", "etype"))
    E = len(src) // 2
    assert E == T // 2 and np.array_equal(src[:E], dst[E:]) and np.array_equal(dst[:E], src[E:])
    assert np.array_equal(et[E:], et[:E] + n_rel) and et[:E].max() < n_rel
    pos_set = {tuple(p) for p in pos}
    assert all((s, r, o) in pos_set for s, r, o in zip(src[:E], et[:E], dst[:E]))
    deg = np.bincount(dst, minlength=sm.n)
    assert np.allclose(a["norm"].cpu().numpy().reshape(-1), 1.0 / deg[dst])
    assert not np.array_equal(smp[T:], b["samples"].cpu().numpy()[T:]) and not np.array_equal(src, b["src"].cpu().numpy())


def test_captured_step_with_its_own_sampler_trains():
    """CapturedTrainStep(sampler=...): the whole loop iteration of kgvae/link_predict.py:200-236 - fresh negatives and
    graph split, edge index, forward, loss, backward, clip, Adam - is one replay; the loss goes down over replays and
    successive replays see different samples (the loss sequence is not constant)."""
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(5)
    n_ent, n_rel, T = 300, 6, 3000
    train = np.stack([rng.integers(0, n_ent, T), rng.integers(0, n_rel, T), rng.integers(0, n_ent, T)], 1)
    sm = K.utils.FullBatchDeviceSampler(torch.from_numpy(train).to(dev), n_rel, 4)
    torch.manual_seed(0)
    m = K.LinkPredict(K.KGVAE, sm.n, 40, n_rel, num_bases=8, dropout=0.2, use_cuda=True, reg_param=0.01,
                      kl_param=1e-3, k=4, n_flows=0).to(dev)
    opt = torch.optim.Adam(m.parameters(), lr=1e-2, fused=True, capturable=True)
    cap = K.link_predict.CapturedTrainStep(m, opt, None, sm.n, warmup=2, sampler=sm)
    losses = [float(cap.step()) for _ in range(40)]
    assert all(np.isfinite(losses)) and len(set(losses)) > 30
    assert np.mean(losses[-5:]) < 0.8 * np.mean(losses[:5])
    with pytest.raises(RuntimeError, match="draws its own inputs"):
        cap.step(labels=torch.zeros(1, device=dev))


def test_driver_capture_step_trains_validates_and_checkpoints(tmp_path, capsys):
    """--capture-step: the reference's loop (kgvae/link_predict.py:200-259) with every iteration - sampler included -
    as one CUDA-graph replay; validation, checkpoints and the printed line are the eager loop's."""
    np.random.seed(0)
    torch.manual_seed(0)
    args = _args(tmp_path, "--capture-step", "--graph-batch-size", "1000000", "--n-flows", "0", "--n-epochs", "6")
    best = K.link_predict.main(args)
    out = capsys.readouterr().out
    assert "Epoch 0002" in out and "Epoch 0006" in out and "start eval" in out and "training done" in out
    losses = [float(line.split("Loss")[1].split("|")[0]) for line in out.splitlines() if line.startswith("Epoch")]
    assert len(losses) == 5 and all(np.isfinite(losses)) and losses[-1] < losses[0]
    assert 0.0 < best <= 1.0
    assert torch.load(args.model_state_file, map_location="cpu")["epoch"] in (2, 4, 6)
    with pytest.raises(RuntimeError, match="full-batch"):
        K.link_predict.main(_args(tmp_path, "--capture-step", "--graph-batch-size", "10"))
