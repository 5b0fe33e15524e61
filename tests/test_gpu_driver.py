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
