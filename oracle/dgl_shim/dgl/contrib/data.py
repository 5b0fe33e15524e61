"""``dgl.contrib.data.load_data`` stand-in: no datasets exist offline."""


def load_data(*args, **kwargs):
    raise RuntimeError("dgl.contrib.data.load_data is unavailable offline; "
                       "use the synthetic generators in gcn-vae_b200/datasets.py")
