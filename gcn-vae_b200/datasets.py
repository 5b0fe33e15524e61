"""Datasets: synthetic knowledge graphs of the named benchmark shapes, and a plain-text loader.

Stands in for ``dgl.contrib.data.load_data`` (kgvae/link_predict.py:105-110), which downloads
FB15k-237 / wn18 and is unavailable offline.  ``load_data(name)`` returns an object with the
same fields the reference reads: ``num_nodes``, ``num_rels``, ``train``, ``valid``, ``test``
(int64 arrays of (s, r, o) rows).
"""
import os
from types import SimpleNamespace

import numpy as np

# name -> (entities, relations, train, valid, test)
SHAPES = {
    "FB15k-237": (14541, 237, 272115, 17535, 20466),
    "wn18": (40943, 18, 141442, 5000, 5000),
    "wikikg2": (2500604, 535, 16109182, 429456, 598543),
    "toy": (500, 12, 4000, 300, 300),
}


def synthetic_kg(name="FB15k-237", seed=0, skew=0.0, scale=1.0):
    """Seeded random triples of a named shape.  ``skew`` > 0 draws entities and relations from a
    Zipf-like law (heavy-tailed degrees, hot relations); 0 is uniform.  ``scale`` shrinks the
    triple counts (not the entity / relation counts) for bounded CPU samples."""
    n_ent, n_rel, n_train, n_valid, n_test = SHAPES[name]
    rng = np.random.default_rng(seed)

    def draw(n_items, size):
        if skew <= 0:
            return rng.integers(0, n_items, size=size)
        p = 1.0 / np.arange(1, n_items + 1) ** skew
        perm = rng.permutation(n_items)
        return perm[rng.choice(n_items, size=size, p=p / p.sum())]

    def triples(n):
        n = max(1, int(n * scale))
        return np.stack([draw(n_ent, n), draw(n_rel, n), draw(n_ent, n)], axis=1).astype(np.int64)

    return SimpleNamespace(name=name, num_nodes=n_ent, num_rels=n_rel, train=triples(n_train),
                           valid=triples(n_valid), test=triples(n_test), synthetic=True)


def _read_triples(path, ent, rel):
    rows = []
    with open(path) as f:
        for line in f:
            s, r, o = line.split()
            rows.append((ent.setdefault(s, len(ent)), rel.setdefault(r, len(rel)), ent.setdefault(o, len(ent))))
    return np.asarray(rows, dtype=np.int64)


def load_data(dataset):
    """``<dir>`` containing train.txt / valid.txt / test.txt (tab-separated s r o), or
    ``synthetic:<shape>[:seed[:skew]]`` / a bare shape name for a synthetic graph."""
    if os.path.isdir(dataset):
        ent, rel = {}, {}
        train, valid, test = (_read_triples(os.path.join(dataset, f + ".txt"), ent, rel)
                              for f in ("train", "valid", "test"))
        return SimpleNamespace(name=dataset, num_nodes=len(ent), num_rels=len(rel), train=train,
                               valid=valid, test=test, synthetic=False)
    parts = dataset.split(":")
    if parts[0] == "synthetic":
        parts = parts[1:]
    if parts[0] not in SHAPES:
        raise ValueError(f"unknown dataset {dataset!r}; known shapes: {sorted(SHAPES)}")
    seed = int(parts[1]) if len(parts) > 1 else 0
    skew = float(parts[2]) if len(parts) > 2 else 0.0
    return synthetic_kg(parts[0], seed, skew)
