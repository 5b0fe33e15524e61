"""The C-ABI library loads and exports every symbol include/kgvae_b200.h declares, and the
ctypes table matches the header's parameter counts.  No compute calls: runs without a GPU."""
import os
import re

import pytest

from conftest import ROOT

import gcn_vae_b200 as K


def header_prototypes():
    text = open(os.path.join(ROOT, "include", "kgvae_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|size_t|const char\*)\s+(kg_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        protos[m.group(1)] = n
    return protos


def test_library_exports_every_declared_symbol():
    protos = header_prototypes()
    assert len(protos) >= 30
    handle = K._lib.lib()
    for name in protos:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"


def test_ctypes_table_matches_header():
    protos = header_prototypes()
    table = K._lib._SIGNATURES
    assert set(table) == set(protos), set(table) ^ set(protos)
    for name, n in protos.items():
        assert len(table[name][1]) == n, f"{name}: header has {n} parameters, ctypes table {len(table[name][1])}"


def test_version_and_error_string():
    handle = K._lib.lib()
    assert handle.kg_version() >= 100
    assert isinstance(handle.kg_last_error(), bytes)


def test_ops_fail_loudly_without_cuda_tensors():
    import torch
    with pytest.raises(RuntimeError, match="CUDA"):
        K.ops.gemm(torch.zeros(2, 2), torch.zeros(2, 2), torch.zeros(2, 2))
    layer = K.RelGraphConv(4, 4, 2, "bdd", 2, self_loop=True)
    g = K.Graph()
    g.add_nodes(2)
    g.add_edges([0], [1])
    with pytest.raises(RuntimeError, match="CUDA"):
        layer(g, torch.zeros(2, 4), torch.zeros(1, dtype=torch.long), torch.ones(1, 1))
