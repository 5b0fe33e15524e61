"""Pure-PyTorch stand-in for the slice of DGL 0.4.x that the reference touches.

TEST INFRASTRUCTURE ONLY.  This package exists so that the reference's own
``kgvae/{model,utils,link_predict,flow_network}.py`` can be imported *verbatim*
from ``/root/reference`` inside the build container (DGL itself is not
installable here: no network, un-pinned dependency).
It is used by ``tests/golden/make_golden.py`` to generate the committed golden
vectors and by nothing on the product path.

Surface restated (DGL 0.4.x, PyPI ``dgl``; the reference pins no version - API
usage and the 2019-12 bytecode imply 0.4.1):

* ``dgl.DGLGraph()``: ``add_nodes``, ``add_edges``, ``local_var``, ``ndata``,
  ``edata``, ``apply_edges``, ``in_degrees``, ``number_of_nodes``, ``__len__``
  (call sites: reference kgvae/utils.py:127-150, kgvae/link_predict.py:95-100,216)
* ``dgl.nn.pytorch.RelGraphConv`` (bdd + basis)  (kgvae/model.py:54-59)
* ``dgl.contrib.data.load_data`` (raises: there are no datasets offline)
"""
import numpy as np
import torch


class _EdgeBatch:
    """What DGL hands to an ``apply_edges`` UDF: views of src/dst/edge frames."""

    def __init__(self, g):
        self.src = {k: v[g._src] for k, v in g.ndata.items()}
        self.dst = {k: v[g._dst] for k, v in g.ndata.items()}
        self.data = dict(g.edata)


class DGLGraph:
    def __init__(self):
        self._n = 0
        self._src = torch.zeros(0, dtype=torch.long)
        self._dst = torch.zeros(0, dtype=torch.long)
        self.ndata = {}
        self.edata = {}

    # -- construction -----------------------------------------------------
    def add_nodes(self, n):
        self._n += int(n)

    def add_edges(self, src, dst):
        src = torch.as_tensor(np.asarray(src), dtype=torch.long)
        dst = torch.as_tensor(np.asarray(dst), dtype=torch.long)
        self._src = torch.cat([self._src, src])
        self._dst = torch.cat([self._dst, dst])

    # -- queries ----------------------------------------------------------
    def number_of_nodes(self):
        return self._n

    def number_of_edges(self):
        return int(self._src.numel())

    def __len__(self):
        return self._n

    def in_degrees(self, v=None):
        deg = torch.bincount(self._dst, minlength=self._n)
        if v is None:
            return deg
        return deg[torch.as_tensor(list(v), dtype=torch.long)]

    def local_var(self):
        g = DGLGraph()
        g._n, g._src, g._dst = self._n, self._src, self._dst
        g.ndata = dict(self.ndata)
        g.edata = dict(self.edata)
        return g

    # -- message passing --------------------------------------------------
    def apply_edges(self, func):
        self.edata.update(func(_EdgeBatch(self)))

    def sum_messages(self, msg):
        """``update_all(udf, fn.sum)``: zero-filled scatter-sum over destinations."""
        out = torch.zeros((self._n,) + tuple(msg.shape[1:]), dtype=msg.dtype)
        return out.index_add(0, self._dst, msg)
