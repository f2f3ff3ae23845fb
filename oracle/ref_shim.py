"""Import shim that runs the UNMODIFIED reference modules from ``/root/reference``.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Only usable in the build container (the GPU
box has no ``/root/reference``); it is used by ``tests/golden/make_golden.py`` to mint golden
vectors, never at test / bench run time.

The reference imports packages that are absent here (``torch_geometric``, ``matplotlib``,
``seaborn``).  The stubs below provide exactly the names the reference touches:
``Data`` / ``Dataset`` attribute bags, ``GCNConv`` (= the restatement in ``oracle/gcn.py``),
and inert plotting modules.  ``src/setup.py`` parses ``sys.argv`` at import (``:53``), writes
``pangnn.log`` into the CWD and ``pangnn.py`` ``rmtree``s ``temp/`` — so the caller must give a
scratch working directory; ``data`` is symlinked into it.
"""
import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("PANGNN_REFERENCE_ROOT", "/root/reference")


class Data:
    """``torch_geometric.data.Data(x, edge_index, edge_attr, y)`` as an attribute bag."""

    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, **kw):
        self.x, self.edge_index, self.edge_attr, self.y = x, edge_index, edge_attr, y
        for k, v in kw.items():
            setattr(self, k, v)

    def to_dict(self):
        return dict(self.__dict__)

    def from_dict(self, d):
        self.__dict__.update(d)
        return self

    def to(self, device):
        import torch
        for k, v in self.__dict__.items():
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self

    def cpu(self):
        return self.to("cpu")


class Dataset:
    def __init__(self, root=None, transform=None, pre_transform=None, pre_filter=None):
        pass


def install_stubs():
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if here not in sys.path:
        sys.path.insert(0, here)
    from oracle.gcn import GCNConv

    tg = types.ModuleType("torch_geometric")
    tg_data = types.ModuleType("torch_geometric.data")
    tg_data.Data, tg_data.Dataset = Data, Dataset
    tg_nn = types.ModuleType("torch_geometric.nn")
    tg_nn.GCNConv, tg_nn.ChebConv, tg_nn.MessagePassing = GCNConv, GCNConv, object
    tg_loader = types.ModuleType("torch_geometric.loader")
    tg_loader.DataLoader = MagicMock()
    tg_utils = types.ModuleType("torch_geometric.utils")
    tg_utils.to_networkx = MagicMock()
    tg_conv = types.ModuleType("torch_geometric.utils.convert")
    tg_conv.to_scipy_sparse_matrix = MagicMock()
    tg_tf = types.ModuleType("torch_geometric.transforms")
    tg_tf.RemoveDuplicatedEdges = MagicMock()
    tg.data, tg.nn, tg.loader, tg.utils, tg.transforms = tg_data, tg_nn, tg_loader, tg_utils, tg_tf
    tg_utils.convert = tg_conv
    sys.modules.update({
        "torch_geometric": tg, "torch_geometric.data": tg_data, "torch_geometric.nn": tg_nn,
        "torch_geometric.loader": tg_loader, "torch_geometric.utils": tg_utils,
        "torch_geometric.utils.convert": tg_conv, "torch_geometric.transforms": tg_tf,
    })
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        sys.modules.setdefault(name, MagicMock())


def load_reference(argv, workdir):
    """chdir to ``workdir``, set ``sys.argv`` and import the reference's ``src`` package.
    Returns a namespace with the imported modules.  One call per process (argparse at import)."""
    import torch.multiprocessing
    torch.multiprocessing.set_sharing_strategy("file_system")   # fd limit here is 20 000 (F8)
    os.makedirs(workdir, exist_ok=True)
    link = os.path.join(workdir, "data")
    if not os.path.exists(link):
        os.symlink(os.path.join(REFERENCE_ROOT, "data"), link)
    os.chdir(workdir)
    for d in ("plots", "temp", "runs"):
        os.makedirs(d, exist_ok=True)
    install_stubs()
    sys.argv = ["pangnn.py"] + list(argv)
    sys.path.insert(0, REFERENCE_ROOT)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):            # the banner
        import src.setup as setup
    import src.preprocessing as preprocessing
    import src.helper as helper
    import src.simulate as simulate
    import src.dataset as dataset
    import src.gnn as gnn
    return types.SimpleNamespace(setup=setup, args=setup.args, preprocessing=preprocessing,
                                 helper=helper, simulate=simulate, dataset=dataset, gnn=gnn)
