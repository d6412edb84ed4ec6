"""The PyG drop-in surface (B1): importability of every name the reference's Code/sag
uses, state-dict compatibility, and -- on the GPU -- the op-level drop-ins against the oracle when
driven exactly the way layers.py / network.py drive them."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SAG = "/root/reference/Code/sag"


@pytest.fixture()
def shim():
    from tsg import run
    run.install()
    yield


def test_surface_names_import(shim):
    from torch_geometric.nn import GCNConv, GraphConv, TopKPooling                       # network.py:2-3
    from torch_geometric.nn import global_mean_pool as gap, global_max_pool as gmp       # network.py:4
    from torch_geometric.nn.pool.topk_pool import topk, filter_adj                       # layers.py:2
    from torch_geometric.datasets import TUDataset                                       # train*.py:5
    from torch_geometric.data import DataLoader                                          # train*.py:6
    from torch_geometric import utils                                                    # train*.py:7
    conv = GCNConv(89, 32)
    assert tuple(conv.weight.shape) == (89, 32) and tuple(conv.bias.shape) == (32,)
    assert set(conv.state_dict()) == {"weight", "bias"}
    assert float(conv.bias.abs().max()) == 0.0
    a = (6.0 / (89 + 32)) ** 0.5
    assert float(conv.weight.abs().max()) <= a + 1e-6                                    # glorot-uniform bound


def test_dataset_and_loader(shim):
    os.environ["TSG_SYNTH_GRAPHS"] = "12"
    os.environ["TSG_ALLOW_SYNTH"] = "1"
    try:
        from torch_geometric.data import DataLoader
        from torch_geometric.datasets import TUDataset
        ds = TUDataset("data/DD", name="DD")
        assert len(ds) == 12 and ds.num_classes == 2 and ds.num_features == 89
        tr, va = torch.utils.data.random_split(ds, [9, 3])
        batch = next(iter(DataLoader(tr, batch_size=4, shuffle=False)))
        assert batch.x.size(1) == 89 and batch.edge_index.dtype == torch.int64
        assert int(batch.batch.max()) == 3 and batch.y.numel() == 4
        assert int(batch.edge_index.max()) < batch.x.size(0)
    finally:
        del os.environ["TSG_SYNTH_GRAPHS"], os.environ["TSG_ALLOW_SYNTH"]


def test_dataset_missing_files_raise(shim, tmp_path):
    """a wrong data path must not silently train on synthetic graphs."""
    from torch_geometric.datasets import TUDataset
    os.environ.pop("TSG_ALLOW_SYNTH", None)
    with pytest.raises(FileNotFoundError):
        TUDataset(str(tmp_path), name="DD")
    os.environ["TSG_ALLOW_SYNTH"] = "1"
    try:
        with pytest.raises(FileNotFoundError):
            TUDataset(str(tmp_path), name="NOT_A_DATASET")
    finally:
        del os.environ["TSG_ALLOW_SYNTH"]


@pytest.mark.skipif(not os.path.isdir(REF_SAG), reason="reference checkout not present on this box")
def test_reference_modules_import_through_the_shim(shim):
    """The UNMODIFIED Code/sag/layers.py and network.py import and construct on top of the shim and
    produce the state-dict keys PackedSAGNet uses (so `latest.pth` round-trips)."""
    sys.path.insert(0, REF_SAG)
    try:
        for m in ("layers", "network"):
            sys.modules.pop(m, None)
        import network
        net = network.Net(89, 32, 32, 0.5, 0.5)
        from tsg.nn import PackedSAGNet
        mine = PackedSAGNet(89, 32, 32, 0.5, 0.5)
        assert list(net.state_dict().keys()) == list(mine.state_dict().keys())
        mine.load_state_dict(net.state_dict())
    finally:
        sys.path.remove(REF_SAG)
        for m in ("layers", "network"):
            sys.modules.pop(m, None)


@pytest.mark.gpu
def test_drop_in_ops_like_layers_py(cuda, shim):
    """Drive the shim exactly as SAGPool.forward (layers.py:14-26) and Net.forward (network.py:34-36)
    do -- edge_index / batch tensors in, PyG-shaped tensors out -- and compare with the oracle."""
    from conftest import rel_err
    from oracle import pyg_ref as R
    from tsg import synth
    from torch_geometric.nn import GCNConv, global_max_pool as gmp, global_mean_pool as gap
    from torch_geometric.nn.pool.topk_pool import filter_adj, topk
    c = synth.make_corpus("DD", 5, seed=3); b = synth.pack(c)
    x, ei, batch = torch.from_numpy(b["x"]), torch.from_numpy(b["edge_index"]), torch.from_numpy(b["batch"])
    torch.manual_seed(0)
    conv, score_layer = GCNConv(89, 32).to(cuda), GCNConv(32, 1).to(cuda)
    xg, eig, bg = x.to(cuda), ei.to(cuda), batch.to(cuda)
    h = torch.relu(conv(xg, eig))
    score = score_layer(h, eig).squeeze()
    perm = topk(score, 0.5, bg)
    xo = h[perm] * torch.tanh(score[perm]).view(-1, 1)
    bo = bg[perm]
    ei2, _ = filter_adj(eig, None, perm, num_nodes=score.size(0))
    out = torch.cat([gmp(xo, bo), gap(xo, bo)], dim=1)
    ho = torch.relu(R.gcn_conv(x, ei, conv.weight.detach().cpu(), conv.bias.detach().cpu()))
    so = R.gcn_conv(ho, ei, score_layer.weight.detach().cpu(), score_layer.bias.detach().cpu()).squeeze()
    po = R.topk(so, 0.5, batch)
    assert torch.equal(perm.cpu(), po)
    eo, _ = R.filter_adj(ei, None, po, x.size(0))
    assert torch.equal(ei2.cpu(), eo) and ei2.shape == eo.shape
    xr = ho[po] * torch.tanh(so[po]).view(-1, 1)
    ref = torch.cat([R.global_max_pool(xr, batch[po]), R.global_mean_pool(xr, batch[po])], dim=1)
    assert rel_err(out, ref) <= 1e-5
    out.sum().backward()
    assert conv.weight.grad is not None and score_layer.weight.grad is not None


@pytest.mark.gpu
def test_device_hygiene_shims_on_cuda(shim, cuda):
    """What Code/sag/train_triplet.py:104-130 does on a GPU box: CPU embeddings into a CUDA MLP, CPU target into
    cross_entropy, CUDA prediction compared with a CPU label, .numpy() on a CUDA tensor."""
    import torch.nn.functional as F
    m = torch.nn.Sequential(torch.nn.Linear(3, 2).to(cuda))
    out = m(torch.randn(3))
    assert out.is_cuda
    loss = F.cross_entropy(out.unsqueeze(0), torch.LongTensor([1]))
    loss.backward()
    assert out.argmax(dim=0).eq(torch.Tensor([1.0])).sum().item() in (0, 1)
    assert out.numpy().shape == (2,)
