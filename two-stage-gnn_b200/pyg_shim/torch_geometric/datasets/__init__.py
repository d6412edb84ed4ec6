"""torch_geometric.datasets.TUDataset (train*.py:5,161-163).  <root>/<name>/[raw/]<name>_A.txt is read by the native
TU loader (tsg.tu, H1).  Missing files raise FileNotFoundError (upstream PyG would download them; there is no network
here and training on made-up data must never happen silently).  Only with TSG_ALLOW_SYNTH=1 in the environment are
seeded synthetic graphs of the named dataset's SHAPE generated instead (tsg.synth; DD and PROTEINS shapes only), with
a warning on stderr; TSG_SYNTH_GRAPHS then overrides the graph count.  Node features are one-hot node labels."""
import os
import sys

import torch

from tsg import synth
from ..data import Data

_ALIAS = {"DD": "DD", "PROTEINS": "PROTEINS"}
_COUNT = {"DD": 1168, "PROTEINS": 1113}


class TUDataset(torch.utils.data.Dataset):
    def __init__(self, root, name, **kw):
        from tsg import tu
        prefix = tu.find_prefix(root, name) if root else None
        if prefix is not None:
            # real TU files on disk: native one-pass loader (H1), PyG reader semantics
            self.name = name
            self.corpus, self.attr, self.num_classes = tu.load(prefix, "pyg")
            self.num_features = self.corpus.num_node_labels
        else:
            if os.environ.get("TSG_ALLOW_SYNTH", "0") != "1":
                raise FileNotFoundError(
                    f"TUDataset: no TU files for {name!r} under {root!r} (looked for {name}_A.txt in <root>, "
                    f"<root>/{name}, <root>/{name}/raw); there is no download path.  Set TSG_ALLOW_SYNTH=1 to train on "
                    "seeded SYNTHETIC graphs of the dataset's shape instead.")
            if name not in _ALIAS:
                raise FileNotFoundError(f"TUDataset: no synthetic shape is defined for {name!r} (known: {sorted(_ALIAS)})")
            shape = _ALIAS[name]
            print(f"[tsg] WARNING: TUDataset({root!r}, {name!r}): files not found, TSG_ALLOW_SYNTH=1 -> using SYNTHETIC "
                  f"{shape}-shape graphs (seed 777); accuracies are meaningless", file=sys.stderr)
            n = int(os.environ.get("TSG_SYNTH_GRAPHS", _COUNT.get(shape, 1168)))
            self.name, self.corpus = name, synth.make_corpus(shape, n, seed=777)
            self.num_classes, self.num_features = 2, self.corpus.num_node_labels
        self.num_node_features = self.num_features

    def __len__(self):
        return self.corpus.num_graphs

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        pk = synth.pack(self.corpus, [int(i)])
        return Data(torch.from_numpy(pk["x"]), torch.from_numpy(pk["edge_index"]), torch.from_numpy(pk["y"]))
