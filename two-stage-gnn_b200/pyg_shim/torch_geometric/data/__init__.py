"""torch_geometric.data.{Data, Batch, DataLoader} (train*.py:6; Batch layout SURVEY A.1.5)."""
import torch


class Data:
    def __init__(self, x=None, edge_index=None, y=None, **kw):
        self.x, self.edge_index, self.y = x, edge_index, y
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        return self.x.size(0)

    @property
    def num_graphs(self):
        return 1

    def keys(self):
        return [k for k, v in self.__dict__.items() if torch.is_tensor(v)]

    def to(self, device, **kw):
        out = self.__class__.__new__(self.__class__)
        out.__dict__ = {k: (v.to(device, **kw) if torch.is_tensor(v) else v) for k, v in self.__dict__.items()}
        return out

    def __repr__(self):
        return f"Data(x={tuple(self.x.shape)}, edge_index={tuple(self.edge_index.shape)}, y={self.y})"


class Batch(Data):
    @staticmethod
    def from_data_list(data_list):
        xs, eis, ys, batch, off = [], [], [], [], 0
        for g, d in enumerate(data_list):
            xs.append(d.x); eis.append(d.edge_index + off); ys.append(d.y.view(-1))
            batch.append(torch.full((d.x.size(0),), g, dtype=torch.long, device=d.x.device))
            off += d.x.size(0)
        b = Batch(torch.cat(xs, 0), torch.cat(eis, 1), torch.cat(ys, 0))
        b.batch = torch.cat(batch, 0)
        b._num_graphs = len(data_list)
        return b

    @property
    def num_graphs(self):
        return self._num_graphs


class DataLoader(torch.utils.data.DataLoader):
    def __init__(self, dataset, batch_size=1, shuffle=False, **kw):
        kw.pop("collate_fn", None)
        super().__init__(dataset, batch_size, shuffle, collate_fn=Batch.from_data_list, **kw)
