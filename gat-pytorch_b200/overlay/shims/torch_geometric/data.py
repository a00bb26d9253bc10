"""`Data` (attribute bag with `.to`) and `DataLoader` (PyG batching: node tensors concatenated, `edge_index` offset per
graph, `batch` = graph id per node) -- what GATModel.py:262-279 and the `*_step` methods consume."""
import torch


class Data:
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None]

    @property
    def num_nodes(self):
        return self.x.size(0)

    def to(self, device, **kwargs):
        return Data(**{k: (v.to(device, **kwargs) if torch.is_tensor(v) else v) for k, v in self.__dict__.items()})

    def __repr__(self):
        return "Data(" + ", ".join(f"{k}={list(v.shape) if torch.is_tensor(v) else v}" for k, v in self.__dict__.items()) + ")"


class Batch(Data):
    @staticmethod
    def from_data_list(graphs):
        out, offset, parts, batch = {}, 0, {}, []
        for gi, g in enumerate(graphs):
            for k in g.keys():
                v = getattr(g, k)
                if not torch.is_tensor(v):
                    continue
                parts.setdefault(k, []).append(v + offset if k == "edge_index" else v)
            batch.append(torch.full((g.num_nodes,), gi, dtype=torch.long))
            offset += g.num_nodes
        for k, vs in parts.items():
            out[k] = torch.cat(vs, dim=1 if k == "edge_index" else 0)
        out["batch"] = torch.cat(batch)
        out["num_graphs"] = len(graphs)
        return Batch(**out)


class DataLoader:
    def __init__(self, dataset, batch_size=1, shuffle=False, **kwargs):
        self.dataset, self.batch_size, self.shuffle = dataset, max(int(batch_size), 1), shuffle

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        order = torch.randperm(len(self.dataset)).tolist() if self.shuffle else list(range(len(self.dataset)))
        for i in range(0, len(order), self.batch_size):
            yield Batch.from_data_list([self.dataset[j] for j in order[i:i + self.batch_size]])
