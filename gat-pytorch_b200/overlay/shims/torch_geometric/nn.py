"""`GATConv` is imported by GATModel.py:10 and only constructed for LayerType.PyTorch_Geometric, which no run_config entry
selects; the stand-in exists for the import."""


class GATConv:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("torch_geometric.nn.GATConv is not part of the offline stand-in (LayerType.GATLayer is the path)")
