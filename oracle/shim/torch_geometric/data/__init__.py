class Data:
    """Attribute bag (torch_geometric.data.Data is import-only on the hot path:
    sampling/methods/tensor_utils.py:4)."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)
