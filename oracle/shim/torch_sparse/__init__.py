"""TEST INFRASTRUCTURE — placeholder for torch_sparse==0.6.18 (requirements.txt:16).
The reference only imports `SparseTensor, set_diag` for isinstance checks
(models/mgat_v2_conv.py:12,205,237); no arithmetic is restated."""


class SparseTensor:  # never instantiated on the hot path
    def __init__(self, *a, **k):
        raise NotImplementedError("torch_sparse shim: SparseTensor is import-only")


def set_diag(*a, **k):
    raise NotImplementedError("torch_sparse shim: set_diag is import-only")
