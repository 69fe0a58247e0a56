def _unavailable(*_a, **_k):
    raise RuntimeError("torchtext shim: GloVe / vocab are not available offline (not needed by the tested conversion)")


GloVe = vocab = _unavailable
