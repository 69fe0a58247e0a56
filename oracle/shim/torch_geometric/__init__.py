"""TEST INFRASTRUCTURE — pure-torch restatement of the torch_geometric==2.6.1 symbols the
reference hot path imports (requirements.txt:14; SURVEY.md §8c lists every call site).
It exists so `/root/reference/ISubGVQA/models/{mgat,mgat_v2_conv,masking}.py` and
`ISubGVQA/sampling/**` import and run UNMODIFIED on CPU in the authoring container.
Never imported by the product package."""
from . import nn, utils, typing, data  # noqa: F401

__version__ = "2.6.1+isg-shim"
