"""Host-side mirror of the reference interface for the hot path (ISubGVQA/models/{mgat,mgat_v2_conv,
masking}.py, ISubGVQA/sampling/**): same names, arguments, return shapes and state_dict keys, with the
arithmetic done by libisg.so.  `install()` makes the reference's own import statements resolve to these
modules, so main.py / training / eval run unchanged."""
import sys

from . import att_pooling, masking, mgat, mgat_v2_conv, node_edge_masks, samplers, scene_graph_encoder  # noqa: F401
from .att_pooling import GlobalAttention  # noqa: F401
from .masking import MaskingModel, get_aimle_samplers, get_imle_samplers  # noqa: F401
from .mgat import MGAT  # noqa: F401
from .mgat_v2_conv import MaskingGATv2Conv  # noqa: F401
from .node_edge_masks import NodeMaskToEdgeMask  # noqa: F401
from .scene_graph_encoder import GraphNorm64, SceneGraphEncodingLayer, encode_scene_graph  # noqa: F401

_ALIASES = {
    "ISubGVQA.models.mgat": mgat,
    "ISubGVQA.models.mgat_v2_conv": mgat_v2_conv,
    "ISubGVQA.models.masking": masking,
    "ISubGVQA.models.att_pooling": att_pooling,
    "ISubGVQA.sampling.node_edge_masks": node_edge_masks,
    "ISubGVQA.sampling.methods.wrapper": samplers,
    "ISubGVQA.sampling.methods.aimle": samplers,
    "ISubGVQA.sampling.methods.noise": samplers,
    "ISubGVQA.sampling.methods.target": samplers,
    "ISubGVQA.sampling.methods.target_aimle": samplers,
    "ISubGVQA.sampling.methods.imle_scheme": samplers,
    "ISubGVQA.sampling.methods.deterministic_scheme": samplers,
    "ISubGVQA.sampling.methods.gumbel_scheme": samplers,
    "ISubGVQA.sampling.methods.simple_scheme": samplers,
}


def install():
    """Route the reference's hot-path imports to isg_b200 (call before `import ISubGVQA.models...`)."""
    for name, mod in _ALIASES.items():
        sys.modules[name] = mod


def uninstall():
    for name, mod in _ALIASES.items():
        if sys.modules.get(name) is mod:
            del sys.modules[name]
