"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference hot-path modules from
/root/reference on top of the pure-torch shim in oracle/shim (authoring container only;
/root/reference does not exist on the GPU box, so nothing under tests/ -m gpu, smoke()
or bench.py may call this).  Used by oracle/make_golden.py and tests/test_oracle_vs_reference.py
to pin oracle/isg_oracle.py against the real reference code.

Patches applied to the *harness*, not to the reference (SURVEY.md §8c last row):
  * GumbelDistribution.sample (sampling/methods/noise.py:86-89): device="cuda" is hard-coded
    at models/masking.py:97,106,114 -> keep samples on CPU and allow noise injection.
  * TORCHDYNAMO_DISABLE=1 for sampling/methods/simple.py (Inductor CPU compile is broken here).
  * simple.py:114-120 writes ./simple_configs/ into CWD -> callers chdir to a scratch dir.
"""
import contextlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "shim")


def _find_root():
    """/root/reference in the authoring container; on the GPU box the unmodified hot-path files staged by
    oracle/stage_reference.py into the git-ignored oracle/_ref/ (they travel with the gpurun snapshot)."""
    env = os.environ.get("ISG_REFERENCE_ROOT")
    for cand in ([env] if env else []) + ["/root/reference", os.path.join(_HERE, "_ref")]:
        if os.path.isdir(os.path.join(cand, "ISubGVQA", "models")):
            return cand
    return env or "/root/reference"


REFERENCE_ROOT = _find_root()

_injected_noise = []  # FIFO of tensors consumed by the patched GumbelDistribution.sample


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ISubGVQA", "models"))


def load():
    """Returns the reference `ISubGVQA` package (models.mgat etc. importable afterwards)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
    for p in (_SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ISubGVQA.sampling.methods.noise as noise_mod

    if not getattr(noise_mod.GumbelDistribution, "_isg_patched", False):
        orig = noise_mod.GumbelDistribution.sample

        def sample(self, shape):
            if _injected_noise:
                t = _injected_noise.pop(0)
                assert tuple(t.shape) == tuple(shape), (t.shape, shape)
                return t.clone()
            dev, self.device = self.device, "cpu"
            try:
                return orig(self, shape)
            finally:
                self.device = dev

        noise_mod.GumbelDistribution.sample = sample
        noise_mod.GumbelDistribution._isg_patched = True
    import ISubGVQA.models.mgat  # noqa: F401  (pulls mgat_v2_conv, masking, sampling/**)
    import ISubGVQA

    return ISubGVQA


@contextlib.contextmanager
def inject_noise(*tensors):
    """Queue noise tensors ([B, S, Nmax, 1]) returned by successive GumbelDistribution.sample calls."""
    _injected_noise.extend(tensors)
    try:
        yield
    finally:
        del _injected_noise[:]


@contextlib.contextmanager
def scratch_cwd(path="/tmp/isg_oracle_scratch"):
    os.makedirs(path, exist_ok=True)
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


@contextlib.contextmanager
def inject_theta_dropout(mask):
    """Replace F.dropout (models/masking.py:159, p=0.2 on theta [N,1]) by multiplication with a
    fixed Bernoulli keep-mask (values 0 or 1/0.8); mask=None -> identity."""
    import torch.nn.functional as F

    orig = F.dropout

    def fake(x, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return x
        if mask is None:
            return x
        assert x.shape == mask.shape, (x.shape, mask.shape)
        return x * mask

    F.dropout = fake
    try:
        yield
    finally:
        F.dropout = orig


@contextlib.contextmanager
def inject_device_gumbel(g):
    """GumbelSampler draws Gumbel(0,1) on-device (sampling/methods/gumbel_scheme.py:65-70) and
    SIMPLE draws uniform keys in simple.py:91-96; both are replaced by the fixed tensor `g`
    (gumbel: [B,Nmax]; simple: [B,n_pad])."""
    import torch
    import ISubGVQA.sampling.methods.simple as simple_mod

    orig_sample = torch.distributions.gumbel.Gumbel.sample
    orig_keys = simple_mod.gumbel_keys

    def fake_sample(self, sample_shape=torch.Size()):
        assert tuple(self.loc.shape) == tuple(g.shape), (self.loc.shape, g.shape)
        return g.clone()

    def fake_keys(w, time_sampled):
        assert time_sampled == 1
        return (w + g)[None]

    torch.distributions.gumbel.Gumbel.sample = fake_sample
    simple_mod.gumbel_keys = fake_keys
    try:
        yield
    finally:
        torch.distributions.gumbel.Gumbel.sample = orig_sample
        simple_mod.gumbel_keys = orig_keys


def load_scene_graph_encoding_layer(num_node_features=300, num_edge_features=300, hidden_dim=300):
    """The reference's MetaLayer (models/scene_graph_encoder.py:107-146), built by its own factory function.
    Harness patch (not an edit of the reference): models/scene_graph_encoder.py:5 imports
    ..datasets.scene_graph.GQASceneGraphs at module level, which needs torchtext / GloVe / the GQA JSON files;
    a stub module is registered under that name first — the factory function never touches it."""
    import types

    load()
    name = "ISubGVQA.datasets.scene_graph"
    if name not in sys.modules:
        pkg = sys.modules.get("ISubGVQA.datasets")
        if pkg is None:
            pkg = types.ModuleType("ISubGVQA.datasets")
            pkg.__path__ = []
            sys.modules["ISubGVQA.datasets"] = pkg
        stub = types.ModuleType(name)
        stub.GQASceneGraphs = type("GQASceneGraphs", (), {})
        sys.modules[name] = stub
    from ISubGVQA.models.scene_graph_encoder import get_gt_scene_graph_encoding_layer

    return get_gt_scene_graph_encoding_layer(num_node_features, num_edge_features, hidden_dim)
