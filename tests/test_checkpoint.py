"""Drop-in surface on the host: reference-layout checkpoints (strict=True, with and without the DDP `module.`
prefix, training/train_loop.py:84-130, main.py:125-139, run_token_coo.py:43) and isubgvqa.install().  CPU only —
nothing here launches a kernel."""
import importlib
import os
import sys

import pytest
import torch

import reference_loader as rl
from isg_b200 import checkpoint, synth
from isg_b200 import isubgvqa as drop


def _mgat(sampler="aimle", C=16):
    return drop.MGAT(channels=C, num_ins=4, heads=4, use_instr=True, masking_thresholds=[1.0, 1.0, 1.0, 0.1],
                     use_topk=True, interpretable_mode=False, sampler_type=sampler, sample_k=2)


class _Host(torch.nn.Module):
    """Stands in for ISubGVQA (models/isubgvqa.py:159: `self.gat_seq = MGAT(...)`)."""

    def __init__(self, sampler="aimle"):
        super().__init__()
        self.gat_seq = _mgat(sampler)
        self.logit_fc = torch.nn.Linear(16, 5)


class _DDPLike(torch.nn.Module):
    def __init__(self, inner):
        super().__init__()
        self.module = inner


def test_reference_layout_checkpoint_loads_strict(tmp_path):
    ref_sd = synth.make_state_dict(16, 4, 4, seed=3)  # exactly the reference MGAT's keys and shapes
    host = _Host()
    full = {"module.gat_seq." + k: v for k, v in ref_sd.items()}
    full.update({"module.logit_fc.weight": torch.randn(5, 16), "module.logit_fc.bias": torch.randn(5)})
    path = tmp_path / "checkpoint.pth"
    torch.save({"model": full, "epoch": 7, "args": {"sampler_type": "aimle"}}, path)
    # (1) into a DDP-wrapped model: keys match verbatim, strict
    ddp = _DDPLike(_Host())
    ddp.load_state_dict(torch.load(path, weights_only=False)["model"], strict=True)
    for k, v in ref_sd.items():
        assert torch.equal(ddp.module.gat_seq.state_dict()[k], v), k
    # (2) into a bare model through the helper (prefix stripped), strict
    ck = checkpoint.load(path, host, strict=True)
    assert ck["epoch"] == 7
    for k, v in ref_sd.items():
        assert torch.equal(host.gat_seq.state_dict()[k], v), k
    # (3) a wrong key still fails loudly
    bad = dict(full)
    bad["module.gat_seq.convs.0.lin_q.weight"] = torch.zeros(1)
    with pytest.raises(RuntimeError):
        ddp.load_state_dict(bad, strict=True)


def test_checkpoint_round_trip_keeps_aimle_state_and_exports_reference_keys(tmp_path):
    host = _DDPLike(_Host("aimle"))
    tgt = host.module.gat_seq.convs[3].mask.sampler_train.target
    tgt.beta = 0.0123
    tgt.grad_norm = 0.77
    opt = torch.optim.Adam(host.parameters(), lr=1e-3)
    path = tmp_path / "ck.pth"
    checkpoint.save(path, host, optimizer=opt, epoch=2)
    again = _DDPLike(_Host("aimle"))
    checkpoint.load(path, again, optimizer=torch.optim.Adam(again.parameters(), lr=1e-3), strict=True)
    t2 = again.module.gat_seq.convs[3].mask.sampler_train.target
    assert t2.beta == 0.0123 and t2.grad_norm == 0.77
    for (k, a), (_, b) in zip(host.state_dict().items(), again.state_dict().items()):
        assert torch.equal(a, b), k
    # export for the unmodified reference: exactly the reference's key set, nothing else
    exported = checkpoint.sub_state(checkpoint.strip_ddp_prefix(checkpoint.strip_isg_keys(host.state_dict())),
                                    "gat_seq.")
    assert set(exported) == set(synth.mgat_param_shapes(16, 4, 4))
    assert all(tuple(exported[k].shape) == s for k, s in synth.mgat_param_shapes(16, 4, 4).items())


@pytest.mark.skipif(not rl.available(), reason="needs /root/reference (authoring container)")
@pytest.mark.parametrize("sampler", ["imle", "aimle", "gumbel"])
def test_state_dicts_interchange_with_the_live_reference(sampler):
    rl.load()
    from ISubGVQA.models.mgat import MGAT as RefMGAT

    with rl.scratch_cwd():
        ref = RefMGAT(channels=16, num_ins=4, heads=4, use_instr=True, masking_thresholds=[1.0, 1.0, 1.0, 0.1],
                      use_topk=True, interpretable_mode=False, sampler_type=sampler, sample_k=2, nb_samples=1,
                      alpha=1.0, beta=10.0, tau=1.0)
    ours = _mgat(sampler)
    ours.load_state_dict(ref.state_dict(), strict=True)              # reference -> isg_b200
    ref.load_state_dict(checkpoint.strip_isg_keys(ours.state_dict()), strict=True)  # isg_b200 -> reference
    assert list(checkpoint.strip_isg_keys(ours.state_dict()).keys()) == list(ref.state_dict().keys())


def test_install_routes_reference_imports(tmp_path):
    """isubgvqa.install() makes `from ISubGVQA.models.mgat import MGAT` (models/isubgvqa.py:10) and the sampler
    imports of models/masking.py resolve to isg_b200 while the rest of the ISubGVQA package stays the
    reference's.  A skeleton package with empty __init__ files stands in for the reference tree."""
    for pkg in ("ISubGVQA", "ISubGVQA/models", "ISubGVQA/sampling", "ISubGVQA/sampling/methods", "ISubGVQA/utils"):
        os.makedirs(tmp_path / pkg, exist_ok=True)
        (tmp_path / pkg / "__init__.py").write_text("")
    (tmp_path / "ISubGVQA" / "utils" / "marker.py").write_text("WHO = 'reference'\n")
    saved = {k: v for k, v in sys.modules.items() if k == "ISubGVQA" or k.startswith("ISubGVQA.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, str(tmp_path))
    try:
        drop.install()
        from ISubGVQA.models.mgat import MGAT
        from ISubGVQA.models.mgat_v2_conv import MaskingGATv2Conv
        from ISubGVQA.models.masking import MaskingModel, get_aimle_samplers, get_imle_samplers  # noqa: F401
        from ISubGVQA.models.att_pooling import GlobalAttention
        from ISubGVQA.sampling.node_edge_masks import NodeMaskToEdgeMask
        from ISubGVQA.sampling.methods.wrapper import imle
        from ISubGVQA.sampling.methods.aimle import aimle
        from ISubGVQA.sampling.methods.gumbel_scheme import GumbelSampler
        from ISubGVQA.sampling.methods.simple_scheme import EdgeSIMPLEBatched
        from ISubGVQA.sampling.methods.noise import GumbelDistribution  # noqa: F401
        from ISubGVQA.sampling.methods.target import TargetDistribution  # noqa: F401
        from ISubGVQA.sampling.methods.target_aimle import AdaptiveTargetDistribution  # noqa: F401
        from ISubGVQA.sampling.methods.imle_scheme import IMLEScheme  # noqa: F401
        from ISubGVQA.sampling.methods.deterministic_scheme import select_from_edge_candidates  # noqa: F401

        assert MGAT is drop.MGAT and MaskingGATv2Conv is drop.MaskingGATv2Conv and MaskingModel is drop.MaskingModel
        assert GlobalAttention is drop.GlobalAttention and NodeMaskToEdgeMask is drop.NodeMaskToEdgeMask
        assert imle is drop.samplers.imle and aimle is drop.samplers.aimle
        assert GumbelSampler is drop.samplers.GumbelSampler and EdgeSIMPLEBatched is drop.samplers.EdgeSIMPLEBatched
        assert importlib.import_module("ISubGVQA.utils.marker").WHO == "reference"  # untouched modules stay
        m = MGAT(channels=16, num_ins=4, heads=4, use_instr=True, masking_thresholds=[1.0, 1.0, 1.0, 0.1],
                 use_topk=True, interpretable_mode=False, sampler_type="imle", sample_k=2)
        assert set(checkpoint.strip_isg_keys(m.state_dict())) == set(synth.mgat_param_shapes(16, 4, 4))
        drop.uninstall()
        assert "ISubGVQA.models.mgat" not in sys.modules
    finally:
        drop.uninstall()
        sys.path.remove(str(tmp_path))
        for k in [k for k in sys.modules if k == "ISubGVQA" or k.startswith("ISubGVQA.")]:
            del sys.modules[k]
        sys.modules.update(saved)
