"""CPU: the oracle restatement (oracle/isg_oracle.py) must reproduce the committed golden vectors,
which were produced by the UNMODIFIED reference (oracle/make_golden.py)."""
import pytest
import torch

import util


@pytest.mark.parametrize("path", util.golden_files(), ids=lambda p: p.split("/")[-1][:-3])
def test_oracle_matches_reference_golden(path):
    torch.set_num_threads(max(1, torch.get_num_threads()))
    fix = util.load_golden(path)
    cfg = fix["config"]
    outs = util.run_oracle_case(cfg)
    assert len(outs) == len(fix["steps"])
    for got, want in zip(outs, fix["steps"]):
        util.compare_step(got, want, cfg["sampler"])


def test_golden_set_covers_samplers_and_modes():
    cfgs = [util.load_golden(p)["config"] for p in util.golden_files()]
    assert {c["sampler"] for c in cfgs} >= {"imle", "aimle", "gumbel"}
    assert {c["train"] for c in cfgs} == {True, False}
