"""CPU: the switches that decide whether gradient producers may run on a side stream (isg_b200.ops) — the advisor's
round-1 finding was that the ban for hook-based reducers only held when a process group existed.  The ban is a
counter that `allow_side_stream` honours unconditionally; `OverlappedGradAllReduce` holds it for its lifetime, also at
world size 1 without any process group."""
import torch

from isg_b200 import ops
from isg_b200.dp import GradAllReduce, OverlappedGradAllReduce


def _reset():
    ops._side_forbidden = 0
    ops.set_side_stream_with_dist(False)
    ops.allow_side_stream(False)


def test_ban_is_a_counter_and_needs_no_process_group():
    _reset()
    assert not torch.distributed.is_initialized()
    ops.allow_side_stream(True)
    assert ops._side_ok is True
    ops.forbid_side_stream(True)
    ops.forbid_side_stream(True)
    ops.allow_side_stream(True)
    assert ops._side_ok is False, "a banned side stream must stay off whatever the caller vouches for"
    ops.forbid_side_stream(False)
    ops.allow_side_stream(True)
    assert ops._side_ok is False, "two consumers banned it, one is still alive"
    ops.forbid_side_stream(False)
    ops.allow_side_stream(True)
    assert ops._side_ok is True
    ops.forbid_side_stream(False)  # an unbalanced release never drives the counter negative
    assert ops._side_forbidden == 0
    _reset()


def test_hook_based_reducer_holds_the_ban_for_its_lifetime_at_world_size_one():
    _reset()
    model = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.GELU(), torch.nn.Linear(8, 4))
    flat = GradAllReduce(model)  # reads gradients after backward() has returned: no ban
    assert ops._side_forbidden == 0 and flat.world == 1
    red = OverlappedGradAllReduce(model, bucket_bytes=256)
    assert ops._side_forbidden == 1
    ops.allow_side_stream(True)
    assert ops._side_ok is False
    # the reducer still does its job at world size 1: gradients pass through the flat bucket unchanged
    x = torch.randn(5, 8)
    model(x).square().mean().backward()
    want = [p.grad.clone() for p in model.parameters()]
    red.finish()
    for p, w in zip(model.parameters(), want):
        assert torch.equal(p.grad, w)
    red.remove_hooks()
    assert ops._side_forbidden == 0
    red.remove_hooks()  # idempotent
    assert ops._side_forbidden == 0
    ops.allow_side_stream(True)
    assert ops._side_ok is True
    _reset()
