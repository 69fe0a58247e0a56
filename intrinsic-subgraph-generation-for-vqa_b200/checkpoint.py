"""Checkpoint compatibility with the reference (training/train_loop.py:84-130 writes {"model":
model.state_dict(), "optimizer": ..., "lr_scheduler": ..., "epoch": ..., "args": ...}; main.py:125-139 and
run_token_coo.py:43 load it back, the latter with strict=True; under DistributedDataParallel every key carries
a `module.` prefix).

The drop-in modules use the reference's parameter names, so a reference checkpoint loads into a model whose hot
path is isg_b200 without translation.  The only addition is the AIMLE adaptive state (`...mask.aimle_state`,
float64[8]) which the reference keeps in Python attributes and loses on resume (target_aimle.py:101-109): it is
optional on load (MaskingModel.OPTIONAL_STATE_KEYS) and `strip_isg_keys` removes it when a checkpoint has to be
read by the unmodified reference."""
import torch

ISG_ONLY_SUFFIXES = (".aimle_state",)
DDP_PREFIX = "module."


def strip_isg_keys(state_dict):
    """state_dict without the keys only isg_b200 knows (same object types, insertion order preserved)."""
    return type(state_dict)((k, v) for k, v in state_dict.items() if not k.endswith(ISG_ONLY_SUFFIXES))


def strip_ddp_prefix(state_dict):
    """`module.x.y` -> `x.y` (a DistributedDataParallel checkpoint loaded into a bare model)."""
    return type(state_dict)((k[len(DDP_PREFIX):] if k.startswith(DDP_PREFIX) else k, v)
                            for k, v in state_dict.items())


def add_ddp_prefix(state_dict):
    return type(state_dict)((k if k.startswith(DDP_PREFIX) else DDP_PREFIX + k, v) for k, v in state_dict.items())


def sub_state(state_dict, prefix):
    """The entries below `prefix` (e.g. 'gat_seq.' — ISubGVQA.gat_seq is the MGAT, models/isubgvqa.py:159),
    with the prefix removed."""
    return type(state_dict)((k[len(prefix):], v) for k, v in state_dict.items() if k.startswith(prefix))


def save(path, model, optimizer=None, lr_scheduler=None, epoch=None, args=None, for_reference=False):
    """Writes the reference's checkpoint layout (training/train_loop.py:89-99)."""
    sd = model.state_dict()
    ckpt = {"model": strip_isg_keys(sd) if for_reference else sd}
    if optimizer is not None:
        ckpt["optimizer"] = optimizer.state_dict()
    if lr_scheduler is not None:
        ckpt["lr_scheduler"] = lr_scheduler.state_dict()
    if epoch is not None:
        ckpt["epoch"] = epoch
    if args is not None:
        ckpt["args"] = args
    torch.save(ckpt, path)
    return ckpt


def load(path, model, optimizer=None, lr_scheduler=None, strict=True, map_location="cpu"):
    """main.py:125-139.  Accepts checkpoints written with or without the DDP `module.` prefix, whichever the
    receiving model uses."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    sd = ckpt["model"]
    want_prefix = any(k.startswith(DDP_PREFIX) for k in model.state_dict().keys())
    have_prefix = any(k.startswith(DDP_PREFIX) for k in sd.keys())
    if have_prefix and not want_prefix:
        sd = strip_ddp_prefix(sd)
    elif want_prefix and not have_prefix:
        sd = add_ddp_prefix(sd)
    model.load_state_dict(sd, strict=strict)
    if optimizer is not None and "optimizer" in ckpt:
        optimizer.load_state_dict(ckpt["optimizer"])
    if lr_scheduler is not None and "lr_scheduler" in ckpt:
        lr_scheduler.load_state_dict(ckpt["lr_scheduler"])
    return ckpt
