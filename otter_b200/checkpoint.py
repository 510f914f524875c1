"""Trainable-only checkpoint format of the reference's training loop.

  * `get_checkpoint(model)`        pipeline/train/train_utils.py:60-67 — state_dict minus every parameter that does not
    require grad (the frozen CLIP tower and LM never hit the disk: 1.2 B of 8 B parameters at the 7B shape);
  * `save_final_weights(...)`      train_utils.py:234-262 (the non-HF branch): `<dir>/final_weights.pt` + the config;
  * `load_trainable_checkpoint`    the `--trained_ckpt` path of instruction_following.py (`load_state_dict(strict=False)`),
    with the checks the reference leaves to the user: unknown keys and shape mismatches raise.
Keys are the reference's (tests/golden/state_dict_keys.json), so files written by either side load on the other.
"""
import os

import torch


def get_checkpoint(model):
    state_dict = model.state_dict()
    for name, p in model.named_parameters():
        if not p.requires_grad:
            del state_dict[name]
    return state_dict


def save_final_weights(model, external_save_dir, is_main_process=True, save_function=torch.save):
    os.makedirs(external_save_dir, exist_ok=True)
    if hasattr(model, "config") and hasattr(model.config, "save_pretrained") and is_main_process:
        model.config.save_pretrained(external_save_dir)
    checkpoint_dict = {k: v.detach().to("cpu") for k, v in get_checkpoint(model).items()}
    path = os.path.join(external_save_dir, "final_weights.pt")
    if is_main_process:
        save_function(checkpoint_dict, path)
    return path


def load_trainable_checkpoint(model, path, strict_shapes=True):
    sd = torch.load(path, map_location="cpu", weights_only=True)
    if "model_state_dict" in sd:                      # instruction_following.py accepts both layouts
        sd = sd["model_state_dict"]
    own = model.state_dict()
    unknown = [k for k in sd if k not in own]
    if unknown:
        raise KeyError(f"checkpoint has keys the model does not: {unknown[:5]}{' ...' if len(unknown) > 5 else ''}")
    if strict_shapes:
        bad = [(k, tuple(v.shape), tuple(own[k].shape)) for k, v in sd.items() if tuple(v.shape) != tuple(own[k].shape)]
        if bad:
            raise ValueError(f"shape mismatch: {bad[:3]}")
    missing = model.load_state_dict(sd, strict=False)
    from . import params
    params.clear_caches()                             # the bf16 compute copies must be re-derived from the new weights
    return missing
