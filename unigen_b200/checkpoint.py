"""Weight interchange for the native denoisers (SURVEY.md §8f rank 3, §8b "Weight interchange"): everything the reference's
`infer.py:124-141` accepts as `--transformer`, plus the diffusers `transformer/` folder `from_pretrained` reads
(`infer.py:115-119`) and the per-module files `save_all_model_hook` writes (`src/hook.py:10-27`).

  read_state_dict(path)          one {reference key: tensor} dict from
      * a DeepSpeed ZeRO checkpoint directory (has a `latest` tag file, infer.py:125-128): the consolidated
        `<tag>/pytorch_model_fp32.bin` that `script/infer.sh:44-46` produces when present, else the rank shards are
        consolidated here (ZeRO stage 2 and 3, trainable + frozen parameters) — a restatement of
        deepspeed 0.16.5 `utils/zero_to_fp32.py::get_fp32_state_dict_from_zero_checkpoint` ("parity unpinned": the
        package is absent here; tests/test_checkpoint_cpu.py round-trips shards written in the published layout);
      * a `torch.save` file (infer.py:133-135), incl. the `<module>_weights_<idx>.bin` family of src/hook.py;
      * a directory of `*.safetensors` files (infer.py:136-140), incl. HF sharded checkpoints with an index json.
  load_pretrained(model, base, control)   base diffusers weights (strict on the base keys) then the control-branch
                                           checkpoint with `strict=False`, returning the reference-style load result.
  save_modules(model, out_dir, modules)    src/hook.py:10-27 (`save_all_model_hook`) for the trainable control modules.

Host-side I/O only: tensors are read on the CPU and copied into the model's fused device storage by
`model.load_state_dict` (views under the reference's key names, no re-layout pass)."""
from __future__ import annotations

import glob
import json
import math
import os
import pickle
import re
import types
from collections import OrderedDict
from typing import Dict, Iterable, List, Optional, Sequence

import torch

Tensor = torch.Tensor

ZERO_STAGE = "zero_stage"
PARTITION_COUNT = "partition_count"
FP32_FLAT_GROUPS = "fp32_flat_groups"
SINGLE_PARTITION = "single_partition_of_fp32_groups"
PARAM_SHAPES = "param_shapes"
FROZEN_PARAM_SHAPES = "frozen_param_shapes"
FROZEN_PARAM_FRAGMENTS = "frozen_param_fragments"
BUFFER_NAMES = "buffer_names"
SHARED_PARAMS = "shared_params"


def _natural(s: str):
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", s)]


def _zero3_partition(numel: int, world: int):
    """zero_to_fp32.zero3_partitioned_param_info: every parameter is padded to a multiple of the world size."""
    rem = numel % world
    return math.ceil(numel / world), (world - rem) if rem else 0


def consolidate_zero_checkpoint(ckpt_dir: str, tag: Optional[str] = None, trust_pickle: bool = False) -> Dict[str, Tensor]:
    """fp32 state dict from the rank shards of a DeepSpeed ZeRO-2 / ZeRO-3 checkpoint.
    Layout read: `<dir>/<tag>/*_optim_states.pt` (optimizer_state_dict: zero_stage, partition_count, fp32 flat groups) and
    `<dir>/<tag>/*_model_states.pt` (module buffers, param_shapes per optimizer group, frozen_param_shapes / fragments,
    shared_params)."""
    if tag is None:
        with open(os.path.join(ckpt_dir, "latest")) as f:
            tag = f.read().strip()
    d = os.path.join(ckpt_dir, tag)
    optim_files = sorted(glob.glob(os.path.join(d, "*_optim_states.pt")), key=_natural)
    if not optim_files:
        raise FileNotFoundError(f"no *_optim_states.pt under {d}")
    optim = [_torch_load(f, trust_pickle)["optimizer_state_dict"] for f in optim_files]
    stage = optim[0][ZERO_STAGE]
    world = optim[0][PARTITION_COUNT]
    world = max(world) if isinstance(world, (list, tuple)) else world
    if world != len(optim_files):
        raise ValueError(f"expected {world} optimizer shards under {d}, found {len(optim_files)}")
    key = SINGLE_PARTITION if stage <= 2 else FP32_FLAT_GROUPS
    if stage not in (1, 2, 3):
        raise ValueError(f"unknown zero stage {stage}")
    flat = [o[key] for o in optim]
    model_files = sorted(glob.glob(os.path.join(d, "*_model_states.pt")), key=_natural)
    if not model_files:
        raise FileNotFoundError(f"no *_model_states.pt under {d}")
    states = [_torch_load(f, trust_pickle) for f in model_files]
    s0 = states[0]
    out: "OrderedDict[str, Tensor]" = OrderedDict()
    for name in s0.get(BUFFER_NAMES, []):
        out[name] = s0["module"][name].float()
    param_shapes: Sequence[Dict[str, torch.Size]] = s0[PARAM_SHAPES]
    frozen_shapes = s0.get(FROZEN_PARAM_SHAPES) or {}
    if stage <= 2:
        # frozen parameters are stored whole on rank 0; each optimizer group is ONE flat buffer split evenly over the ranks
        for name in frozen_shapes:
            out[name] = s0[FROZEN_PARAM_FRAGMENTS][name].float()
        for g, shapes in enumerate(param_shapes):
            full = torch.cat([flat[r][g] for r in range(world)], 0)
            off = 0
            for name, shape in shapes.items():
                n = math.prod(shape)
                out[name] = full.narrow(0, off, n).view(shape).clone()
                off += n
            if off > full.numel() or full.numel() - off >= 2 * world:  # stage-2 groups carry < 2 * world_size padding elements
                raise ValueError(f"group {g}: consumed {off} of {full.numel()} elements")
    else:
        if len(states) != world:
            raise ValueError(f"ZeRO-3 needs one model_states file per rank ({world}), found {len(states)}")
        for name, shape in frozen_shapes.items():
            n = math.prod(shape)
            frag = torch.cat([st[FROZEN_PARAM_FRAGMENTS][name].reshape(-1) for st in states], 0)
            out[name] = frag.narrow(0, 0, n).view(shape).float().clone()
        per_rank = [torch.cat([t.reshape(-1) for t in flat[r]], 0) for r in range(world)]
        off = 0
        for shapes in param_shapes:
            for name, shape in shapes.items():
                n = math.prod(shape)
                part, _pad = _zero3_partition(n, world)
                out[name] = torch.cat([per_rank[r].narrow(0, off, part) for r in range(world)], 0).narrow(0, 0, n).view(shape).clone()
                off += part
        if off != per_rank[0].numel():
            raise ValueError(f"consumed {off} of {per_rank[0].numel()} elements per rank")
    for pair in s0.get(SHARED_PARAMS, []) or []:
        if pair[1] in out:
            out[pair[0]] = out[pair[1]]
    return out


def _torch_load(path: str, trust_pickle: bool = False):
    """`torch.load(..., weights_only=True)`: tensors and plain containers only. DeepSpeed rank shards (`*_optim_states.pt`,
    `*_model_states.pt`) and older pickles carry python objects and need the full unpickler the reference itself uses
    (`torch.load(path)`, infer.py:133) — that executes code from the file, so it is taken ONLY with `trust_pickle=True`
    (an explicit opt-in of the caller), never as a silent retry; a truncated / corrupt file raises either way."""
    if trust_pickle:
        return torch.load(path, map_location="cpu", weights_only=False)
    try:
        return torch.load(path, map_location="cpu", weights_only=True)
    except pickle.UnpicklingError as e:
        raise pickle.UnpicklingError(
            f"{path} is not a plain tensor file ({e}); if it is a DeepSpeed shard or an older pickle from a source you trust, "
            "pass trust_pickle=True") from e


def _read_safetensors_dir(path: str, variant: Optional[str] = None) -> Dict[str, Tensor]:
    from safetensors.torch import load_file
    files = sorted(glob.glob(os.path.join(path, "*.safetensors")), key=_natural)
    if variant is not None:  # diffusers naming: diffusion_pytorch_model.<variant>.safetensors / ...<variant>-00001-of-0000N...
        files = [f for f in files if f".{variant}" in os.path.basename(f)]
        if not files:
            raise FileNotFoundError(f"no *.{variant}*.safetensors under {path}")
    index = glob.glob(os.path.join(path, "*.safetensors.index.json"))
    if index:  # HF sharded layout: only the shards the index names, each key from the shard it is mapped to
        with open(index[0]) as f:
            weight_map = json.load(f)["weight_map"]
        files = sorted({os.path.join(path, v) for v in weight_map.values()}, key=_natural)
    if not files:
        raise FileNotFoundError(f"no *.safetensors under {path}")
    out: Dict[str, Tensor] = {}
    for f in files:
        out.update(load_file(f))
    return out


def read_state_dict(path: str, trust_pickle: bool = False, variant: Optional[str] = None) -> Dict[str, Tensor]:
    """Branch order of infer.py:124-141. `trust_pickle` (default off) allows the full unpickler for files that are not plain
    tensor containers (DeepSpeed ZeRO rank shards always need it)."""
    if os.path.isdir(path) and os.path.exists(os.path.join(path, "latest")):
        with open(os.path.join(path, "latest")) as f:
            tag = f.read().strip()
        merged = os.path.join(path, tag, "pytorch_model_fp32.bin")  # script/infer.sh:44-46
        if os.path.exists(merged):
            return _torch_load(merged, trust_pickle)
        return consolidate_zero_checkpoint(path, tag, trust_pickle)
    if os.path.isfile(path):
        if path.endswith(".safetensors"):
            from safetensors.torch import load_file
            return load_file(path)
        sd = _torch_load(path, trust_pickle)
        return sd.get("state_dict", sd) if isinstance(sd, dict) else sd
    if os.path.isdir(path):
        bins = sorted(glob.glob(os.path.join(path, "*_weights_*.bin")), key=_natural)  # src/hook.py:19-25
        if bins and not glob.glob(os.path.join(path, "*.safetensors")):
            out: Dict[str, Tensor] = {}
            for f in bins:
                out.update(_torch_load(f, trust_pickle))
            return out
        return _read_safetensors_dir(path, variant)
    raise FileNotFoundError(path)


def load_pretrained(model, base: Optional[str] = None, control: Optional[str] = None, trust_pickle: bool = False):
    """`cls.from_pretrained(<base>/transformer)` + `load_state_dict(<control ckpt>, strict=False)` (infer.py:115-141).
    Base keys must all be present in the base checkpoint (a missing block would silently run on zeros otherwise); the
    control checkpoint is loaded non-strictly exactly like the reference and its load result is returned."""
    own = set(model.state_dict().keys())
    result = types.SimpleNamespace(missing_keys=[], unexpected_keys=[])
    if base is not None:
        sd = read_state_dict(base, trust_pickle)
        ctrl_prefixes = tuple(getattr(model, "trainable_control_modules", {}) or ())
        need = [k for k in own if not k.startswith(ctrl_prefixes)] if ctrl_prefixes else list(own)
        missing = [k for k in need if k not in sd]
        if missing:
            raise RuntimeError(f"base checkpoint {base} lacks {len(missing)} base-model keys, e.g. {missing[:4]}")
        model.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
    if control is not None:
        sd = read_state_dict(control, trust_pickle)
        result = model.load_state_dict(sd, strict=False)
    return result


def save_modules(model, out_dir: str, modules: Iterable[str], idx: int = 0) -> List[str]:
    """src/hook.py:10-27: one `<module>_weights_<idx>.bin` per trainable module, keys filtered by substring."""
    os.makedirs(out_dir, exist_ok=True)
    sd = model.state_dict()
    written = []
    for module in modules:
        part = {k: v.detach().cpu() for k, v in sd.items() if module in k}
        f = os.path.join(out_dir, f"{module}_weights_{idx}.bin")
        torch.save(part, f)
        written.append(f)
    return written
