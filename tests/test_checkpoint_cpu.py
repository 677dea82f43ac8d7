"""CPU: weight interchange (unigen_b200/checkpoint.py) — every `--transformer` form infer.py:124-141 accepts."""
import json
import math
import os
import types
from collections import OrderedDict

import pytest
import torch

from unigen_b200 import checkpoint as ck


def _params(seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = OrderedDict([("control_x_embedder.weight", (6, 5)), ("control_x_embedder.bias", (6,)),
                          ("moe.moe_layer.gate.wg.weight", (3, 7)), ("controlnet_add_joint_blocks.0.weight", (4, 4)),
                          ("shared_expert.0.norm1.linear.weight", (11, 3))])
    return OrderedDict((k, torch.randn(*s, generator=g)) for k, s in shapes.items())


def _write_zero3(tmp, params, frozen, world, tag="global_step7"):
    d = tmp / tag
    d.mkdir()
    (tmp / "latest").write_text(tag)
    groups = [list(params.items())[:2], list(params.items())[2:]]  # two optimizer groups
    flat = [[[] for _ in groups] for _ in range(world)]
    for gi, grp in enumerate(groups):
        for _, p in grp:
            n = p.numel()
            part = math.ceil(n / world)
            padded = torch.cat([p.reshape(-1), torch.zeros(part * world - n)])
            for r in range(world):
                flat[r][gi].append(padded[r * part:(r + 1) * part])
    for r in range(world):
        torch.save({"optimizer_state_dict": {ck.ZERO_STAGE: 3, ck.PARTITION_COUNT: world,
                                             ck.FP32_FLAT_GROUPS: [torch.cat(x) for x in flat[r]]}},
                   d / f"bf16_zero_pp_rank_{r}_mp_rank_00_optim_states.pt")
        frag = {}
        for k, p in frozen.items():
            part = math.ceil(p.numel() / world)
            padded = torch.cat([p.reshape(-1), torch.zeros(part * world - p.numel())])
            frag[k] = padded[r * part:(r + 1) * part].to(torch.bfloat16)
        torch.save({"module": {"pos_embed.pos_embed": torch.arange(4.0)}, ck.BUFFER_NAMES: ["pos_embed.pos_embed"],
                    ck.PARAM_SHAPES: [OrderedDict((k, p.shape) for k, p in grp) for grp in groups],
                    ck.FROZEN_PARAM_SHAPES: OrderedDict((k, p.shape) for k, p in frozen.items()), ck.FROZEN_PARAM_FRAGMENTS: frag,
                    ck.SHARED_PARAMS: [["alias.weight", "control_x_embedder.weight"]]},
                   d / f"zero_pp_rank_{r}_mp_rank_00_model_states.pt")


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_zero3_shards_consolidate_to_the_original_tensors(tmp_path, world):
    params = _params()
    frozen = OrderedDict([("transformer_blocks.0.attn.to_q.weight", torch.randn(5, 5).bfloat16().float())])
    _write_zero3(tmp_path, params, frozen, world)
    sd = ck.read_state_dict(str(tmp_path))
    for k, v in params.items():
        assert torch.equal(sd[k], v), k
    assert torch.equal(sd["transformer_blocks.0.attn.to_q.weight"], frozen["transformer_blocks.0.attn.to_q.weight"])
    assert torch.equal(sd["alias.weight"], params["control_x_embedder.weight"]) and torch.equal(sd["pos_embed.pos_embed"], torch.arange(4.0))
    # the pre-merged file script/infer.sh writes wins when present
    torch.save({"only": torch.ones(1)}, tmp_path / "global_step7" / "pytorch_model_fp32.bin")
    assert list(ck.read_state_dict(str(tmp_path))) == ["only"]


def test_zero2_shards(tmp_path):
    params, world, tag = _params(1), 4, "global_step3"
    (tmp_path / tag).mkdir()
    (tmp_path / "latest").write_text(tag)
    flat = torch.cat([p.reshape(-1) for p in params.values()])
    pad = (-flat.numel()) % (2 * world)
    flat = torch.cat([flat, torch.zeros(pad)])
    part = flat.numel() // world
    for r in range(world):
        torch.save({"optimizer_state_dict": {ck.ZERO_STAGE: 2, ck.PARTITION_COUNT: [world],
                                             ck.SINGLE_PARTITION: [flat[r * part:(r + 1) * part]]}},
                   tmp_path / tag / f"zero_pp_rank_{r}_mp_rank_00_optim_states.pt")
    torch.save({"module": {}, ck.BUFFER_NAMES: [], ck.PARAM_SHAPES: [OrderedDict((k, p.shape) for k, p in params.items())],
                ck.FROZEN_PARAM_SHAPES: {"frozen.w": torch.Size([2])}, ck.FROZEN_PARAM_FRAGMENTS: {"frozen.w": torch.tensor([1.0, 2.0])}},
               tmp_path / tag / "mp_rank_00_model_states.pt")
    sd = ck.consolidate_zero_checkpoint(str(tmp_path))
    assert all(torch.equal(sd[k], v) for k, v in params.items()) and torch.equal(sd["frozen.w"], torch.tensor([1.0, 2.0]))


def test_safetensors_directory_sharded_index_and_torch_file(tmp_path):
    from safetensors.torch import save_file
    params = _params(2)
    keys = list(params)
    d = tmp_path / "transformer"
    d.mkdir()
    save_file({k: params[k] for k in keys[:2]}, str(d / "diffusion_pytorch_model-00001-of-00002.safetensors"))
    save_file({k: params[k] for k in keys[2:]}, str(d / "diffusion_pytorch_model-00002-of-00002.safetensors"))
    sd = ck.read_state_dict(str(d))
    assert set(sd) == set(keys) and all(torch.equal(sd[k], params[k]) for k in keys)
    save_file({"stale": torch.zeros(1)}, str(d / "zz_unreferenced.safetensors"))
    (d / "diffusion_pytorch_model.safetensors.index.json").write_text(json.dumps({"weight_map": {
        **{k: "diffusion_pytorch_model-00001-of-00002.safetensors" for k in keys[:2]},
        **{k: "diffusion_pytorch_model-00002-of-00002.safetensors" for k in keys[2:]}}}))
    assert "stale" not in ck.read_state_dict(str(d))
    torch.save(dict(params), tmp_path / "model.bin")
    assert set(ck.read_state_dict(str(tmp_path / "model.bin"))) == set(keys)
    with pytest.raises(FileNotFoundError):
        ck.read_state_dict(str(tmp_path / "nope"))


class _FakeModel:
    def __init__(self, sd):
        self._sd = {k: v.clone() for k, v in sd.items()}
        self.trainable_control_modules = {"control_x_embedder": None, "moe": None, "controlnet_add_joint_blocks": None, "shared_expert": None}

    def state_dict(self):
        return self._sd

    def load_state_dict(self, sd, strict=True):
        missing = [k for k in self._sd if k not in sd]
        unexpected = [k for k in sd if k not in self._sd]
        for k, v in sd.items():
            if k in self._sd:
                self._sd[k] = v.clone()
        return types.SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)


def test_load_pretrained_and_save_modules_roundtrip(tmp_path):
    params = _params(3)
    base = OrderedDict([("x_embedder.weight", torch.randn(3, 3)), ("proj_out.bias", torch.randn(3))])
    model = _FakeModel({**{k: torch.zeros_like(v) for k, v in params.items()}, **{k: torch.zeros_like(v) for k, v in base.items()}})
    torch.save(dict(base), tmp_path / "base.bin")
    src = _FakeModel({**params, **base})
    files = ck.save_modules(src, str(tmp_path / "ckpt"), ["control_x_embedder", "moe", "controlnet_add_joint_blocks", "shared_expert"])
    assert sorted(os.path.basename(f) for f in files) == ["control_x_embedder_weights_0.bin", "controlnet_add_joint_blocks_weights_0.bin",
                                                           "moe_weights_0.bin", "shared_expert_weights_0.bin"]
    res = ck.load_pretrained(model, base=str(tmp_path / "base.bin"), control=str(tmp_path / "ckpt"))
    assert set(res.missing_keys) == set(base) and not res.unexpected_keys  # strict=False: the control checkpoint has no base keys
    assert all(torch.equal(model.state_dict()[k], v) for k, v in {**params, **base}.items())
    torch.save({"x_embedder.weight": base["x_embedder.weight"]}, tmp_path / "partial.bin")
    with pytest.raises(RuntimeError):
        ck.load_pretrained(model, base=str(tmp_path / "partial.bin"))


class _NotATensor:  # a python object inside a checkpoint: needs the full (code-executing) unpickler
    pass


def test_full_unpickler_is_an_explicit_opt_in(tmp_path):
    """A file that is not a plain tensor container raises instead of silently retrying with `weights_only=False`
    (arbitrary code execution from an untrusted control checkpoint); `trust_pickle=True` is the caller's opt-in.
    A corrupt file raises either way."""
    import pickle
    f = tmp_path / "model.bin"
    torch.save({"w": torch.ones(2), "extra": _NotATensor()}, f)
    with pytest.raises(pickle.UnpicklingError, match="trust_pickle"):
        ck.read_state_dict(str(f))
    sd = ck.read_state_dict(str(f), trust_pickle=True)
    assert torch.equal(sd["w"], torch.ones(2))
    bad = tmp_path / "bad.bin"
    bad.write_bytes(b"not a checkpoint")
    with pytest.raises(Exception):
        ck.read_state_dict(str(bad))
    with pytest.raises(Exception):
        ck.read_state_dict(str(bad), trust_pickle=True)
