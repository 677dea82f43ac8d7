"""B200-native AutoencoderKL (SURVEY.md §8 (f)4 tail: VAE encode of the condition image / decode of the final latents) against
the torch fp32 restatement in oracle/vae_oracle.py on identical bf16-representable weights and inputs.

Bars: the HBM-bound kernels match torch on the same bf16 inputs to bf16 rounding; the implicit-GEMM convolution matches
F.conv2d (fp32) to rel-L2 <= 3e-3; the model per block, TEACHER-FORCED (every oracle stage is fed the native trace's input of that
stage): rel-L2 <= 1e-2; free-running through the ~30 bf16 layers of an encode / decode: cosine >= 0.999 and rel-L2 <= 2.5e-2 on the
decoded image / the latents."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_l2(got, want):
    got, want = got.float().cpu(), want.float().cpu()
    return ((got - want).norm() / want.norm().clamp_min(1e-12)).item()


def _bf(t):
    return t.to(torch.bfloat16)


@pytest.mark.parametrize("B,H,W,Ci,Co,variant", [
    (2, 16, 32, 64, 128, 0), (1, 8, 128, 128, 64, 0), (1, 4, 256, 64, 64, 0), (1, 24, 64, 192, 256, 0), (1, 6, 16, 64, 64, 0),
    (1, 32, 128, 128, 256, 1), (1, 32, 128, 128, 256, 2), (1, 32, 128, 128, 256, 3), (1, 32, 128, 128, 256, 4),
    (1, 32, 128, 128, 256, 5), (1, 32, 128, 128, 256, 6), (3, 10, 8, 64, 72, 0), (1, 32, 128, 128, 128, 7), (1, 48, 128, 64, 72, 7),
    (1, 256, 256, 64, 128, 0)])
def test_implicit_conv3x3_matches_conv2d(B, H, W, Ci, Co, variant):
    from unigen_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + H * 10 + W + Ci + Co + variant)
    x = _bf(torch.randn(B, Ci, H, W, generator=g))
    w = _bf(torch.randn(Co, Ci, 3, 3, generator=g) / (3 * Ci ** 0.5))
    bias = _bf(torch.randn(Co, generator=g))
    res = _bf(torch.randn(B, Co, H, W, generator=g))
    want = F.conv2d(x.float(), w.float(), bias.float(), padding=1)
    xn = x.permute(0, 2, 3, 1).contiguous().cuda()
    wm = w.permute(0, 2, 3, 1).reshape(Co, 9 * Ci).contiguous().cuda()
    got = ops.conv3x3(xn, wm, bias=bias.cuda(), variant=variant)
    assert got.shape == (B, H, W, Co)
    assert rel_l2(got.permute(0, 3, 1, 2), want) < 3e-3
    # residual + alpha in the epilogue, output pixel stride wider than c_out (the padding channels are left untouched)
    out = torch.full((B, H, W, Co + 8), 7.0, device="cuda", dtype=torch.bfloat16)
    rn = torch.zeros(B, H, W, Co + 8, device="cuda", dtype=torch.bfloat16)
    rn[..., :Co] = res.permute(0, 2, 3, 1).cuda()
    ops.conv3x3(xn, wm, bias=bias.cuda(), residual=rn, out=out, alpha=0.5, variant=variant)
    assert rel_l2(out[..., :Co].permute(0, 3, 1, 2), 0.5 * want + res.float()) < 3e-3
    assert (out[..., Co:] == 7.0).all()


def test_conv3x3_rejects_shapes_the_implicit_path_does_not_tile():
    from unigen_b200 import ops
    x = torch.zeros(1, 8, 96, 64, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(ops.UgError):
        ops.conv3x3(x, torch.zeros(64, 9 * 64, device="cuda", dtype=torch.bfloat16))
    assert not ops.conv3x3_implicit_ok(64, 96) and not ops.conv3x3_implicit_ok(16, 128) and ops.conv3x3_implicit_ok(512, 1024)
    with pytest.raises(ops.UgError):
        ops.conv3x3(torch.zeros(1, 8, 32, 64, dtype=torch.bfloat16), torch.zeros(64, 576, dtype=torch.bfloat16))  # CPU tensors


@pytest.mark.parametrize("layout,dtype,C,stride", [("nhwc", torch.bfloat16, 64, 2), ("nhwc", torch.bfloat16, 24, 1),
                                                   ("nchw", torch.float32, 3, 1), ("nchw", torch.bfloat16, 16, 1)])
def test_im2col_matches_unfold(layout, dtype, C, stride):
    from unigen_b200 import ops
    g = torch.Generator().manual_seed(C + stride)
    B, H, W = 2, 10, 12
    x = torch.randn(B, C, H, W, generator=g).to(dtype)
    if stride == 2:  # Downsample2D: pad (0, 1, 0, 1), stride 2, no further padding
        want = F.unfold(F.pad(x.float(), (0, 1, 0, 1)), 3, stride=2)
        ho, wo, pad = H // 2, W // 2, 0
    else:
        want = F.unfold(x.float(), 3, padding=1)
        ho, wo, pad = H, W, 1
    # unfold rows are (c, ky, kx); ours are (ky, kx, c)
    want = want.view(B, C, 9, ho * wo).permute(0, 3, 2, 1).reshape(B * ho * wo, 9 * C)
    xin = (x.permute(0, 2, 3, 1).contiguous() if layout == "nhwc" else x).cuda()
    k_pad = (9 * C + 63) // 64 * 64
    got = ops.im2col(xin, layout, 3, 3, stride, pad, pad, ho, wo, k_pad)
    assert got.shape == (B * ho * wo, k_pad)
    assert torch.equal(got[:, :9 * C].float().cpu(), want.to(torch.bfloat16).float())
    assert (got[:, 9 * C:] == 0).all()
    if layout == "nchw":  # affine on in-image values only: the zero padding stays zero
        got2 = ops.im2col(xin, layout, 3, 3, stride, pad, pad, ho, wo, k_pad, alpha=2.0, beta=0.5)
        inside = F.unfold(torch.ones(B, C, H, W), 3, padding=1).view(B, C, 9, ho * wo).permute(0, 3, 2, 1).reshape(B * ho * wo, 9 * C)
        want2 = ((2.0 * want + 0.5) * inside).to(torch.bfloat16).float()
        assert torch.allclose(got2[:, :9 * C].float().cpu(), want2, atol=2e-2, rtol=1e-2)


@pytest.mark.parametrize("B,H,W,C,G,silu", [(2, 9, 7, 64, 32, True), (1, 32, 32, 128, 32, False), (1, 64, 64, 512, 32, True),
                                            (3, 5, 8, 192, 32, True)])
def test_groupnorm_matches_torch(B, H, W, C, G, silu):
    from unigen_b200 import ops
    g = torch.Generator().manual_seed(C + H)
    x = _bf(torch.randn(B, C, H, W, generator=g) * 2 + 0.7)
    gamma, beta = _bf(1 + 0.2 * torch.randn(C, generator=g)), _bf(0.3 * torch.randn(C, generator=g))
    want = F.group_norm(x.float(), G, gamma.float(), beta.float(), eps=1e-6)
    if silu:
        want = F.silu(want)
    xn = x.permute(0, 2, 3, 1).contiguous().cuda()
    got = ops.groupnorm(xn, gamma.cuda(), beta.cuda(), G, silu_act=silu)
    assert rel_l2(got.permute(0, 3, 1, 2), want) < 4e-3
    ops.groupnorm(xn, gamma.cuda(), beta.cuda(), G, silu_act=silu, out=xn)  # in place
    assert torch.equal(xn, got)


def test_upsample_softmax_layout_and_sampling_kernels():
    from unigen_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = _bf(torch.randn(2, 64, 6, 5, generator=g))
    xn = x.permute(0, 2, 3, 1).contiguous().cuda()
    up = ops.upsample2x(xn)
    assert torch.equal(up.permute(0, 3, 1, 2).cpu(), F.interpolate(x.float(), scale_factor=2.0, mode="nearest").to(torch.bfloat16))
    s = _bf(torch.randn(37, 1024, generator=g) * 3).cuda()
    want = torch.softmax(s.float(), dim=-1)
    ops.softmax_rows_(s)
    assert rel_l2(s, want) < 4e-3 and abs(s.float().sum(-1) - 1).max() < 2e-2
    back = ops.nhwc_to_nchw(xn, 48, dtype=torch.float32)
    assert torch.equal(back.cpu(), x[:, :48].float())
    assert torch.equal(ops.nhwc_to_nchw(xn, 64).cpu(), x)
    m = _bf(torch.randn(2, 4, 4, 64, generator=g)).cuda()  # 16 latent channels: mean | logvar | padding
    noise = torch.randn(2, 16, 4, 4, generator=g).cuda()
    mm = m.float().permute(0, 3, 1, 2)
    want_z = (mm[:, :16] + torch.exp(0.5 * mm[:, 16:32].clamp(-30, 20)) * noise - 0.1159) * 0.3611
    assert rel_l2(ops.vae_sample(m, 16, noise, 0.1159, 0.3611), want_z) < 4e-3
    assert torch.equal(ops.vae_sample(m, 16, None).float(), mm[:, :16].contiguous())


def _setup(cfg=None, seed=0):
    from oracle import vae_oracle as O
    from unigen_b200.vae import AutoencoderKL
    cfg = cfg or O.VAEConfig.tiny()
    sd = {k: v.to(torch.bfloat16).float() for k, v in O.init_state_dict(cfg, seed=seed).items()}
    oracle = O.VAEOracle(cfg, sd)
    oracle.record = True
    model = AutoencoderKL(in_channels=cfg.in_channels, out_channels=cfg.out_channels, latent_channels=cfg.latent_channels,
                          block_out_channels=cfg.block_out_channels, layers_per_block=cfg.layers_per_block,
                          norm_num_groups=cfg.norm_num_groups, scaling_factor=cfg.scaling_factor, shift_factor=cfg.shift_factor,
                          device="cuda")
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return cfg, sd, oracle, model


def _teacher_forced(stages, first_input, trace):
    """rel-L2 of every native trace point against the oracle stage evaluated on the NATIVE input of that stage."""
    worst, x = {}, first_input
    for name, fn in stages:
        with torch.no_grad():
            worst[name] = rel_l2(trace[name], fn(x.float()))
        x = trace[name].to(first_input.device)
    return worst


def test_state_dict_keys_and_shapes_are_the_diffusers_names():
    cfg, sd, oracle, model = _setup()
    mine = model.state_dict()
    assert set(mine) == set(sd)
    assert all(tuple(mine[k].shape) == tuple(sd[k].shape) for k in sd)
    assert mine["encoder.conv_in.weight"].shape == (64, 3, 3, 3) and mine["decoder.conv_out.weight"].shape == (3, 64, 3, 3)
    assert torch.equal(mine["decoder.up_blocks.0.resnets.0.conv1.weight"].float().cpu(), sd["decoder.up_blocks.0.resnets.0.conv1.weight"])


@pytest.mark.parametrize("H,W", [(64, 64), (96, 64), (32, 256)])
def test_tiny_encode_decode_match_oracle_per_block(H, W):
    """(96, 64): the level-0 width 96 neither divides nor is a multiple of 128 -> the patch-gather path of every convolution."""
    cfg, sd, oracle, model = _setup()
    g = torch.Generator().manual_seed(H + W)
    img = _bf(torch.rand(2, 3, H, W, generator=g) * 2 - 1).float()
    noise = torch.randn(2, cfg.latent_channels, H // 2, W // 2, generator=g)
    want_lat = oracle.encode(img, noise)
    model.trace = {}
    dist = model.encode(img.cuda()).latent_dist
    got_lat = ops_lat = None
    from unigen_b200 import ops
    got_lat = ops.vae_sample(dist._m, cfg.latent_channels, noise.cuda(), cfg.shift_factor, cfg.scaling_factor)
    assert got_lat.shape == want_lat.shape == (2, 16, H // 2, W // 2)
    lat_in = _bf(want_lat).float()
    want_img = oracle.decode(lat_in)
    got_img = model.decode_latents(lat_in.cuda())
    torch.cuda.synchronize()
    assert got_img.shape == want_img.shape == (2, 3, H, W)
    trace = {k: v.cpu() for k, v in model.trace.items()}
    assert set(trace) == set(oracle.trace) and len(trace) >= 10
    worst = _teacher_forced(oracle.encoder_stages(), img, trace)
    worst.update(_teacher_forced(oracle.decoder_stages(), lat_in / cfg.scaling_factor + cfg.shift_factor, trace))
    bad = {k: v for k, v in worst.items() if v > 1e-2}
    assert not bad, f"teacher-forced per-block rel-L2 above 1e-2: {bad}"
    cos = lambda a, b: F.cosine_similarity(a.float().cpu().flatten(), b.flatten(), dim=0).item()  # noqa: E731
    assert cos(got_lat, want_lat) >= 0.999 and cos(got_img, want_img) >= 0.999
    assert rel_l2(got_lat, want_lat) < 2.5e-2 and rel_l2(got_img, want_img) < 2.5e-2
    # the diffusers-shaped API: unscaled sample / mode, decode of raw z
    z = model.encode(img.cuda()).latent_dist.mode()
    assert rel_l2(z, oracle.encoder(img)[:, :16]) < 2.5e-2
    raw = model.decode(lat_in.cuda(), return_dict=False)[0]
    assert rel_l2(raw, oracle.decoder(lat_in)) < 2.5e-2
    assert model.last_launches > 20


def test_encode_condition_is_the_pipeline_three_liner_and_rejects_bad_inputs():
    from unigen_b200 import ops
    from unigen_b200.vae import AutoencoderKL
    cfg, sd, oracle, model = _setup()
    img = _bf(torch.rand(1, 3, 64, 64) * 2 - 1).float().cuda()
    gen = torch.Generator(device="cuda").manual_seed(7)
    a = model.encode_condition(img, generator=gen)
    gen.manual_seed(7)
    z = model.encode(img).latent_dist.sample(gen)
    b = ((z.float() - cfg.shift_factor) * cfg.scaling_factor)
    assert rel_l2(a, b) < 1e-2
    assert torch.equal(model.encode_condition(img, sample=False, use_shift_factor=False),
                       model.encode(img).latent_dist.mode(0.0, cfg.scaling_factor))
    with pytest.raises(ops.UgError):
        model.encode(torch.zeros(1, 4, 64, 64))
    with pytest.raises(ops.UgError):
        model.decode(torch.zeros(1, 4, 8, 8))
    with pytest.raises(ops.UgError):
        AutoencoderKL(device="cpu")
    with pytest.raises(ops.UgError):
        AutoencoderKL(use_quant_conv=True, device="cuda")


def test_flux_vae_full_size_decode_and_encode_match_fp32_oracle_on_gpu():
    """FLUX.1 VAE architecture (128-256-512-512, 16 latent channels) at 512 x 512: the oracle runs in fp32 ON the GPU (TF32 off) over
    the same bf16-representable weights; teacher-forced per-block rel-L2 <= 1e-2, free-running image / latent cosine >= 0.999."""
    from oracle import vae_oracle as O
    cfg, sd, oracle, model = _setup(O.VAEConfig.flux(), seed=3)
    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        oracle.sd = {k: v.cuda() for k, v in sd.items()}
        g = torch.Generator().manual_seed(1)
        lat = _bf(torch.randn(1, 16, 64, 64, generator=g)).float().cuda()
        img = _bf(torch.rand(1, 3, 512, 512, generator=g) * 2 - 1).float().cuda()
        model.trace = {}
        got_img = model.decode_latents(lat)
        got_lat = model.encode_condition(img, sample=False)
        with torch.no_grad():
            want_img = oracle.decode(lat)
            want_lat = oracle.encode(img, None)
        worst = _teacher_forced(oracle.encoder_stages(), img, model.trace)
        worst.update(_teacher_forced(oracle.decoder_stages(), lat / cfg.scaling_factor + cfg.shift_factor, model.trace))
        assert set(worst) == set(oracle.trace)
        bad = {k: v for k, v in worst.items() if v > 1e-2}
        assert not bad, f"teacher-forced per-block rel-L2 above 1e-2: {bad}"
        cos = lambda a, b: F.cosine_similarity(a.float().flatten(), b.float().flatten(), dim=0).item()  # noqa: E731
        assert cos(got_img, want_img) >= 0.999 and cos(got_lat, want_lat) >= 0.999
        assert rel_l2(got_img, want_img) < 2.5e-2 and rel_l2(got_lat, want_lat) < 2.5e-2
        import json
        import os
        if os.environ.get("UG_PARITY_OUT"):
            with open(os.path.join(os.environ["UG_PARITY_OUT"], "r02_parity_vae_flux_512.json"), "w") as f:
                json.dump(dict(teacher_forced_rel_l2=worst, image_cosine=cos(got_img, want_img), latent_cosine=cos(got_lat, want_lat),
                               image_rel_l2=rel_l2(got_img, want_img), latent_rel_l2=rel_l2(got_lat, want_lat)), f, indent=1)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def test_flux_pipeline_encodes_the_condition_image_and_decodes_the_latents_through_the_native_vae():
    """`UniGenFLUXPipeline.__call__` with pixels on both sides (src/UniGenPipeline.py:954-961 vae.encode -> sample -> shift / scale ->
    pack; :1123-1125 unpack -> latents / scaling + shift -> vae.decode) == the same steps done by hand around the latent-space call."""
    from oracle import unigen_oracle as FO
    from unigen_b200 import ops
    from unigen_b200.condition import Condition
    from unigen_b200.model import FluxArch, UniGenFlux, canonical_control_params
    from unigen_b200.pipeline import UniGenFLUXPipeline
    cfg, sd, oracle, vae = _setup()
    tr = UniGenFlux(FluxArch.tiny(), device="cuda")
    tr.init_condition_block(condition_nums=1, control_params=canonical_control_params())
    tr.init_random_(seed=2)
    E = tr.expert_nums
    H = W = 64  # pixels: latents 8 x 8 (the tiny VAE has one down-sampling level: factor 2) -> use vae_scale_factor = 2
    g = torch.Generator().manual_seed(9)
    img = _bf(torch.rand(1, 3, H, W, generator=g) * 2 - 1).float().cuda()
    pe, pp, cp = torch.randn(1, 32, 4096, generator=g).cuda(), torch.randn(1, 768, generator=g).cuda(), torch.randn(1, 768, generator=g).cuda()
    lat0 = _bf(torch.randn(1, 16, H // 2, W // 2, generator=g)).cuda()
    N = (H // 4) * (W // 4)
    rts = [torch.rand(N, E, generator=g).cuda() for _ in range(2)]
    pipe = UniGenFLUXPipeline(tr, vae=vae, vae_scale_factor=2)
    kw = dict(prompt_embeds=pe, pooled_prompt_embeds=pp, condition_pooled_prompt_embeds=cp, height=H, width=W, num_inference_steps=2,
              latents=ops.pack_latents(lat0), rts_uniform=rts)
    gen = torch.Generator(device="cuda").manual_seed(4)
    image = pipe(control_image=img, condition_types=["canny"], generator=gen, output_type="pt", **kw).images
    assert image.shape == (1, 3, H, W)
    # by hand: encode -> pack, latent-space pipeline call, unpack -> decode
    gen.manual_seed(4)
    tokens, ids, type_id = Condition("canny", img).encode(pipe, gen)
    assert tokens.shape == (1, N, 64) and ids.shape == (N, 3) and int(type_id[0, 0]) == 1
    assert torch.equal(ids.cpu(), FO.prepare_latent_image_ids(H // 4, W // 4))
    lat_out = pipe(control_image=tokens, condition_ids=ids, output_type="latent", **kw).images
    want = vae.decode_latents(ops.unpack_latents(lat_out, H // 2, W // 2))
    assert torch.equal(image, want)
    sub_tokens, sub_ids, _ = Condition("subject", img).encode(pipe, gen)
    assert torch.equal(sub_ids[:, 2].cpu(), ids[:, 2].cpu() + (H // 2) // 2)  # :109-110


def test_from_pretrained_reads_a_local_diffusers_vae_folder(tmp_path):
    """`AutoencoderKL.from_pretrained(base, subfolder="vae")`: config.json + safetensors of a local diffusers folder; attention
    projections stored as 1x1 convolutions (checkpoints converted from the LDM layout) are accepted."""
    import json
    from safetensors.torch import save_file
    from unigen_b200 import ops
    from unigen_b200.vae import AutoencoderKL
    cfg, sd, oracle, model = _setup()
    d = tmp_path / "base" / "vae"
    d.mkdir(parents=True)
    (d / "config.json").write_text(json.dumps(dict(
        _class_name="AutoencoderKL", act_fn="silu", in_channels=3, out_channels=3, latent_channels=16,
        block_out_channels=list(cfg.block_out_channels), layers_per_block=cfg.layers_per_block, norm_num_groups=32,
        down_block_types=["DownEncoderBlock2D"] * 2, up_block_types=["UpDecoderBlock2D"] * 2, scaling_factor=cfg.scaling_factor,
        shift_factor=cfg.shift_factor, use_quant_conv=False, use_post_quant_conv=False, sample_size=1024, force_upcast=True)))
    as_conv = {k: (v[:, :, None, None] if ".attentions.0.to_" in k and k.endswith(".weight") else v) for k, v in sd.items()}
    save_file({k: v.contiguous() for k, v in as_conv.items()}, str(d / "diffusion_pytorch_model.safetensors"))
    m2 = AutoencoderKL.from_pretrained(str(tmp_path / "base"), subfolder="vae", torch_dtype=torch.bfloat16).to("cuda")
    m2.requires_grad_(False).eval()
    assert m2.config.block_out_channels == tuple(cfg.block_out_channels) and m2.config.scaling_factor == cfg.scaling_factor
    lat = _bf(torch.randn(1, 16, 32, 32)).cuda()
    assert torch.equal(m2.decode_latents(lat), model.decode_latents(lat))
    with pytest.raises(OSError):
        AutoencoderKL.from_pretrained(str(tmp_path / "nowhere"))
    (d / "config.json").write_text(json.dumps(dict(block_out_channels=[64, 128], layers_per_block=1, act_fn="gelu")))
    with pytest.raises(ops.UgError):
        AutoencoderKL.from_pretrained(str(d))
    with pytest.raises(ops.UgError):
        m2.to("cpu")


def test_mid_attention_score_chunks_are_bit_identical_to_one_chunk():
    cfg, sd, oracle, model = _setup()
    lat = _bf(torch.randn(2, 16, 32, 32, generator=torch.Generator().manual_seed(2))).cuda()
    whole = model.decode_latents(lat)
    model.max_score_elems = 192 * 1024  # 1024 tokens -> chunks of 192 query rows (the last one ragged)
    assert torch.equal(model.decode_latents(lat), whole)
