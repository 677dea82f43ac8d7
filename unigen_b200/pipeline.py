"""Denoise loop around the B200-native transformer with the scheduler kept on the device.

`UniGenFLUXPipeline.__call__` mirrors the part of the reference `UniGenFLUXPipeline.__call__` (src/UniGenPipeline.py:810-1134)
between prompt encoding and VAE decode: latent preparation (:978-987), the sigma schedule (:989-1006), the loop body
(:1050-1116: transformer call with `timestep / 1000`, optional true-CFG second call + combine, Euler flow-match step) and the
`output_type="latent"` exit (:1120-1121) or, with `vae=` (the native `unigen_b200.vae.AutoencoderKL`), the VAE encode of the
condition image (:954-961) and the decode of the final latents (:1123-1125). Text encoders and image pre-processing (PIL ->
[-1, 1] tensors) are callers of the path and stay out of scope: prompts arrive as embeddings, condition images as packed latents
(`unigen_b200.condition.Condition`) or pre-processed pixel tensors.

The whole loop is ONE CUDA graph (`graph_loop=True`, default): the sigma / timestep tables live in device memory
(`ug_timestep_embedding` reads entry i, `ug_euler_step_table` reads entries i, i+1), the per-step RTS uniform draws are a
device table, the CFG combine runs in-graph — so an image costs exactly one graph launch and no host work between steps.
`denoise_sd3` is the same for the SD3.5 loop (src/UniGenPipeline.py:377-412): batch-doubled classifier-free guidance with
`noise_uncond + g * (noise_text - noise_uncond)` (:405-412) inside the graph."""
from __future__ import annotations

import math
import types
from typing import Any, Dict, List, Optional, Sequence

import torch

from . import condition as cond_mod
from . import ops


def calculate_shift(image_seq_len: int, base_seq_len: int = 256, max_seq_len: int = 4096, base_shift: float = 0.5,
                    max_shift: float = 1.15) -> float:
    """diffusers pipeline_flux.calculate_shift as called at src/UniGenPipeline.py:991-997 (the reference passes
    `scheduler.config.get("max_shift", 1.15)`)."""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


def flow_match_sigmas(num_inference_steps: int, image_seq_len: int, use_dynamic_shifting: bool = True,
                      sigmas: Optional[Sequence[float]] = None, shift: float = 1.0, num_train_timesteps: int = 1000) -> List[float]:
    """FlowMatchEulerDiscreteScheduler.set_timesteps. Flux (dynamic shifting): the pipeline passes sigmas = linspace(1, 1/n, n)
    (src/UniGenPipeline.py:989) and the scheduler applies the exponential time shift with mu = calculate_shift(seq_len). SD3.5 (static
    `shift`, 3.0): the pipeline passes NO sigmas (:345-351), so the scheduler starts from linspace(sigma_max, sigma_min, n) with
    sigma_max / sigma_min the ends of its ALREADY shifted training schedule (1 and shift / 1000 / (1 + (shift - 1) / 1000)) and
    shifts once more. A terminal 0 is appended; timestep_i = 1000 sigma_i."""
    n = num_inference_steps
    if sigmas is not None:
        sig = list(sigmas)
    elif use_dynamic_shifting:
        sig = [1.0 + (1.0 / n - 1.0) * i / max(n - 1, 1) for i in range(n)]
    else:
        lo = 1.0 / num_train_timesteps
        s_min = shift * lo / (1.0 + (shift - 1.0) * lo)
        sig = [1.0 + (s_min - 1.0) * i / max(n - 1, 1) for i in range(n)]
    if use_dynamic_shifting:
        mu = calculate_shift(image_seq_len)
        sig = [math.exp(mu) / (math.exp(mu) + (1.0 / s - 1.0)) for s in sig]
    elif shift != 1.0:
        sig = [shift * s / (1.0 + (shift - 1.0) * s) for s in sig]
    return sig + [0.0]


class _LoopGraphs:
    """CUDA graphs of whole sampling loops, one per (shape, steps, flags) key, with their static input buffers."""

    def __init__(self):
        self.graphs: Dict[Any, Any] = {}

    def run(self, key, inputs: Dict[str, Optional[torch.Tensor]], body, use_graph: bool):
        if not use_graph:
            return body(inputs)
        g = self.graphs.get(key)
        if g is None:
            static = {k: (v.clone() if v is not None else None) for k, v in inputs.items()}
            snapshot = {k: (v.clone() if v is not None else None) for k, v in static.items()}
            body(static)  # warm-up on the static buffers: workspace + job-table allocation happen outside the capture
            torch.cuda.synchronize()
            for k, v in snapshot.items():  # the warm-up advanced the latents in place: restore before capturing
                if v is not None:
                    static[k].copy_(v)
            n0 = ops.launch_count()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = body(static)
            g = self.graphs[key] = (graph, static, out, ops.launch_count() - n0)
        graph, static, out, n_launch = g
        for k, v in inputs.items():
            if v is not None:
                static[k].copy_(v)
        graph.replay()
        ops.add_launches(n_launch)
        return out


_GRAPHS_ATTR = "_ug_loop_graphs"


def _loop_graphs(transformer) -> _LoopGraphs:
    g = getattr(transformer, _GRAPHS_ATTR, None)
    if g is None:
        g = _LoopGraphs()
        object.__setattr__(transformer, _GRAPHS_ATTR, g)
    return g


def _f32(t, dev):
    return t.to(device=dev, dtype=torch.float32).contiguous()


@torch.no_grad()
def denoise(transformer, latents: torch.Tensor, condition_latents, encoder_hidden_states, pooled_projections,
            condition_pooled_projections, img_ids, txt_ids, condition_ids, num_inference_steps: int = 4,
            guidance: Optional[torch.Tensor] = None, conditioning_scale: float = 1.0, use_dynamic_shifting: bool = True,
            rts_uniform: Optional[Sequence] = None, sigmas: Optional[Sequence[float]] = None, true_cfg_scale: float = 1.0,
            negative_encoder_hidden_states=None, negative_pooled_projections=None, negative_txt_ids=None,
            graph_loop: bool = True) -> torch.Tensor:
    """Runs the sampling loop on packed latents (B, N, 64) and returns the final packed latents (bf16, on the device).
    Per step: velocity = transformer(latents, ..., timestep = sigma_i)[0] (+ the true-CFG second call and combine,
    src/UniGenPipeline.py:1076-1091); latents += (sigma_{i+1} - sigma_i) * velocity.
    `rts_uniform`: per step (and, under true CFG, per call: [steps][2]) the uniform draw of the MoE gate; None = fresh draws."""
    dev = transformer.device
    n_steps = int(num_inference_steps)
    multi = isinstance(condition_latents, (list, tuple))
    conds = list(condition_latents) if multi else [condition_latents]
    cpools = list(condition_pooled_projections) if multi else [condition_pooled_projections]
    cids = list(condition_ids) if multi else [condition_ids]
    n_cond = len(conds)
    x0 = ops.to_bf16(latents.to(dev).contiguous()).clone()  # the loop updates the latents in place: never the caller's tensor
    B, N, _ = x0.shape
    T = encoder_hidden_states.shape[1]
    E = transformer.expert_nums
    do_cfg = true_cfg_scale > 1 and negative_encoder_hidden_states is not None
    n_calls = 2 if do_cfg else 1
    sig = flow_match_sigmas(n_steps, N, use_dynamic_shifting, sigmas)
    sig32 = torch.tensor(sig, dtype=torch.float32)
    # the pipeline hands the transformer `timestep / 1000` with timestep = sigma * 1000 (fp32, :1063): same two roundings here
    t_in = (sig32[:n_steps] * 1000.0) / 1000.0
    sq = lambda t: t[0] if t.dim() == 3 else t  # noqa: E731
    inputs: Dict[str, Optional[torch.Tensor]] = dict(
        x=x0, es=ops.to_bf16(encoder_hidden_states.to(dev).contiguous()), pooled=_f32(pooled_projections, dev),
        guidance=_f32(guidance, dev) if guidance is not None else None, txt_ids=_f32(sq(txt_ids), dev), img_ids=_f32(sq(img_ids), dev),
        sigmas=sig32.to(dev), t_in=t_in.to(dev))
    for c in range(n_cond):
        cl = conds[c].to(dev)
        inputs[f"cs{c}"] = ops.to_bf16((cl if cl.dim() == 3 else cl.unsqueeze(0)).contiguous())
        cp = _f32(cpools[c], dev)
        inputs[f"cp{c}"] = cp if cp.dim() == 2 else cp.unsqueeze(0)
        inputs[f"cid{c}"] = _f32(sq(cids[c]), dev)
        if rts_uniform is None:  # DeepSpeed draws a fresh uniform tensor in every MoE call: one table per loop
            u = torch.rand(n_steps, n_calls, B * N, E, device=dev, dtype=torch.float32)
        else:
            rows = []
            for i in range(n_steps):
                ui = rts_uniform[i]
                ui = ui[c] if multi else ui
                per_call = list(ui) if isinstance(ui, (list, tuple)) else [ui]  # [conditional call, negative call] or one draw
                per_call = (per_call * n_calls)[:n_calls]
                rows.append(torch.stack([_f32(t, dev) for t in per_call], 0))
            u = torch.stack(rows, 0)
        inputs[f"u{c}"] = u
    if do_cfg:
        inputs["neg_es"] = ops.to_bf16(negative_encoder_hidden_states.to(dev).contiguous())
        inputs["neg_pooled"] = _f32(negative_pooled_projections, dev)
        inputs["neg_txt_ids"] = _f32(sq(negative_txt_ids if negative_txt_ids is not None else txt_ids), dev)

    def call(st, i, es, pooled, txt, call_idx):
        kw = {}
        for c in range(n_cond):
            kw[f"cs{c}"], kw[f"cp{c}"], kw[f"cid{c}"] = st[f"cs{c}"], st[f"cp{c}"], st[f"cid{c}"]
            kw[f"u{c}"] = st[f"u{c}"][i, call_idx]
        # a ONE-element view of the device timestep table: the embedding kernel broadcasts it to the batch
        return transformer._forward_impl(float(conditioning_scale), st["x"], es, pooled, st["t_in"][i:i + 1], st["guidance"], txt,
                                         st["img_ids"], **kw)[0]

    def body(st):
        x = st["x"]
        for i in range(n_steps):
            v = call(st, i, st["es"], st["pooled"], st["txt_ids"], 0)
            if do_cfg:
                # the workspace's velocity buffer is reused by the second call: keep the conditional prediction
                v_pos = ops.copy(v, st["v_pos"]) if "v_pos" in st else v.clone()
                v_neg = call(st, i, st["neg_es"], st["neg_pooled"], st["neg_txt_ids"], 1)
                v = ops.cfg_combine(v_neg, v_pos, float(true_cfg_scale), out=v_pos)
            ops.euler_step_table(x, v if v.is_contiguous() else v.contiguous(), st["sigmas"], i)
        return x

    if do_cfg:
        inputs["v_pos"] = torch.empty_like(x0)
    key = (B, N, T, n_steps, n_cond, float(conditioning_scale), guidance is not None, do_cfg, float(true_cfg_scale),
           negative_encoder_hidden_states.shape[1] if do_cfg else 0)
    use_graph = bool(graph_loop) and getattr(transformer, "trace", None) is None and not getattr(transformer, "sp_world", 0)
    out = _loop_graphs(transformer).run(key, inputs, body, use_graph)
    if hasattr(transformer, "check_peer_errors"):
        transformer.check_peer_errors()
    return out.clone()


def sd3_controlnet_keep(i: int, n_steps: int, start: float = 0.0, end: float = 1.0) -> float:
    """src/UniGenPipeline.py:367-373: `1.0 - float(i / len(timesteps) < s or (i + 1) / len(timesteps) > e)`."""
    return 1.0 - float(i / n_steps < start or (i + 1) / n_steps > end)


@torch.no_grad()
def denoise_sd3(transformer, latents: torch.Tensor, condition_latents: torch.Tensor, encoder_hidden_states, pooled_projections,
                condition_pooled_projections, num_inference_steps: int = 28, guidance_scale: float = 7.0,
                negative_encoder_hidden_states=None, negative_pooled_projections=None, conditioning_scale: float = 1.0,
                shift: float = 3.0, rts_uniform: Optional[Sequence] = None, sigmas: Optional[Sequence[float]] = None,
                graph_loop: bool = True, control_guidance_start: float = 0.0, control_guidance_end: float = 1.0) -> torch.Tensor:
    """`UniGenSD3Pipeline.__call__` loop (src/UniGenPipeline.py:377-412) on latents (B, 16, H, W): with classifier-free guidance
    (`guidance_scale > 1` and negative embeddings) every step runs ONE batch-doubled forward over [uncond | text] (:380, prompt
    embeddings concatenated negative-first) and combines `uncond + g * (text - uncond)` (:405-407) on the device; Euler update
    with the static-shift flow-match schedule (SD3.5: shift 3.0). `control_guidance_start / end` gate the control branch per
    step exactly as `controlnet_keep` does (:367-373, :387-391: scale_i = conditioning_scale * keep_i). Returns the final
    latents (bf16)."""
    dev = transformer.device
    n_steps = int(num_inference_steps)
    scales = [float(conditioning_scale) * sd3_controlnet_keep(i, n_steps, control_guidance_start, control_guidance_end)
              for i in range(n_steps)]
    x0 = ops.to_bf16(latents.to(dev).contiguous()).clone()
    B = x0.shape[0]
    p = transformer.arch.patch_size
    N = (x0.shape[2] // p) * (x0.shape[3] // p)
    E = transformer.expert_nums
    do_cfg = guidance_scale > 1 and negative_encoder_hidden_states is not None
    Bf = 2 * B if do_cfg else B
    sig = flow_match_sigmas(n_steps, N, use_dynamic_shifting=False, sigmas=sigmas, shift=shift)
    sig32 = torch.tensor(sig, dtype=torch.float32)
    cat2 = lambda neg, pos: torch.cat([neg.to(dev), pos.to(dev)], 0) if do_cfg else pos.to(dev)  # noqa: E731
    cl = ops.to_bf16(condition_latents.to(dev).contiguous())
    inputs: Dict[str, Optional[torch.Tensor]] = dict(
        x=x0, xin=torch.empty(Bf, *x0.shape[1:], device=dev, dtype=x0.dtype),
        cs=torch.cat([cl, cl], 0) if do_cfg else cl,
        es=ops.to_bf16(cat2(negative_encoder_hidden_states, encoder_hidden_states).contiguous()),
        pooled=_f32(cat2(negative_pooled_projections, pooled_projections), dev),
        cpooled=_f32(torch.cat([condition_pooled_projections] * 2, 0) if do_cfg else condition_pooled_projections, dev),
        sigmas=sig32.to(dev), timesteps=(sig32[:n_steps] * 1000.0).to(dev))  # scheduler: timesteps = sigmas * num_train_timesteps
    if rts_uniform is None:
        inputs["u"] = torch.rand(n_steps, Bf * N, E, device=dev, dtype=torch.float32)
    else:
        inputs["u"] = torch.stack([_f32(rts_uniform[i], dev) for i in range(n_steps)], 0)
    if do_cfg:
        inputs["v"] = torch.empty_like(x0)

    def body(st):
        x = st["x"]
        xf, xi = x.view(B, 1, -1), st["xin"].view(Bf, 1, -1)
        for i in range(n_steps):
            ops.copy(xf, xi[:B])  # latent_model_input = cat([latents] * 2) (:380)
            if do_cfg:
                ops.copy(xf, xi[B:])
            out = transformer._forward_impl(scales[i], st["xin"], st["cs"], st["es"], st["pooled"], st["cpooled"],
                                            st["timesteps"][i:i + 1], st["u"][i])[0]
            if do_cfg:  # noise_pred_uncond, noise_pred_text = noise_pred.chunk(2)
                v = ops.cfg_combine(out[:B], out[B:], float(guidance_scale), out=st["v"])
            else:
                v = out
            ops.euler_step_table(x, v, st["sigmas"], i)
        return x

    key = ("sd3", tuple(x0.shape), encoder_hidden_states.shape[1], n_steps, tuple(scales), do_cfg, float(guidance_scale))
    use_graph = bool(graph_loop) and getattr(transformer, "trace", None) is None
    return _loop_graphs(transformer).run(key, inputs, body, use_graph).clone()


class UniGenFLUXPipeline:
    """`UniGenFLUXPipeline`-shaped entry (src/UniGenPipeline.py:810-1134) for the B200-native transformer. Same call-site
    keywords for everything the denoise path consumes; what needs a text encoder / VAE must arrive pre-computed:

      prompt_embeds, pooled_prompt_embeds         instead of prompt / prompt_2 (a string prompt raises: no text encoder here)
      condition_pooled_prompt_embeds              (tensor, or a list with one entry per condition)
      control_image                               packed condition latents (B, Nc, 64), a `condition.Condition`, or a list of either;
                                                  with `vae=` also preprocessed pixel images (B, 3, H, W) in [-1, 1]
      output_type="latent"                        or "pt" with `vae=` (unigen_b200.vae.AutoencoderKL: native encode / decode) or a
                                                  caller-owned `vae_decode=` callable

    Returns `SimpleNamespace(images=latents)` (FluxPipelineOutput-shaped) or `(latents,)` with return_dict=False."""

    def __init__(self, transformer, vae_decode=None, vae_scale_factor: int = 8, default_sample_size: int = 128, vae=None):
        self.transformer = transformer
        self.vae = vae
        if vae_decode is None and vae is not None:  # :1123-1125: latents / scaling_factor + shift_factor -> vae.decode
            vae_decode = vae.decode_latents
        self.vae_decode = vae_decode
        self.vae_scale_factor, self.default_sample_size = vae_scale_factor, default_sample_size
        self.joint_attention_kwargs = None

    @property
    def device(self):
        return self.transformer.device

    @property
    def dtype(self):
        return self.transformer.dtype

    @staticmethod
    def _pack_latents(latents):
        return ops.pack_latents(latents)

    @staticmethod
    def _unpack_latents(latents, height, width, vae_scale_factor=8):
        return ops.unpack_latents(latents, 2 * (int(height) // (vae_scale_factor * 2)), 2 * (int(width) // (vae_scale_factor * 2)))

    @staticmethod
    def _prepare_latent_image_ids(batch_size, height, width, device, dtype):
        return cond_mod.prepare_latent_image_ids(height, width, device=device, dtype=dtype)

    def prepare_latents(self, batch_size, num_channels_latents, height, width, dtype, device, generator, latents=None):
        """src/UniGenPipeline.py:662-686 (FluxPipeline.prepare_latents): noise (B, C, 2*(h//16), 2*(w//16)) -> packed tokens + ids."""
        h, w = 2 * (int(height) // (self.vae_scale_factor * 2)), 2 * (int(width) // (self.vae_scale_factor * 2))
        ids = self._prepare_latent_image_ids(batch_size, h // 2, w // 2, device, torch.float32)
        if latents is not None:
            return latents.to(device=device, dtype=dtype), ids
        gdev = generator.device if generator is not None else device
        noise = torch.randn((batch_size, num_channels_latents, h, w), generator=generator, device=gdev, dtype=torch.float32)
        return self._pack_latents(noise.to(device=device, dtype=torch.bfloat16)), ids

    @torch.no_grad()
    def __call__(self, prompt=None, prompt_2=None, condition_prompt=None, control_image=None, conditioning_scale=1.0,
                 true_cfg_scale: float = 1.0, height: Optional[int] = None, width: Optional[int] = None,
                 num_inference_steps: int = 28, sigmas: Optional[List[float]] = None, guidance_scale: float = 3.5,
                 negative_prompt=None, negative_prompt_2=None, num_images_per_prompt: int = 1, generator=None, latents=None,
                 prompt_embeds=None, condition_prompt_embeds=None, negative_prompt_embeds=None, pooled_prompt_embeds=None,
                 condition_pooled_prompt_embeds=None, negative_pooled_prompt_embeds=None, output_type: str = "latent",
                 return_dict: bool = True, joint_attention_kwargs=None, condition_ids=None, condition_types=None,
                 rts_uniform=None, graph_loop: bool = True, **kwargs):
        if prompt is not None or prompt_2 is not None or negative_prompt is not None or negative_prompt_2 is not None:
            raise ops.UgError("text encoders are outside the B200-native path: pass prompt_embeds / pooled_prompt_embeds "
                              "(and the negative_* embeddings for true CFG)")
        if prompt_embeds is None or pooled_prompt_embeds is None or control_image is None or condition_pooled_prompt_embeds is None:
            raise ValueError("prompt_embeds, pooled_prompt_embeds, control_image and condition_pooled_prompt_embeds are required")
        if isinstance(conditioning_scale, (list, tuple)):
            conditioning_scale = conditioning_scale[0]
        tr = self.transformer
        dev = self.device
        height = height or self.default_sample_size * self.vae_scale_factor
        width = width or self.default_sample_size * self.vae_scale_factor
        B = prompt_embeds.shape[0] * num_images_per_prompt
        rep = lambda t: t.repeat_interleave(num_images_per_prompt, 0) if num_images_per_prompt > 1 else t  # noqa: E731
        prompt_embeds, pooled_prompt_embeds = rep(prompt_embeds), rep(pooled_prompt_embeds)
        # ---- conditions (:930-975): packed latents + ids per condition ----
        multi = isinstance(control_image, (list, tuple))
        imgs = list(control_image) if multi else [control_image]
        types_ = list(condition_types) if condition_types is not None else [None] * len(imgs)
        given_ids = (list(condition_ids) if isinstance(condition_ids, (list, tuple)) else [condition_ids]) if condition_ids is not None \
            else [None] * len(imgs)
        cond_tokens, cond_ids = [], []
        for im, ty, gi in zip(imgs, types_, given_ids):
            if isinstance(im, cond_mod.Condition):
                tok, ids, _ = im.encode(self, generator)
            elif im.dim() == 4:  # pixel image: :954-961 (vae.encode -> sample -> shift / scale -> pack)
                tok, ids, _ = cond_mod.Condition(ty or "depth", im).encode(self, generator)
                ids = gi if gi is not None else ids
            else:
                tok = im
                ids = gi if gi is not None else (cond_mod.condition_ids(ty, height, width)[0] if ty is not None else
                                                 self._prepare_latent_image_ids(B, height // 16, width // 16, dev, torch.float32))
            cond_tokens.append(rep(tok) if tok.shape[0] != B else tok)
            cond_ids.append(ids)
        cpool = condition_pooled_prompt_embeds
        cpool = [rep(c) if c.shape[0] != B else c for c in (list(cpool) if isinstance(cpool, (list, tuple)) else [cpool])]
        if len(cpool) != len(cond_tokens):
            raise ValueError(f"{len(cond_tokens)} condition image(s) but {len(cpool)} condition_pooled_prompt_embeds")
        # ---- latents + ids (:978-987) ----
        latents, latent_image_ids = self.prepare_latents(B, tr.config.in_channels // 4, height, width, torch.bfloat16, dev, generator,
                                                         latents)
        text_ids = torch.zeros(prompt_embeds.shape[1], 3, device=dev, dtype=torch.float32)
        guidance = torch.full([B], float(guidance_scale), device=dev, dtype=torch.float32) if tr.config.guidance_embeds else None  # :1011-1015
        do_true_cfg = true_cfg_scale > 1 and negative_prompt_embeds is not None
        out = denoise(tr, latents, cond_tokens if multi else cond_tokens[0], prompt_embeds, pooled_prompt_embeds,
                      cpool if multi else cpool[0], latent_image_ids, text_ids, cond_ids if multi else cond_ids[0],
                      num_inference_steps=num_inference_steps, guidance=guidance, conditioning_scale=float(conditioning_scale),
                      rts_uniform=rts_uniform, sigmas=sigmas, true_cfg_scale=float(true_cfg_scale) if do_true_cfg else 1.0,
                      negative_encoder_hidden_states=rep(negative_prompt_embeds) if do_true_cfg else None,
                      negative_pooled_projections=rep(negative_pooled_prompt_embeds) if do_true_cfg else None, graph_loop=graph_loop)
        if output_type == "latent":
            image = out
        else:
            if self.vae_decode is None:
                raise ops.UgError("no VAE bound: use output_type='latent', or pass vae= (unigen_b200.vae.AutoencoderKL) / vae_decode=")
            image = self.vae_decode(self._unpack_latents(out, height, width, self.vae_scale_factor))
        return types.SimpleNamespace(images=image) if return_dict else (image,)


class UniGenSD3Pipeline:
    """`UniGenSD3Pipeline`-shaped entry (src/UniGenPipeline.py:145-449) for the B200-native `UniGenSD3`. Same call-site keywords
    for everything the denoise path consumes; what needs a text encoder / VAE must arrive pre-computed:

      prompt_embeds, pooled_prompt_embeds (+ negative_*)   instead of prompt / prompt_2 / prompt_3 (a string prompt raises)
      condition_pooled_prompt_embeds                        the pooled embedding of the condition prompt (:274-286)
      control_image                                         condition LATENTS (B, 16, H/8, W/8), already `(vae.encode(x) - shift) *
                                                            scaling` (:306-308), or pixels when `vae=` / `vae_encode=` is given
      output_type="latent"                                  or "pt" with `vae=` (unigen_b200.vae.AutoencoderKL, native) / `vae_decode=`

    Classifier-free guidance is on for `guidance_scale > 1` (`do_classifier_free_guidance`) and then needs the negative embeddings,
    as the reference's `encode_prompt` would produce them. Returns `SimpleNamespace(images=latents)` or `(latents,)`."""

    def __init__(self, transformer, vae_encode=None, vae_decode=None, vae_scale_factor: int = 8, default_sample_size: int = 128,
                 shift: float = 3.0, vae=None):
        self.transformer = transformer
        self.vae = vae  # unigen_b200.vae.AutoencoderKL: native encode of the condition image / decode of the final latents
        if vae is not None and vae_decode is None:
            vae_decode = vae.decode_latents
        self.vae_encode, self.vae_decode = vae_encode, vae_decode
        self.vae_scale_factor, self.default_sample_size, self.shift = vae_scale_factor, default_sample_size, shift
        self.joint_attention_kwargs = None

    @property
    def device(self):
        return self.transformer.device

    @property
    def dtype(self):
        return self.transformer.dtype

    def prepare_latents(self, batch_size, num_channels_latents, height, width, dtype, device, generator, latents=None):
        """StableDiffusion3Pipeline.prepare_latents: noise (B, C, h // 8, w // 8) unless latents are given."""
        if latents is not None:
            return latents.to(device=device, dtype=dtype)
        shape = (batch_size, num_channels_latents, int(height) // self.vae_scale_factor, int(width) // self.vae_scale_factor)
        gdev = generator.device if generator is not None else device
        return torch.randn(shape, generator=generator, device=gdev, dtype=torch.float32).to(device=device, dtype=dtype)

    @torch.no_grad()
    def __call__(self, prompt=None, prompt_2=None, prompt_3=None, condition_prompt=None, control_image=None,
                 control_use_vae_shift_factor: bool = True, conditioning_scale=1.0, height: Optional[int] = None,
                 width: Optional[int] = None, num_inference_steps: int = 28, sigmas: Optional[List[float]] = None,
                 guidance_scale: float = 7.0, control_guidance_start=0.0, control_guidance_end=1.0, negative_prompt=None,
                 negative_prompt_2=None, negative_prompt_3=None, condition_negative_prompt=None, num_images_per_prompt: int = 1,
                 generator=None, latents=None, prompt_embeds=None, condition_prompt_embeds=None, negative_prompt_embeds=None,
                 condition_negative_prompt_embeds=None, pooled_prompt_embeds=None, condition_pooled_prompt_embeds=None,
                 negative_pooled_prompt_embeds=None, condition_negative_pooled_prompt_embeds=None, ip_adapter_image=None,
                 ip_adapter_image_embeds=None, output_type: str = "latent", return_dict: bool = True, joint_attention_kwargs=None,
                 rts_uniform=None, graph_loop: bool = True, **kwargs):
        if any(p is not None for p in (prompt, prompt_2, prompt_3, negative_prompt, negative_prompt_2, negative_prompt_3)):
            raise ops.UgError("text encoders are outside the B200-native path: pass prompt_embeds / pooled_prompt_embeds "
                              "(and the negative_* embeddings for classifier-free guidance)")
        if ip_adapter_image is not None or ip_adapter_image_embeds is not None:
            raise ops.UgError("ip_adapter_image(_embeds) is not covered by the B200-native path")
        if prompt_embeds is None or pooled_prompt_embeds is None or control_image is None or condition_pooled_prompt_embeds is None:
            raise ValueError("prompt_embeds, pooled_prompt_embeds, control_image and condition_pooled_prompt_embeds are required")
        first = lambda v: v[0] if isinstance(v, (list, tuple)) else v  # noqa: E731  (:226-236: lists collapse to their first entry)
        conditioning_scale = first(conditioning_scale)
        start, end = float(first(control_guidance_start)), float(first(control_guidance_end))
        tr, dev = self.transformer, self.device
        do_cfg = guidance_scale > 1  # do_classifier_free_guidance
        if do_cfg and (negative_prompt_embeds is None or negative_pooled_prompt_embeds is None):
            raise ValueError("guidance_scale > 1 needs negative_prompt_embeds and negative_pooled_prompt_embeds")
        B = prompt_embeds.shape[0] * num_images_per_prompt
        rep = lambda t: t.repeat_interleave(num_images_per_prompt, 0) if num_images_per_prompt > 1 else t  # noqa: E731
        if control_image.dim() == 4 and control_image.shape[1] != tr.config.in_channels:
            if self.vae_encode is not None:
                control_image = self.vae_encode(control_image)  # caller-owned VAE: returns the shifted / scaled latents (:306-308)
            elif self.vae is not None:
                control_image = self.vae.encode_condition(control_image, generator=generator, use_shift_factor=control_use_vae_shift_factor)
        cond_latents = control_image if control_image.shape[0] == B else rep(control_image)
        if cond_latents.dim() != 4 or cond_latents.shape[1] != tr.config.in_channels:
            raise ops.UgError("control_image must be condition latents (B, in_channels, H/8, W/8): VAE encode is outside the "
                              "B200-native path (or pass vae_encode=)")
        # the reference takes height / width from the prepared control image (:304); here from its latents
        height, width = cond_latents.shape[2] * self.vae_scale_factor, cond_latents.shape[3] * self.vae_scale_factor
        latents = self.prepare_latents(B, tr.config.in_channels, height, width, torch.bfloat16, dev, generator, latents)
        cpool = condition_pooled_prompt_embeds if condition_pooled_prompt_embeds.shape[0] == B else rep(condition_pooled_prompt_embeds)
        out = denoise_sd3(tr, latents, cond_latents, rep(prompt_embeds), rep(pooled_prompt_embeds), cpool,
                          num_inference_steps=num_inference_steps, guidance_scale=float(guidance_scale),
                          negative_encoder_hidden_states=rep(negative_prompt_embeds) if do_cfg else None,
                          negative_pooled_projections=rep(negative_pooled_prompt_embeds) if do_cfg else None,
                          conditioning_scale=float(conditioning_scale), shift=self.shift, rts_uniform=rts_uniform, sigmas=sigmas,
                          graph_loop=graph_loop, control_guidance_start=start, control_guidance_end=end)
        if output_type == "latent":
            image = out
        else:
            if self.vae_decode is None:
                raise ops.UgError("no VAE bound: use output_type='latent', or pass vae= (unigen_b200.vae.AutoencoderKL) / vae_decode=")
            image = self.vae_decode(out)  # caller-owned: `latents / scaling_factor + shift_factor` -> vae.decode (:430-433)
        return types.SimpleNamespace(images=image) if return_dict else (image,)
