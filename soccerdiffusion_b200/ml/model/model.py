"""End2EndDiffusionTransformer (reference: soccer_diffusion/ml/model/model.py:16-179).

Same constructor keywords, same parameter / buffer names (checkpoints of the reference load
unchanged), same three entry points:

    forward(input_data, noisy_action_predictions, step)            model.py:150-157
    encode_input_data(input_data) -> list[Tensor]                  model.py:123-148
    forward_with_context(context, noisy_action_predictions, step)  model.py:159-179

plus the B200-native sampler entry ``sample(context, x_T, scheduler)`` that runs the whole DDIM loop
(ros.py:301-310) as one persistent kernel.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from soccerdiffusion_b200 import _lib, ops
from soccerdiffusion_b200.functional import AssembleContextFn
from soccerdiffusion_b200.ml.model.decoder import DiffusionActionGenerator
from soccerdiffusion_b200.ml.model.encoder.game_state import GameStateEncoder
from soccerdiffusion_b200.ml.model.encoder.image import (
    ImageEncoderType,
    SequenceEncoderType,
    image_sequence_encoder_factory,
)
from soccerdiffusion_b200.ml.model.encoder.imu import IMUEncoder
from soccerdiffusion_b200.ml.model.encoder.joint import JointEncoder
from soccerdiffusion_b200.ml.model.misc import StepToken


class _Plan:
    """Owner of one native ``sd_plan`` (packed denoiser weights, K/V caches, schedule tables)."""

    def __init__(self, d, heads, layers, T, J, ctx_tokens):
        cfg = _lib.PlanConfig(d, heads, layers, T, J, ctx_tokens)
        self.handle = C.c_void_p()
        _lib.check(_lib.lib().sd_plan_create(C.byref(cfg), C.byref(self.handle)), "sd_plan_create")
        self.weights_sig = None
        self.context_sig = None
        self.schedule_sig = None

    def close(self):
        if self.handle:
            handle, self.handle = self.handle, None
            _lib.check(_lib.lib().sd_plan_destroy(handle), "sd_plan_destroy")

    def __del__(self):
        # interpreter shutdown may already have torn the library binding down: that (and only that) is ignored
        try:
            self.close()
        except (AttributeError, TypeError, ImportError):
            pass


def _sig(tensors):
    """Identity of the CONTENTS of ``tensors`` as far as the host can know it, or None when it cannot.

    libsd_b200 writes through raw pointers, which never bumps ``Tensor._version``, and the caching allocator hands the
    same addresses out again tick after tick — so (pointer, version, shape) alone would call two different contexts
    equal.  Tensors produced by this package carry a process-wide generation stamp (``runtime.stamp``); weights carry
    the optimizer generation (``runtime.weights_generation``, advanced by every FusedAdamW step / graph replay).
    A tensor without a stamp (a caller's own ``torch.randn`` context, train.py:222) gives None: never cached."""
    from soccerdiffusion_b200 import runtime

    out = []
    for t in tensors:
        gen = getattr(t, "_sd_gen", None)
        if gen is None:
            return None
        out.append((t.data_ptr(), t._version, tuple(t.shape), gen))
    return tuple(out) + (runtime.weights_generation(),)


def _weights_sig(tensors):
    from soccerdiffusion_b200 import runtime

    return tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in tensors) + (runtime.weights_generation(),)


class End2EndDiffusionTransformer(nn.Module):
    def __init__(
        self,
        num_joints: int,
        hidden_dim: int,
        use_action_history: bool,
        num_action_history_encoder_layers: int,
        max_action_context_length: int,
        encoder_patch_size: int,
        use_imu: bool,
        imu_orientation_embedding_method: IMUEncoder.OrientationEmbeddingMethod,
        num_imu_encoder_layers: int,
        imu_context_length: int,
        use_joint_states: bool,
        joint_state_encoder_layers: int,
        joint_state_context_length: int,
        use_images: bool,
        image_encoder_type: ImageEncoderType,
        image_sequence_encoder_type: SequenceEncoderType,
        num_image_sequence_encoder_layers: int,
        image_context_length: int,
        image_use_final_avgpool: bool,
        image_resolution: int,
        use_gamestate: bool,
        num_decoder_layers: int,
        trajectory_prediction_length: int,
    ):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.num_joints = num_joints

        # diffusion-step token (model.py:46)
        self.step_encoding = StepToken(hidden_dim)

        self.action_history_encoder = (
            JointEncoder(
                num_joints=num_joints,
                patch_size=encoder_patch_size,
                hidden_dim=hidden_dim,
                num_layers=num_action_history_encoder_layers,
                num_heads=4,
                max_seq_len=max_action_context_length,
            )
            if use_action_history
            else None
        )
        self.imu_encoder = (
            IMUEncoder(
                orientation_embedding_method=imu_orientation_embedding_method,
                patch_size=encoder_patch_size,
                hidden_dim=hidden_dim,
                num_layers=num_imu_encoder_layers,
                num_heads=4,
                max_seq_len=imu_context_length,
            )
            if use_imu
            else None
        )
        self.joint_states_encoder = (
            JointEncoder(
                num_joints=num_joints,
                patch_size=encoder_patch_size,
                hidden_dim=hidden_dim,
                num_layers=joint_state_encoder_layers,
                num_heads=4,
                max_seq_len=joint_state_context_length,
            )
            if use_joint_states
            else None
        )
        self.image_sequence_encoder = (
            image_sequence_encoder_factory(
                encoder_type=SequenceEncoderType(getattr(image_sequence_encoder_type, "value", image_sequence_encoder_type)),
                image_encoder_type=ImageEncoderType(getattr(image_encoder_type, "value", image_encoder_type)),
                hidden_dim=hidden_dim,
                num_layers=num_image_sequence_encoder_layers,
                max_seq_len=image_context_length,
                use_final_avgpool=image_use_final_avgpool,
                resolution=image_resolution,
            )
            if use_images
            else None
        )
        self.game_state_encoder = GameStateEncoder(hidden_dim) if use_gamestate else None

        self.diffusion_action_generator = DiffusionActionGenerator(
            num_joints=num_joints,
            hidden_dim=hidden_dim,
            num_layers=num_decoder_layers,
            num_heads=4,
            max_seq_len=trajectory_prediction_length,
        )

        # normalisation parameters travel with the checkpoint (model.py:120-121; train.py:142-143)
        self.register_buffer("mean", torch.zeros(num_joints))
        self.register_buffer("std", torch.ones(num_joints))

        self._plans: dict[int, _Plan] = {}
        self._tc_graphs: dict = {}   # tensor-core sampler: captured DDIM loops by input signature

    # ----------------------------------------------------------------------------------------
    def encode_input_data(self, input_data: dict[str, torch.Tensor]) -> list[torch.Tensor]:
        """Runs the enabled encoders; returns their token tensors in the reference's fixed order (model.py:123-148).

        The encoders share no data, so on a CUDA device the sequence encoders are launched on side streams beside the
        image path (which stays on the caller's stream) and joined before returning; autograd replays every node on
        the stream of its forward, so the backward passes overlap the same way.  Works inside CUDA-graph capture
        (the side streams fork from / join into the capturing stream)."""
        from soccerdiffusion_b200 import runtime

        jobs = []   # (encoder, input, stays on the caller's stream)
        if self.action_history_encoder is not None:
            jobs.append((self.action_history_encoder, input_data["joint_command_history"], False))
        if self.imu_encoder is not None:
            jobs.append((self.imu_encoder, input_data["rotation"], False))
        if self.joint_states_encoder is not None:
            jobs.append((self.joint_states_encoder, input_data["joint_state"], False))
        if self.image_sequence_encoder is not None:
            if "image_tokens" in input_data and "image_data" not in input_data:
                # per-frame embeddings computed when the frames arrived (FrameEmbeddingCache; the TODO at
                # ml/inference/ros.py:180-183): only the frame-sequence encoder runs per tick
                jobs.append((self.image_sequence_encoder.transformer_encoder, input_data["image_tokens"], True))
            else:
                jobs.append((self.image_sequence_encoder, input_data["image_data"], True))
        if self.game_state_encoder is not None:
            jobs.append((self.game_state_encoder, input_data["game_state"], True))
        n_side = sum(1 for _, x, main in jobs if not main)
        if not (runtime.concurrent_encoders() and n_side >= 1 and len(jobs) >= 2 and all(x.is_cuda for _, x, _ in jobs)):
            return [runtime.stamp(enc(x)) for enc, x, _ in jobs]
        cur = torch.cuda.current_stream()
        side = runtime.side_streams(cur.device, n_side)
        out: list = [None] * len(jobs)
        k = 0
        for i, (enc, x, main) in enumerate(jobs):
            if main:
                continue
            s = side[k]
            k += 1
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                out[i] = enc(x)
        for i, (enc, x, main) in enumerate(jobs):
            if main:
                out[i] = enc(x)
        for s in side[:n_side]:
            cur.wait_stream(s)
        for i, (_, _, main) in enumerate(jobs):
            if not main:
                out[i].record_stream(cur)   # allocated on a side stream, consumed on the caller's
        return [runtime.stamp(o) for o in out]

    def forward(
        self, input_data: dict[str, torch.Tensor], noisy_action_predictions: torch.Tensor, step: torch.Tensor
    ) -> torch.Tensor:
        context = self.encode_input_data(input_data)
        return self.forward_with_context(context, noisy_action_predictions, step)

    def forward_with_context(
        self, context: list[torch.Tensor], noisy_action_predictions: torch.Tensor, step: torch.Tensor
    ) -> torch.Tensor:
        x = noisy_action_predictions
        _lib.require_cuda(x, self.step_encoding.token, *context)
        if self._use_fused_inference(context, x):
            return self._denoise_fused(context, x, step)
        from soccerdiffusion_b200.ml.model.misc import normalize_steps

        steps = normalize_steps(step, x.device)
        mem = AssembleContextFn.apply(steps, self.step_encoding.freqs, self.step_encoding.token, self.hidden_dim,
                                      *[c.float() for c in context])
        return self.diffusion_action_generator(x, mem)

    # ----------------------------------------------------------------------------------------
    # fused inference path (no autograd, no dropout): persistent sampler kernel + cached context K/V
    def _use_fused_inference(self, context, x) -> bool:
        if torch.is_grad_enabled() and (any(c.requires_grad for c in context) or x.requires_grad
                                        or any(p.requires_grad for p in self.diffusion_action_generator.parameters())):
            return False
        if self.training:
            return False  # dropout is live in train mode (reference never calls .eval() while training)
        from soccerdiffusion_b200 import runtime

        if runtime.get_precision() != ops.PREC_FP32:
            return False
        return x.shape[1] <= 32

    def _plan_for(self, ctx_tokens: int, T: int) -> _Plan:
        key = (ctx_tokens, T)
        plan = self._plans.get(key)
        dag = self.diffusion_action_generator
        if plan is None:
            plan = _Plan(self.hidden_dim, dag.num_heads, len(dag.transformer_decoder.layers), T, self.num_joints, ctx_tokens)
            self._plans[key] = plan
        tensors = [dag.embedding.weight, dag.embedding.bias, dag.fc_out.weight, dag.fc_out.bias,
                   self.step_encoding.token, self.mean, self.std, *dag.transformer_decoder.tensors()]
        sig = _weights_sig(tensors)
        if plan.weights_sig != sig:
            st = _lib.stream_ptr()
            lib = _lib.lib()
            for i, layer in enumerate(dag.transformer_decoder.layers):
                w = _lib.DecoderLayerWeights()
                sa, ca = layer.self_attn, layer.multihead_attn
                w.sa_in_w, w.sa_in_b = sa.in_proj_weight.data_ptr(), sa.in_proj_bias.data_ptr()
                w.sa_out_w, w.sa_out_b = sa.out_proj.weight.data_ptr(), sa.out_proj.bias.data_ptr()
                w.ca_in_w, w.ca_in_b = ca.in_proj_weight.data_ptr(), ca.in_proj_bias.data_ptr()
                w.ca_out_w, w.ca_out_b = ca.out_proj.weight.data_ptr(), ca.out_proj.bias.data_ptr()
                w.lin1_w, w.lin1_b = layer.linear1.weight.data_ptr(), layer.linear1.bias.data_ptr()
                w.lin2_w, w.lin2_b = layer.linear2.weight.data_ptr(), layer.linear2.bias.data_ptr()
                w.norm1_w, w.norm1_b = layer.norm1.weight.data_ptr(), layer.norm1.bias.data_ptr()
                w.norm2_w, w.norm2_b = layer.norm2.weight.data_ptr(), layer.norm2.bias.data_ptr()
                w.norm3_w, w.norm3_b = layer.norm3.weight.data_ptr(), layer.norm3.bias.data_ptr()
                _lib.check(lib.sd_plan_set_layer(plan.handle, i, C.byref(w), st), "sd_plan_set_layer")
            pe = dag.positional_encoding.table(T).contiguous()
            _lib.check(lib.sd_plan_set_io(plan.handle, dag.embedding.weight.data_ptr(), dag.embedding.bias.data_ptr(),
                                          dag.fc_out.weight.data_ptr(), dag.fc_out.bias.data_ptr(), pe.data_ptr(),
                                          self.step_encoding.freqs.data_ptr(), self.step_encoding.token.data_ptr(),
                                          self.mean.data_ptr(), self.std.data_ptr(), st), "sd_plan_set_io")
            ops._count(20 * len(dag.transformer_decoder.layers))
            plan.weights_sig = sig
            plan.context_sig = None
            plan.schedule_sig = None
        return plan

    def _set_context(self, plan: _Plan, context: list[torch.Tensor]):
        sig = _sig(context)
        if sig is not None and plan.context_sig == sig:
            return
        B = context[0].shape[0]
        for c in context:
            if c.shape[0] != B:
                raise RuntimeError(f"context tensors disagree on the batch size: {[tuple(c.shape) for c in context]}")
        d = self.hidden_dim
        lens = [c.shape[1] for c in context]
        Mc = sum(lens)
        if len(context) == 1 and context[0].is_contiguous() and context[0].dtype == torch.float32:
            ctx = context[0]
        else:
            ctx = torch.empty((B, Mc, d), device=context[0].device, dtype=torch.float32)
            off = 0
            for c, n in zip(context, lens):
                cc = c.float().contiguous()
                ops.copy_rows(cc.data_ptr(), n * d, d, ctx.data_ptr() + 4 * off * d, Mc * d, d, B, n, d)
                off += n
        _lib.check(_lib.lib().sd_plan_set_context(plan.handle, ctx.data_ptr(), B, _lib.stream_ptr()),
                   "sd_plan_set_context")
        ops._count(2)
        plan.context_sig = sig
        plan.context_B = B

    def _denoise_fused(self, context, x, step):
        from soccerdiffusion_b200.ml.model.misc import normalize_steps

        B, T, J = x.shape
        if context[0].shape[0] != B:
            raise RuntimeError(f"batch size of the context ({context[0].shape[0]}) and of the noisy actions ({B}) differ")
        plan = self._plan_for(sum(c.shape[1] for c in context), T)
        self._set_context(plan, context)
        steps = normalize_steps(step, x.device)
        if steps.numel() == 1 and B > 1:
            steps = steps.expand(B).contiguous()
        xin = x.float().contiguous()
        out = torch.empty_like(xin)
        _lib.check(_lib.lib().sd_plan_denoise(plan.handle, xin.data_ptr(), steps.data_ptr(),
                                              1 if steps.dtype == torch.float32 else 0, out.data_ptr(), B,
                                              _lib.stream_ptr()), "sd_plan_denoise")
        ops._count()
        return out

    # ----------------------------------------------------------------------------------------
    # tensor-core batched sampler (bf16 mode): the DDIM loop on the layer-fused kernels, replayed from one CUDA graph
    def tc_sampler_supported(self, context_tokens: int, T: int) -> bool:
        from soccerdiffusion_b200.functional import tc_sampler_supported

        dag = self.diffusion_action_generator
        layers = dag.transformer_decoder.layers
        return tc_sampler_supported(self.hidden_dim, dag.num_heads, T, context_tokens + 1, len(layers),
                                    layers[0].linear1.weight.shape[0] if len(layers) else 0)

    def _sample_tc(self, context, x_T, scheduler, denormalize: bool, return_trace: bool, use_graph: bool = True):
        from soccerdiffusion_b200.functional import ddim_sample_tc

        B, T, J = x_T.shape
        d = self.hidden_dim
        dag = self.diffusion_action_generator
        lens = tuple(int(c.shape[1]) for c in context)
        Mm = sum(lens) + 1
        ts, coef = scheduler.schedule_tables()
        S = len(ts)
        dev = x_T.device
        key = (B, T, J, lens, tuple(ts), coef.tobytes(), bool(denormalize), bool(return_trace), dev.index)
        ent = self._tc_graphs.get(key) if use_graph else None
        tsd = torch.tensor(ts, device=dev, dtype=torch.int64) if ent is None else ent[4]

        def run(ctx_list, x_in):
            mem = torch.empty((B, Mm, d), device=dev, dtype=torch.float32)
            off = 0
            for c, n in zip(ctx_list, lens):
                cc = c.float().contiguous()
                ops.copy_rows(cc.data_ptr(), n * d, d, mem.data_ptr() + 4 * off * d, Mm * d, d, B, n, d)
                off += n
            ops.copy_rows(self.step_encoding.token.data_ptr(), 0, d, mem.data_ptr() + 4 * off * d, Mm * d, d, B, 1, d)  # placeholder row
            tok = torch.empty((S, d), device=dev, dtype=torch.float32)
            ops.step_token(tsd, self.step_encoding.freqs, self.step_encoding.token, tok, d, S, d)
            trace = torch.empty((S, B, T, J), device=dev, dtype=torch.float32) if return_trace else None
            x0 = ddim_sample_tc(B, T, Mm, dag.num_heads, dag.positional_encoding.table(T).contiguous(), x_in.float(),
                                mem.view(B * Mm, d), tok, dag.embedding.weight, dag.embedding.bias, dag.fc_out.weight,
                                dag.fc_out.bias, dag.transformer_decoder.tensors(), [tuple(map(float, c)) for c in coef], trace)
            if denormalize:
                out = torch.empty_like(x0)
                ops.affine_joints(x0.contiguous(), self.mean.contiguous(), self.std.contiguous(), out, 1)
                x0 = out
            return x0, trace

        if not use_graph or torch.cuda.is_current_stream_capturing():
            x0, trace = run(context, x_T)   # kernel by kernel (inside a caller's capture: part of the caller's graph)
            return (x0, trace) if return_trace else x0
        if ent is None:
            if len(self._tc_graphs) >= 4:
                self._tc_graphs.pop(next(iter(self._tc_graphs)))
            s_ctx = [torch.empty((B, n, d), device=dev, dtype=torch.float32) for n in lens]
            s_x = torch.empty((B, T, J), device=dev, dtype=torch.float32)
            for dst, c in zip(s_ctx, context):
                dst.copy_(c)
            s_x.copy_(x_T)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                run(s_ctx, s_x)   # warm-up outside the capture (lazy kernel attributes, allocator)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = run(s_ctx, s_x)
            ent = self._tc_graphs[key] = (g, s_ctx, s_x, out, tsd)
        g, s_ctx, s_x, (x0, trace), _ = ent
        for dst, c in zip(s_ctx, context):
            dst.copy_(c)
        s_x.copy_(x_T)
        g.replay()
        x0 = x0.clone()
        return (x0, trace.clone()) if return_trace else x0

    @torch.no_grad()
    def sample(self, context: list[torch.Tensor], x_T: torch.Tensor, scheduler, num_inference_steps: int | None = None,
               denormalize: bool = False, return_trace: bool = False, sampler: str = "auto"):
        """Runs the complete DDIM loop of ros.py:301-310 / distill.py:179-189.

        ``scheduler`` is a ``soccerdiffusion_b200.schedulers.DDIMScheduler`` whose ``set_timesteps`` has
        been called (or pass ``num_inference_steps``).  Returns x_0 (denormalised like ros.py:313 if asked).

        ``sampler``: "cluster" / "cta" = the fp32 persistent kernels (one launch for the whole loop; 16-CTA cluster or one
        CTA per trajectory), "tc" = the bf16 tensor-core path (layer-fused tcgen05 kernels over the whole batch, context
        K | V projected once, the loop replayed from one CUDA graph), "auto" = "tc" in bf16 precision mode for batches
        of ``runtime.tc_sampler_min_batch()`` trajectories or more when the shapes are supported, else the fp32 kernels.
        """
        _lib.require_cuda(x_T, *context)
        if self.training:
            raise RuntimeError("sample() implements eval-mode semantics; call model.eval() first")
        if num_inference_steps is not None:
            scheduler.set_timesteps(num_inference_steps)
        B, T, J = x_T.shape
        if context[0].shape[0] != B:
            raise RuntimeError(f"batch size of the context ({context[0].shape[0]}) and of x_T ({B}) differ")
        for c in context:
            if c.shape[0] != B:
                raise RuntimeError(f"context tensors disagree on the batch size: {[tuple(c.shape) for c in context]}")
        if sampler not in ("auto", "cta", "cluster", "tc"):
            raise ValueError(f"unknown sampler {sampler!r}")
        from soccerdiffusion_b200 import runtime

        use_tc = sampler == "tc"
        n_layers = len(self.diffusion_action_generator.transformer_decoder.layers)
        # measured on a B200 (tools/sampler_micro.py): one trajectory of the default architecture is fastest on the 16-CTA
        # cluster kernel (3.4 ms vs 4.4 ms); batches, and deeper / longer decoders even at one trajectory (scaled-up
        # config: 8.6 ms vs 15.9 ms), on the tensor-core path
        if sampler == "auto" and runtime.get_precision() == ops.PREC_BF16 and (B >= runtime.tc_sampler_min_batch()
                                                                               or n_layers * T >= 80):
            use_tc = self.tc_sampler_supported(sum(c.shape[1] for c in context), T)
        if use_tc:
            if not self.tc_sampler_supported(sum(c.shape[1] for c in context), T):
                raise _lib.SdError("sampler='tc': shapes outside the fused tensor-core kernels (d=128, 4 heads, T<=64, memory<=384)")
            self.last_sampler = "tc"
            return self._sample_tc(context, x_T, scheduler, denormalize, return_trace)
        plan = self._plan_for(sum(c.shape[1] for c in context), T)
        _lib.check(_lib.lib().sd_plan_set_sampler(plan.handle, {"auto": 0, "cta": 1, "cluster": 2}[sampler]),
                   "sd_plan_set_sampler")
        self._set_context(plan, context)
        ts, coef = scheduler.schedule_tables()
        ssig = (tuple(ts), coef.tobytes())
        if plan.schedule_sig != ssig:
            n = len(ts)
            arr_t = (C.c_longlong * n)(*ts)
            arr_c = (C.c_float * (4 * n))(*coef.reshape(-1).tolist())
            _lib.check(_lib.lib().sd_plan_set_schedule(plan.handle, n, arr_t, arr_c, _lib.stream_ptr()),
                       "sd_plan_set_schedule")
            ops._count(2)
            plan.schedule_sig = ssig
        xin = x_T.float().contiguous()
        out = torch.empty_like(xin)
        trace = torch.empty((len(ts), B, T, J), device=xin.device, dtype=torch.float32) if return_trace else None
        _lib.check(_lib.lib().sd_plan_sample(plan.handle, xin.data_ptr(), out.data_ptr(), _lib.ptr(trace),
                                             1 if denormalize else 0, B, _lib.stream_ptr()), "sd_plan_sample")
        ops._count()
        self.last_sampler = {1: "cta", 2: "cluster"}.get(_lib.lib().sd_plan_last_sampler(plan.handle), "?")
        return (out, trace) if return_trace else out
