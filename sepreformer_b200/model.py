"""Drop-in ``Model`` for the reference's ``models/<name>/model.py`` (reference ``model.py:10-53``): same constructor
kwargs (``configs.yaml: config.model``), same ``state_dict`` keys, same ``forward(mixture[B, n]) -> (audio, audio_aux)``.

``forward`` is ONE call into ``libsepref_b200.so`` (``sepref_model_forward``): AudioEncoder, FeatureProjector, the separator,
OutputLayer and AudioDecoder (reference ``modules/module.py:12-35, 190-218, 237-283``) all run as hand-written sm_100a
kernels, so a waveform goes in and waveforms come out - no ``[B, F, T]`` feature tensor ever crosses the boundary.

The four auxiliary heads (``out_layer_bn`` / ``decoder_bn``, ``model.py:47-51``) only feed training losses.  They are
kept as parameters (checkpoints load) and, when ``compute_aux`` is true (the default, to return what the reference
returns), evaluated with a few plain torch ops on the per-stage outputs of the separator; inference callers
(``engine.py:165-172`` discards them) should set ``compute_aux = False``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List

import torch
from torch import nn

from . import _lib
from .separator import Separator

SHELL_KEYS = ("audio_encoder.conv1d.weight", "feature_projector.norm.weight", "feature_projector.norm.bias",
              "feature_projector.conv1d.weight", "out_layer.end_conv1x1.0.weight", "out_layer.end_conv1x1.0.bias",
              "out_layer.end_conv1x1.2.weight", "out_layer.end_conv1x1.2.bias", "audio_decoder.weight")


class _Holder(nn.Module):
    """Parameter container reproducing the key nesting of a reference layer (``<name>.<child>.weight``)."""

    def __init__(self, **children):
        super().__init__()
        for k, v in children.items():
            self.add_module(k, v)


def _output_layer(in_channels: int, out_channels: int) -> nn.Module:      # module.py:237-248 (Masking has no parameters)
    return _Holder(end_conv1x1=nn.Sequential(nn.Linear(out_channels, 4 * out_channels), nn.GLU(),
                                             nn.Linear(2 * out_channels, in_channels)))


class Model(nn.Module):
    """B200 SepReformer model with the reference's module surface.  INFERENCE ONLY (see ``Separator``)."""

    def __init__(self, num_stages: int, num_spks: int, module_audio_enc: dict, module_feature_projector: dict,
                 module_separator: dict, module_output_layer: dict, module_audio_dec: dict, per_stage_split: bool = False):
        super().__init__()
        enc, prj, out, dec = module_audio_enc, module_feature_projector, module_output_layer, module_audio_dec
        if (enc["in_channels"], enc["out_channels"], enc["kernel_size"], enc["stride"], enc["groups"], enc["bias"]) != (1, 256, 16, 4, 1, False):
            raise ValueError("the CUDA encoder is built for Conv1d(1, 256, k=16, stride=4, bias=False) (configs.yaml:33-39)")
        if prj.get("kernel_size", 1) != 1 or prj.get("bias", False) or dec.get("bias", False) or dec["kernel_size"] != 16 or dec["stride"] != 4:
            raise ValueError("unsupported feature projector / decoder configuration (configs.yaml:40-45, 88-93)")
        self.num_stages, self.num_spks = num_stages, num_spks
        feat = prj["out_channels"]
        self.audio_encoder = _Holder(conv1d=nn.Conv1d(1, 256, 16, stride=4, bias=False))
        self.feature_projector = _Holder(norm=nn.GroupNorm(1, 256, eps=1e-8), conv1d=nn.Conv1d(256, feat, 1, bias=False))
        self.separator = Separator(**module_separator, per_stage_split=per_stage_split)
        self.out_layer = _output_layer(out["in_channels"], out["out_channels"])
        self.audio_decoder = nn.ConvTranspose1d(256, 1, 16, stride=4, bias=False)
        self.out_layer_bn = nn.ModuleList([_output_layer(out["in_channels"], out["out_channels"]) for _ in range(num_stages)])
        self.decoder_bn = nn.ModuleList([nn.ConvTranspose1d(256, 1, 16, stride=4, bias=False) for _ in range(num_stages)])
        self.compute_aux = True
        self.separator._shell_source = self._shell_tensors
        self._ws: Dict[tuple, torch.Tensor] = {}

    # ------------------------------------------------------------------ weights
    def _shell_tensors(self) -> Dict[str, torch.Tensor]:
        sd = self.state_dict()
        return {k: sd[k] for k in SHELL_KEYS}

    def __setstate__(self, d):
        super().__setstate__(d)
        self.separator._shell_source = self._shell_tensors

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.separator.refresh_weights()          # shell tensors moved / converted with the rest
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.separator.refresh_weights()
        return out

    # ------------------------------------------------------------------ forward
    def output_samples(self, samples: int) -> int:
        return ((samples - 16) // 4) * 4 + 16

    def forward(self, x: torch.Tensor):
        """x: mixture ``[B, n]`` (or ``[n]``) fp32 CUDA tensor -> (``[audio_spk0, audio_spk1]`` each ``[B, n_out]``, aux)."""
        if x.dim() == 1:
            x = x[None]
        if x.dim() != 2:
            raise RuntimeError("Model expects a mixture of shape [B, n]")
        if not x.is_cuda:
            raise RuntimeError("sepreformer_b200.Model has no CPU path: input must be a CUDA tensor")
        if self.training:
            raise RuntimeError("sepreformer_b200.Model is inference-only (eval-mode BatchNorm folded into the weights, no dropout, "
                               "no autograd graph): call model.eval() first; train with the reference Model")
        if torch.is_grad_enabled() and x.requires_grad:
            raise RuntimeError("sepreformer_b200.Model builds no autograd graph: wrap the call in torch.no_grad() / inference_mode()")
        mix = x.detach().to(torch.float32).contiguous()
        B, n = mix.shape
        sep, s = self.separator, self.separator.shape_
        if sep.gemm_path < 1:
            raise RuntimeError("the model-level path needs a tensor-core gemm_path (1 or 2)")
        h = sep._handle_for(mix.device)
        lib = _lib.lib()
        S = s.num_spks
        n_out = self.output_samples(n)
        T = (n - 16) // 4 + 1
        Tp = sep.padded_frames(T)
        Td = Tp >> s.num_stages
        with torch.cuda.device(mix.device):
            okey = ("model", B, n, bool(self.compute_aux))
            static = h.static_out.get(okey) if sep.use_cuda_graph else None
            if static is not None:
                audio, stages = static
            else:
                audio = torch.empty(S, B, n_out, device=mix.device, dtype=torch.float32)
                stages = [torch.empty(B * S, s.feat, Td << i, device=mix.device, dtype=torch.float32)
                          for i in range(s.num_stages)] if self.compute_aux else []
                if sep.use_cuda_graph:           # address-stable buffers for graph replay (see Separator.use_cuda_graph)
                    if len(h.static_out) >= 4:
                        h.static_out.clear()
                    h.static_out[okey] = (audio, stages)
            if sep.use_cuda_graph:
                xin = h.static_in.get(("model", B, n))
                if xin is None:
                    if len(h.static_in) >= 4:
                        h.static_in.clear()
                    xin = h.static_in[("model", B, n)] = torch.empty_like(mix)
                if xin.data_ptr() != mix.data_ptr():
                    xin.copy_(mix)
                mix = xin
            ptrs = (C.c_void_p * s.num_stages)()
            for i in range(s.num_stages):
                ptrs[i] = stages[i].data_ptr() if self.compute_aux else None
            key = (mix.device.index, B, n, int(sep.gemm_path), int(sep.cla_fused))
            ws = self._ws.get(key)
            if ws is None:
                nbytes = lib.sepref_model_workspace_bytes(h.ptr, B, n)
                if nbytes == 0:
                    _lib.check(-2, "sepref_model_workspace_bytes")
                self._ws.clear()
                ws = self._ws[key] = torch.empty(nbytes, device=mix.device, dtype=torch.uint8)
            stream = torch.cuda.current_stream(mix.device).cuda_stream
            _lib.check(lib.sepref_model_forward(h.ptr, mix.data_ptr(), B, n, audio.data_ptr(), ptrs, ws.data_ptr(), ws.numel(),
                                                stream), "sepref_model_forward")
        sep.last_launch_count = lib.sepref_last_launch_count(h.ptr)
        audio_list = [audio[i] for i in range(S)]
        aux = self._aux_heads(mix, stages, T, n) if self.compute_aux else []
        return audio_list, aux

    def _aux_heads(self, mix, stages, T, n):
        """model.py:47-51: per stage, OutputLayer(masking=True) on the nearest-upsampled stage output, gated by the
        encoder output (Masking: ReLU(x) * skip, network.py:20-43), then that stage's decoder; cropped to the input length."""
        Fn = torch.nn.functional
        S = self.num_spks
        e = Fn.gelu(self.audio_encoder.conv1d(mix[:, None]))                    # [B, 256, T]
        out = []
        for i, st in enumerate(stages):
            y = Fn.interpolate(st, size=T)[..., :T].transpose(1, 2)             # [B*S, T, F]
            y = self.out_layer_bn[i].end_conv1x1(y).transpose(1, 2)             # [B*S, 256, T]
            B = y.shape[0] // S
            skip = e[:, None].expand(B, S, e.shape[1], e.shape[2]).reshape(B * S, e.shape[1], e.shape[2])
            y = (torch.relu(y) * skip).view(B, S, y.shape[1], y.shape[2]).transpose(0, 1)
            out.append([self.decoder_bn[i](y[j]).squeeze(1)[..., :n] for j in range(S)])
        return out

    # ------------------------------------------------------------------ pipelined host-buffer calls and the metric
    def submit_host(self, mix_host: torch.Tensor, slot: int = 0, device=None, out: torch.Tensor = None):
        """``sepref_model_submit_host``: queue H2D copy of the mixture -> kernels -> D2H copy of the waveforms in staging
        slot ``slot`` and return at once; ``wait_host(slot)`` returns the ``[S, B, n_out]`` CPU tensor."""
        if mix_host.is_cuda:
            raise RuntimeError("submit_host takes a CPU tensor")
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        mix = mix_host.to(torch.float32).contiguous()
        B, n = mix.shape
        sep = self.separator
        h = sep._handle_for(device)
        shape = (sep.shape_.num_spks, B, self.output_samples(n))
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, pin_memory=True)
        elif tuple(out.shape) != shape or out.dtype != torch.float32 or out.is_cuda or not out.is_contiguous():
            raise RuntimeError("out must be a contiguous fp32 CPU tensor of shape [S, B, n_out]")
        with torch.cuda.device(device):
            _lib.check(_lib.lib().sepref_model_submit_host(h.ptr, int(slot), mix.data_ptr(), B, n, out.data_ptr()),
                       "sepref_model_submit_host")
        sep.last_launch_count = _lib.lib().sepref_last_launch_count(h.ptr)
        h.pending[int(slot)] = (mix, out, [])
        return int(slot)

    def wait_host(self, slot: int = 0, device=None) -> torch.Tensor:
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        h = self.separator._handle_for(device)
        if int(slot) not in h.pending:
            raise RuntimeError(f"nothing submitted in slot {slot}")
        _lib.check(_lib.lib().sepref_model_wait_host(h.ptr, int(slot)), "sepref_model_wait_host")
        return h.pending.pop(int(slot))[1]

    def pit_si_snri(self, audio: torch.Tensor, targets: torch.Tensor, mixture: torch.Tensor, eps: float = 1e-15) -> torch.Tensor:
        """Batched PIT SI-SNR improvement on the device (``criterions.py:221-260``, two speakers).
        audio ``[2, B, n_out]`` (as ``forward`` stacks it), targets ``[2, B, n]``, mixture ``[B, n]`` -> ``[B, 3]``:
        best-permutation sum over speakers (dB) and its two per-speaker terms."""
        assert audio.is_cuda and targets.is_cuda and mixture.is_cuda
        audio, targets, mixture = audio.contiguous().float(), targets.contiguous().float(), mixture.contiguous().float()
        B, n = mixture.shape
        h = self.separator._handle_for(mixture.device)
        with torch.cuda.device(mixture.device):
            out = torch.empty(B, 3, device=mixture.device, dtype=torch.float32)
            st = torch.cuda.current_stream(mixture.device).cuda_stream
            _lib.check(_lib.lib().sepref_pit_sisnri(h.ptr, audio.data_ptr(), targets.data_ptr(), mixture.data_ptr(), B, n,
                                                    audio.shape[-1], float(eps), out.data_ptr(), st), "sepref_pit_sisnri")
        return out
