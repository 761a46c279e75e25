"""Drop-in ``Separator`` for the reference's ``Model`` (reference ``modules/module.py:38-234``).

Same constructor kwargs (``configs.yaml:46-83``), same ``state_dict`` keys and shapes, same
``forward(input[B,F,L]) -> (last[B*S,F,L_pad], [stage outputs])`` contract - but ``forward`` is a single
call into ``libsepref_b200.so`` (hand-written sm_100a kernels behind the C ABI of ``include/sepref.h``).
Host code stays Python/PyTorch: torch provides device memory, the stream and the parameter storage only.

There is no CPU path: a CPU tensor, or a machine without the built library, raises.
"""
from __future__ import annotations

import ctypes as C
import threading
import weakref
from typing import Dict, List, Tuple

import torch

from . import _lib
from .configs import SeparatorShape, shape_from_kwargs
from .params import ParamTree, separator_spec


class _Handle:
    """Owns one ``sepref_handle`` (one per CUDA device)."""

    def __init__(self, shape: SeparatorShape, device_index: int):
        L = _lib.lib()
        cfg = _lib.SeprefConfig(shape.feat, shape.heads, shape.num_stages, shape.num_spks, shape.cla_kernel,
                                shape.down_kernel, shape.maxlen, int(shape.per_stage_split))
        self.ptr = C.c_void_p()
        _lib.check(L.sepref_create(C.byref(cfg), device_index, C.byref(self.ptr)), "sepref_create")
        self.version = None
        self.workspaces: Dict[Tuple[int, int], torch.Tensor] = {}
        self.static_out: Dict[tuple, tuple] = {}      # CUDA-graph mode: address-stable outputs per shape
        self.static_in: Dict[tuple, torch.Tensor] = {}
        self.pending: Dict[int, tuple] = {}       # submit_host requests in flight, by slot

    def __del__(self):
        try:
            if self.ptr:
                _lib.lib().sepref_destroy(self.ptr)
                self.ptr = C.c_void_p()
        except Exception:
            pass

    def load(self, state: Dict[str, torch.Tensor], shell: Dict[str, torch.Tensor] = None):
        """``state``: the separator's state_dict; ``shell`` (optional): the model-shell tensors by their key in
        ``Model.state_dict()`` - handed to the library under "@" + key (include/sepref.h)."""
        L = _lib.lib()
        items = list(state.items()) + [("@" + k, v) for k, v in (shell or {}).items()]
        for key, t in items:
            if not t.is_floating_point():
                continue       # BatchNorm.num_batches_tracked
            host = t.detach().to(device="cpu", dtype=torch.float32).contiguous()
            shp = (C.c_int64 * host.dim())(*host.shape)
            _lib.check(L.sepref_set_param(self.ptr, key.encode(), host.data_ptr(), shp, host.dim()),
                       f"sepref_set_param({key})")
        _lib.check(L.sepref_finalize(self.ptr), "sepref_finalize")


def _refresh_after_load(module, incompatible_keys):
    module.refresh_weights()


class _Shared:
    """State that ``torch.nn.parallel.replicate`` must NOT duplicate: replicas are shallow copies of the module's
    ``__dict__`` (``Module._replicate_for_data_parallel``), so this one object - the packed per-device handles, their
    lock, the weight epoch and a weak reference to the module that owns the real parameters - is shared by the master
    and every replica of every ``data_parallel`` call (engine.py:64,98,130,167 with several ``gpuid``s)."""

    def __init__(self, owner: "Separator"):
        self.master = weakref.ref(owner)
        self.handles: Dict[int, _Handle] = {}
        self.lock = threading.Lock()
        self.epoch = 0            # bumped whenever the owner's weights may have changed
        self.packs = 0            # how many times weights were packed and uploaded (tests: no re-pack per call)


class Separator(ParamTree):
    """B200 separator with the reference's module surface.  INFERENCE ONLY: ``forward`` runs eval-mode kernels
    (BatchNorm folded, no dropout) and builds no autograd graph, so it refuses to run in training mode."""

    def __init__(self, num_stages: int, relative_positional_encoding: dict, enc_stage: dict, spk_split_stage: dict,
                 simple_fusion: dict, dec_stage: dict, per_stage_split: bool = False):
        shape = shape_from_kwargs(num_stages, relative_positional_encoding, enc_stage, spk_split_stage,
                                  simple_fusion, dec_stage, per_stage_split)
        super().__init__(separator_spec(shape))
        self.shape_ = shape
        self.num_stages = num_stages
        self._shared = _Shared(self)
        # a parent's load_state_dict() fills this module through _load_from_state_dict, not through our own
        # load_state_dict(): the post-hook (called for every sub-module of the recursion) catches that case
        self.register_load_state_dict_post_hook(_refresh_after_load)
        self._shell_source = None     # set by sepreformer_b200.Model: callable returning the model-shell tensors
        # "hooks": weights are re-packed after load_state_dict / .to() / .cuda() / refresh_weights();
        # "always": additionally compare (data_ptr, _version) of every tensor on each call (~3 ms of host time)
        self.check_weights = "hooks"
        # 2 = tcgen05 kind::f16 (fp16 operands, same 11-bit significand as TF32, fp32 accumulate; default),
        # 1 = tcgen05 kind::tf32, 0 = fp32 CUDA-core kernels
        self.gemm_path = 2
        self.debug_sync = False
        self.cluster = 2              # CTAs sharing each TMA-multicast weight slab (1, 2 or 4)
        self.gcfn_wide = 0            # 1: 160-frame GCFN tiles with single-buffered accumulators (f16 path, F = 128)
        self.gcfn_tm = 0              # 1: frames-as-M GCFN kernel on CTA pairs (cta_group::2), 2: on single CTAs (f16 path, F = 128)
        self.gcfn_pair = 0            # 1: weights-resident CTA-pair GCFN kernel (f16 path, F = 128; measured slower, see profiles/r2_gcfn_pair.md)
        # True: replay a captured CUDA graph instead of ~260 launch calls per forward (SEPREF_OPT_CUDA_GRAPH).  The graph is
        # tied to buffer addresses, so forward() then returns the SAME output tensors on every call of a given shape:
        # consume (or clone) them before calling forward again.
        self.use_cuda_graph = False
        self.cla_fused = 1            # 0: CLA as three kernels with fp32 intermediates (the round-1 schedule)
        self.gcfn_trio = 0            # 1: weights-resident GCFN kernel on clusters of three CTAs (f16 path, F = 128)
        self.raw_f16 = 0              # 0: raw-stream GEMMs = FP16 + run-time range check + conditional TF32 re-computation; 1: FP16 only
        self.write_stage_outputs = True   # the four auxiliary outputs only feed training-time heads (model.py:47-51)
        self.last_launch_count = 0

    # ------------------------------------------------------------------ packing
    def refresh_weights(self):
        """Call after modifying parameters in place (``p.data.copy_``...): the next forward re-packs them."""
        self._sh().epoch += 1

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.refresh_weights()
        return out

    def __getstate__(self):
        d = dict(self.__dict__)
        d.pop("_shared", None)          # ctypes handles do not pickle; they are rebuilt on first use
        d["_shell_source"] = None       # a bound method of the parent model; the parent re-attaches it
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self._shared = _Shared(self)

    def _sh(self) -> _Shared:
        sh = self.__dict__.get("_shared")
        # copy.deepcopy / pickling give a real module whose shared state still points at the original: start afresh
        if sh is None or (not getattr(self, "_is_replica", False) and sh.master() is not self):
            sh = self._shared = _Shared(self)
        return sh

    def _weights_version(self, owner):
        return tuple((t.data_ptr(), t._version) for t in owner.state_dict(keep_vars=True).values())

    def _handle_for(self, device: torch.device) -> _Handle:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        sh = self._sh()
        # data_parallel replicas hold broadcast copies of the weights with an empty _parameters dict; the weights that
        # matter are the owner's (identical by construction), packed once per device and kept across calls
        owner = sh.master() or self
        with sh.lock:
            h = sh.handles.get(idx)
            if h is None:
                h = sh.handles[idx] = _Handle(self.shape_, idx)
            ver = (sh.epoch, self._weights_version(owner) if self.check_weights == "always" else None)
            if h.version != ver:
                src = owner.__dict__.get("_shell_source")
                h.load(owner.state_dict(), src() if src is not None else None)
                h.version = ver
                sh.packs += 1
        L = _lib.lib()
        _lib.check(L.sepref_set_option(h.ptr, _lib.OPT_GEMM_PATH, int(self.gemm_path)))
        _lib.check(L.sepref_set_option(h.ptr, _lib.OPT_DEBUG_SYNC, int(self.debug_sync)))
        _lib.check(L.sepref_set_option(h.ptr, _lib.OPT_CLUSTER, int(self.cluster)))
        _lib.check(L.sepref_set_option(h.ptr, _lib.OPT_GCFN_WIDE, int(self.gcfn_wide)))
        _lib.check(L.sepref_set_option(h.ptr, _lib.OPT_RAW_F16, int(self.raw_f16)))
        _lib.check(L.sepref_set_option(h.ptr, _lib.OPT_GCFN_PAIR, int(self.gcfn_pair)))
        _lib.check(L.sepref_set_option(h.ptr, _lib.OPT_GCFN_TM, int(getattr(self, "gcfn_tm", 0))))
        _lib.check(L.sepref_set_option(h.ptr, _lib.OPT_GCFN_TRIO, int(self.gcfn_trio)))
        _lib.check(L.sepref_set_option(h.ptr, _lib.OPT_CLA_FUSED, int(self.cla_fused)))
        _lib.check(L.sepref_set_option(h.ptr, _lib.OPT_CUDA_GRAPH, int(bool(self.use_cuda_graph))))
        return h

    def handle(self, device=None) -> "C.c_void_p":
        """The raw ``sepref_handle*`` for ``device`` (weights packed), for block-level calls."""
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        return self._handle_for(device).ptr

    # ------------------------------------------------------------------ forward
    def padded_frames(self, frames: int) -> int:
        chunk = self.shape_.chunk
        return frames if frames % chunk == 0 else (frames // chunk + 1) * chunk

    def forward(self, input: torch.Tensor):
        """input: [B, F, L] fp32 CUDA tensor (module.py:190-218)."""
        if input.dim() != 3:
            raise RuntimeError("Separator expects [B, F, L] features")
        s = self.shape_
        if input.shape[1] != s.feat:
            raise RuntimeError(f"expected {s.feat} feature channels, got {input.shape[1]}")
        if not input.is_cuda:
            raise RuntimeError("sepreformer_b200.Separator has no CPU path: input must be a CUDA tensor")
        if self.training:
            raise RuntimeError("sepreformer_b200.Separator is inference-only (eval-mode BatchNorm folded into the weights, "
                               "no dropout, no autograd graph): call model.eval() first; train with the reference Separator")
        if torch.is_grad_enabled() and input.requires_grad:
            raise RuntimeError("sepreformer_b200.Separator builds no autograd graph: its input requires grad, so gradients "
                               "would silently stop here; wrap the call in torch.no_grad() / inference_mode() or detach the input")
        x = input.detach().to(torch.float32).contiguous()
        B, F, L = x.shape
        Tp = self.padded_frames(L)
        Td = Tp >> s.num_stages
        h = self._handle_for(x.device)
        lib = _lib.lib()
        with torch.cuda.device(x.device):
            okey = (B, L, bool(self.write_stage_outputs))
            static = h.static_out.get(okey) if self.use_cuda_graph else None
            if static is not None:
                last, stages = static
            else:
                last = torch.empty(B * s.num_spks, F, Tp, device=x.device, dtype=torch.float32)
                stages = [torch.empty(B * s.num_spks, F, Td << i, device=x.device, dtype=torch.float32)
                          for i in range(s.num_stages)] if self.write_stage_outputs else []
                if self.use_cuda_graph:          # address-stable outputs (and input staging) for graph replay
                    if len(h.static_out) >= 4:
                        h.static_out.clear()
                    h.static_out[okey] = (last, stages)
            if self.use_cuda_graph:
                xin = h.static_in.get((B, L))
                if xin is None:
                    if len(h.static_in) >= 4:
                        h.static_in.clear()
                    xin = h.static_in[(B, L)] = torch.empty_like(x)
                if xin.data_ptr() != x.data_ptr():
                    xin.copy_(x)
                x = xin
            stage_ptrs = (C.c_void_p * s.num_stages)()
            for i in range(s.num_stages):
                stage_ptrs[i] = stages[i].data_ptr() if self.write_stage_outputs else None
            key = (B, L, int(self.gemm_path), int(self.cla_fused))
            ws = h.workspaces.get(key)
            if ws is None:
                nbytes = lib.sepref_workspace_bytes(h.ptr, B, L)
                h.workspaces.clear()
                ws = h.workspaces[key] = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(lib.sepref_separator_forward(h.ptr, x.data_ptr(), B, L, last.data_ptr(), stage_ptrs,
                                                    ws.data_ptr(), ws.numel(), stream), "sepref_separator_forward")
        self.last_launch_count = lib.sepref_last_launch_count(h.ptr)
        return last, stages

    def forward_host(self, x_host: torch.Tensor, device=None, want_stages: bool = False, out: torch.Tensor = None):
        """Same computation through the HOST-buffer C-ABI entry (``sepref_separator_forward_host``):
        pinned/pageable CPU tensor in, pinned CPU tensors out, copies and a stream sync included.
        ``out`` (optional) is a reusable pinned ``[B*S, F, L_pad]`` result buffer - allocating 100s of MB of pinned
        memory per call costs more than the copy itself."""
        if x_host.is_cuda:
            raise RuntimeError("forward_host takes a CPU tensor")
        s = self.shape_
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        x = x_host.to(torch.float32).contiguous()
        B, F, L = x.shape
        Tp = self.padded_frames(L)
        Td = Tp >> s.num_stages
        h = self._handle_for(device)
        lib = _lib.lib()
        if out is None:
            out = torch.empty(B * s.num_spks, F, Tp, dtype=torch.float32, pin_memory=True)
        elif tuple(out.shape) != (B * s.num_spks, F, Tp) or out.dtype != torch.float32 or out.is_cuda or not out.is_contiguous():
            raise RuntimeError("out must be a contiguous fp32 CPU tensor of shape [B*S, F, L_pad]")
        stages, ptrs = [], (C.c_void_p * s.num_stages)()
        for i in range(s.num_stages):
            if want_stages:
                t = torch.empty(B * s.num_spks, F, Td << i, dtype=torch.float32, pin_memory=True)
                stages.append(t)
                ptrs[i] = t.data_ptr()
            else:
                ptrs[i] = None
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            _lib.check(lib.sepref_separator_forward_host(h.ptr, x.data_ptr(), B, L, out.data_ptr(), ptrs, stream),
                       "sepref_separator_forward_host")
        self.last_launch_count = lib.sepref_last_launch_count(h.ptr)
        return out, stages

    def submit_host(self, x_host: torch.Tensor, slot: int = 0, device=None, want_stages: bool = False,
                    out: torch.Tensor = None):
        """Pipelined host-buffer call (``sepref_separator_submit_host``): queue H2D copy -> kernels -> D2H copy for
        one batch in staging slot ``slot`` (0/1) and return at once; ``wait_host(slot)`` returns the CPU tensors.
        Keeping two slots in flight overlaps the copies of neighbouring batches with the kernels - the loop shape of
        the reference's test pass (engine.py:165-167).  ``x_host`` and ``out`` should be pinned."""
        if x_host.is_cuda:
            raise RuntimeError("submit_host takes a CPU tensor")
        s = self.shape_
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        x = x_host.to(torch.float32).contiguous()
        B, F, L = x.shape
        Tp = self.padded_frames(L)
        Td = Tp >> s.num_stages
        h = self._handle_for(device)
        lib = _lib.lib()
        if out is None:
            out = torch.empty(B * s.num_spks, F, Tp, dtype=torch.float32, pin_memory=True)
        elif tuple(out.shape) != (B * s.num_spks, F, Tp) or out.dtype != torch.float32 or out.is_cuda or not out.is_contiguous():
            raise RuntimeError("out must be a contiguous fp32 CPU tensor of shape [B*S, F, L_pad]")
        stages, ptrs = [], (C.c_void_p * s.num_stages)()
        for i in range(s.num_stages):
            if want_stages:
                t = torch.empty(B * s.num_spks, F, Td << i, dtype=torch.float32, pin_memory=True)
                stages.append(t)
                ptrs[i] = t.data_ptr()
            else:
                ptrs[i] = None
        with torch.cuda.device(device):
            _lib.check(lib.sepref_separator_submit_host(h.ptr, int(slot), x.data_ptr(), B, L, out.data_ptr(), ptrs),
                       "sepref_separator_submit_host")
        self.last_launch_count = lib.sepref_last_launch_count(h.ptr)
        h.pending[int(slot)] = (x, out, stages)        # keeps the host buffers alive until wait_host
        return int(slot)

    def wait_host(self, slot: int = 0, device=None):
        """Block until the request submitted in ``slot`` has landed in host memory; returns ``(last, stages)``."""
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        h = self._handle_for(device)
        if int(slot) not in h.pending:
            raise RuntimeError(f"nothing submitted in slot {slot}")
        _lib.check(_lib.lib().sepref_separator_wait_host(h.ptr, int(slot)), "sepref_separator_wait_host")
        _, out, stages = h.pending.pop(int(slot))
        return out, stages

    def profile_kernels(self, x: torch.Tensor, steps: int = 3):
        """Per-kernel device time of one forward (ms, averaged over ``steps``), from CUDA events recorded on the
        launching stream by the library (``SEPREF_OPT_PROFILE``).  Keys: ``<kernel>_ms``, ``<kernel>_launches``."""
        h = self._handle_for(x.device)
        lib = _lib.lib()
        _lib.check(lib.sepref_set_option(h.ptr, _lib.OPT_PROFILE, 1))
        acc: Dict[str, float] = {}
        try:
            for _ in range(steps):
                self.forward(x)
                buf = C.create_string_buffer(1 << 16)
                _lib.check(lib.sepref_profile_report(h.ptr, buf, len(buf)), "sepref_profile_report")
                for line in buf.value.decode().splitlines():
                    name, ms, n = line.split()
                    short = name.split("::")[-1].replace("k_", "")
                    acc[short + "_ms"] = acc.get(short + "_ms", 0.0) + float(ms) / steps
                    acc[short + "_launches"] = int(n)
        finally:
            _lib.check(lib.sepref_set_option(h.ptr, _lib.OPT_PROFILE, 0))
        return acc

    # ------------------------------------------------------------------ block-level calls (unit parity)
    def run_block(self, kind: str, prefix: str, x: torch.Tensor, *, td: int = 0, x_low: torch.Tensor = None):
        """Run one block through its C-ABI entry point on channels-last ``x [rows, T, F]``."""
        assert x.is_cuda and x.dtype == torch.float32
        x = x.contiguous()
        rows, t, f = x.shape
        h = self._handle_for(x.device)
        lib = _lib.lib()
        s = self.shape_
        out_shape = {"down_conv": (rows, t // 2, f), "spk_split": (rows * s.num_spks, t, f)}.get(kind, (rows, t, f))
        with torch.cuda.device(x.device):
            y = torch.empty(out_shape, device=x.device, dtype=torch.float32)
            nbytes = lib.sepref_block_workspace_bytes(h.ptr, rows, t)
            ws = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
            st = torch.cuda.current_stream(x.device).cuda_stream
            p = prefix.encode()
            if kind in ("gcfn", "cla", "local_block", "spk_attention", "spk_split"):
                fn = getattr(lib, f"sepref_{kind}_forward")
                rc = fn(h.ptr, p, x.data_ptr(), rows, t, y.data_ptr(), ws.data_ptr(), ws.numel(), st)
            elif kind in ("ega", "global_block"):
                fn = getattr(lib, f"sepref_{kind}_forward")
                rc = fn(h.ptr, p, x.data_ptr(), rows, t, td, y.data_ptr(), ws.data_ptr(), ws.numel(), st)
            elif kind == "down_conv":
                rc = lib.sepref_down_conv_forward(h.ptr, p, x.data_ptr(), rows, t, y.data_ptr(), st)
            elif kind == "fusion":
                x_low = x_low.contiguous()
                rc = lib.sepref_fusion_forward(h.ptr, p, x_low.data_ptr(), x.data_ptr(), rows, t, y.data_ptr(),
                                               ws.data_ptr(), ws.numel(), st)
            else:
                raise ValueError(kind)
            _lib.check(rc, f"sepref_{kind}_forward")
        return y
