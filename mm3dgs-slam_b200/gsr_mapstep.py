"""Keyframe-sharded mapping step: one process per GPU, replicated Gaussian
parameters, each rank renders + back-propagates its share of the K keyframes of the covisibility
window, and ONE all-reduce(SUM) over a flat fp32 gradient bucket combines them.

The reference optimises one random keyframe per Adam step and has no distributed code
(R/slam/mapper.py:797-828,938-939); the single-GPU ground truth of this construct is plain
gradient accumulation over the same K keyframes followed by one optimizer step, which is exactly
what `ShardedMapStep` computes at world_size == 1.

Gradients land directly in the bucket: every parameter's `.grad` is a view into one contiguous
tensor (14 floats = 56 B per Gaussian at SH degree 0, in the order of the reference's optimizer
groups, R/slam/gaussian_model.py:151-187), autograd accumulates in place, and the collective runs
on that tensor without a gather/flatten copy.  Rank r takes keyframes r, r+G, r+2G, ...; ranks with
no keyframe still join the all-reduce with zeros.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


class PeerExchange:
    """The gradient exchange over peer memory (include/gscomm_b200.h): the bucket lives in a symmetric allocation that
    every rank of the node maps (torch.distributed._symmetric_memory does the handle exchange — plumbing), and ONE
    library kernel per step pulls, reduces and broadcasts it over NVLink (multimem instructions through the NVSwitch
    when the node has NVLS).  No NCCL call on the path.  Raises if symmetric memory cannot be set up; the caller then
    stays on NCCL's all_reduce."""

    def __init__(self, n: int, device, group=None, mode: str = "auto"):
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem

        from diff_gaussian_rasterization import _lib
        self._lib, self._ct = _lib, ctypes
        grp = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(grp), dist.get_rank(grp)
        if self.world not in (2, 4, 8):
            raise RuntimeError("peer exchange needs 2, 4 or 8 ranks")
        _lib.gsr_allreduce_flag_words.restype = ctypes.c_size_t
        _lib.gsr_allreduce_sum_f32.restype = ctypes.c_int
        _lib.gsr_allreduce_sum_f32.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p),
                                               ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int64, ctypes.c_uint32]
        _lib.gsr_allgather_f32.restype = ctypes.c_int
        _lib.gsr_allgather_f32.argtypes = _lib.gsr_allreduce_sum_f32.argtypes
        self.n = n
        n_pad = (n + 3) // 4 * 4
        self._store = symm_mem.empty(n_pad, dtype=torch.float32, device=device)
        self._store.zero_()
        self._h = symm_mem.rendezvous(self._store, grp)
        words = int(_lib.gsr_allreduce_flag_words())
        self._flags = symm_mem.empty(words, dtype=torch.int32, device=device)
        self._flags.zero_()
        self._hf = symm_mem.rendezvous(self._flags, grp)
        torch.cuda.synchronize(device)
        dist.barrier(group=grp)                       # every rank's flags are zero before anybody's first handshake
        self.flat = self._store[:n]
        # NVLS halves the bytes every link carries from 4 ranks up; with 2 ranks the switch has nothing to reduce and the
        # plain peer loads / stores measured faster (106 vs 164 us for the 56 MB bucket)
        want_mc = mode == "nvls" or (mode == "auto" and self.world >= 4)
        mc = int(self._h.multicast_ptr) if want_mc else 0
        if mode == "nvls" and not mc:
            raise RuntimeError("no multicast (NVLS) mapping on this node")
        self.mode = "nvls" if mc else "peer"
        self._mc = mc or None
        self._bp = (ctypes.c_void_p * self.world)(*[int(p) for p in self._h.buffer_ptrs])
        self._fp = (ctypes.c_void_p * self.world)(*[int(p) for p in self._hf.buffer_ptrs])
        self.epoch = 0

    def _call(self, fn):
        self.epoch += 1
        rc = fn(torch.cuda.current_stream(self.flat.device).cuda_stream, self.world, self.rank, self._bp, self._mc, self._fp,
                self.n, self.epoch)
        if rc != 0:
            raise RuntimeError("gsrast_b200: " + self._lib.gsr_last_error().decode())

    def all_reduce(self):
        """SUM over the ranks, in place, on torch's current stream; nothing waits on the host."""
        self._call(self._lib.gsr_allreduce_sum_f32)

    def all_gather(self):
        """Every rank's slice (see slice_range) is copied into every other rank's buffer; current stream, no host wait."""
        self._call(self._lib.gsr_allgather_f32)

    def slice_range(self, rank=None):
        """[lo, hi) in floats of the slice rank `rank` owns (reduces / contributes)."""
        r = self.rank if rank is None else rank
        n4 = (self.n + 3) // 4
        per = (n4 + self.world - 1) // self.world
        return min(per * r * 4, self.n), min(per * (r + 1) * 4, self.n)


class GradBucket:
    """One flat fp32 tensor holding the gradients of all parameters, exposed as per-parameter views.
    `flat`: storage to use (e.g. PeerExchange.flat) instead of a fresh allocation."""

    def __init__(self, params: Dict[str, torch.Tensor], flat: Optional[torch.Tensor] = None):
        self.names = list(params.keys())
        self.params = params
        n = sum(p.numel() for p in params.values())
        first = next(iter(params.values()))
        if flat is not None and (flat.numel() != n or flat.dtype != torch.float32 or flat.device != first.device):
            raise ValueError("bucket storage must be a flat fp32 tensor of the parameters' total size on their device")
        self.flat = flat if flat is not None else torch.zeros(n, dtype=torch.float32, device=first.device)
        self.views: Dict[str, torch.Tensor] = {}
        off = 0
        for k, p in params.items():
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise ValueError(f"parameter {k} must be contiguous fp32")
            self.views[k] = self.flat[off: off + p.numel()].view_as(p)
            off += p.numel()
        self.attach()

    def attach(self):
        """(Re)bind every parameter's .grad to its bucket view so autograd accumulates in place."""
        for k, p in self.params.items():
            p.grad = self.views[k]

    def zero_(self):
        self.flat.zero_()

    @property
    def nbytes(self):
        return self.flat.numel() * 4


class DensificationStats:
    """The per-Gaussian statistics the reference mapper accumulates next to the gradients, made
    consistent across ranks:

      max_radii2D[vis]        = max(max_radii2D[vis], radii[vis])          R/slam/mapper.py:892-895
      xyz_gradient_accum[vis] += || viewspace_points.grad[vis, :2] ||      R/slam/gaussian_model.py:594-598
      denom[vis]              += 1

    The accumulated quantity is the SUM over keyframes of per-keyframe norms (not the norm of the summed
    gradient), so every rank accumulates its own keyframes locally (`add`) and `reduce` then needs one
    all-reduce(SUM) over the packed [2,P] (accum, denom) and one all-reduce(MAX) over [P] (radii)."""

    def __init__(self, num_gaussians: int, device, group: Optional[dist.ProcessGroup] = None):
        self.sum = torch.zeros(2, num_gaussians, dtype=torch.float32, device=device)   # [accum, denom]
        self.max_radii = torch.zeros(num_gaussians, dtype=torch.float32, device=device)
        self.group = group

    def add(self, radii: torch.Tensor, viewspace_grad: torch.Tensor):
        vis = radii > 0
        self.sum[0] += torch.where(vis, viewspace_grad[:, :2].norm(dim=-1), torch.zeros((), device=radii.device))
        self.sum[1] += vis.to(torch.float32)
        self.max_radii = torch.maximum(self.max_radii, torch.where(vis, radii.to(torch.float32), self.max_radii))

    def reduce(self):
        """Returns (xyz_gradient_accum [P,1], denom [P,1], max_radii2D [P]) identical on every rank."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.sum, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(self.max_radii, op=dist.ReduceOp.MAX, group=self.group)
        return self.sum[0].unsqueeze(1), self.sum[1].unsqueeze(1), self.max_radii


def shard_keyframes(num_keyframes: int, rank: int, world: int) -> List[int]:
    """Indices of the keyframes rank `rank` owns: r, r+G, r+2G, ..."""
    return list(range(rank, num_keyframes, world))


class ShardedMapStep:
    """Two ways to describe the per-keyframe work:

    * frame_fn(params, keyframe): runs forward + backward for one keyframe, accumulating into the
      parameters' .grad (it may return a detached scalar loss);
    * forward_fn(params, keyframe) -> (output, grad_output or None): runs only the forward (and the loss,
      if any).  `step` then issues ALL of this rank's forwards first and back-propagates them together
      with one torch.autograd.backward call.  The forward has one host hand-off per frame (the instance
      count R); issuing the forwards back to back lets the GPU work of frame k hide the host latency of
      frame k+1, and the backwards (no host sync at all) then stream without bubbles.  Costs one set of
      rasterizer workspaces per in-flight keyframe (~160 MB at 1M Gaussians / 640x480).
    """

    def __init__(self, params: Dict[str, torch.Tensor], frame_fn: Optional[Callable] = None,
                 group: Optional[dist.ProcessGroup] = None, forward_fn: Optional[Callable] = None,
                 streams: int = 1, direct_targets: bool = False, prepare_fn: Optional[Callable] = None,
                 overwrite_first: bool = False, exchange: str = "auto"):
        """streams > 1 (forward_fn mode, CUDA only): consecutive keyframes alternate between `streams` CUDA
        streams, so the latency-bound binning kernels of one frame overlap the blend kernels of another.
        direct_targets: forward_fn receives a third argument, a dict of gradient accumulators (views of THE flat
        bucket) for kernels that add their gradients in place.  Every stream adds into the same bucket: the library's
        backward kernels accumulate with TMA reduce-adds / float atomics, so concurrent frames do not race.
        overwrite_first (needs direct_targets, ONE stream, and a forward_fn that hands `targets` to the rasterizer's
        grad_targets unchanged): the first frame of a step carries targets["_overwrite"] = True, so the bucket needs no
        zero fill (only its opacity slice, which the blend backward adds into) and that frame's gradients are stored
        instead of added.  (A rank that owns a single keyframe always works this way.)"""
        if (frame_fn is None) == (forward_fn is None):
            raise ValueError("give exactly one of frame_fn / forward_fn")
        self.params = params
        # exchange: "nccl" = dist.all_reduce; "peer" / "nvls" = the library's one-kernel exchange over peer memory
        # (P2P loads/stores / multimem through the switch); "auto" = nvls, else peer, else nccl (CPU / gloo: nccl path).
        self.peer = None
        self.exchange_error = None
        first_ = next(iter(params.values()))
        if (exchange != "nccl" and first_.is_cuda and dist.is_available() and dist.is_initialized()
                and dist.get_world_size(group) > 1):
            try:
                self.peer = PeerExchange(sum(p.numel() for p in params.values()), first_.device, group, exchange)
            except Exception as ex:   # noqa: BLE001 - any failure of the symmetric-memory plumbing means NCCL
                if exchange != "auto":
                    raise
                self.exchange_error = repr(ex)[:300]
        self.exchange = self.peer.mode if self.peer is not None else "nccl"
        self.bucket = GradBucket(params, self.peer.flat if self.peer is not None else None)
        self.frame_fn = frame_fn
        self.forward_fn = forward_fn
        # prepare_fn(params, keyframe) -> handle: enqueues the part of the forward that precedes its host
        # hand-off (diff_gaussian_rasterization.prepare_forward); it is issued for ALL of the rank's keyframes
        # before the first forward_fn call, which then receives the handle as its last argument.
        self.prepare_fn = prepare_fn
        self.group = group
        self.direct_targets = direct_targets
        first = next(iter(params.values()))
        self.nstreams = max(1, int(streams)) if (forward_fn is not None and first.is_cuda) else 1
        # (with several streams an overwriting frame would race with the frames adding on the other streams)
        self.overwrite_first = bool(overwrite_first and direct_targets and self.nstreams == 1)
        self.single_frame_overwrite = True    # see _step
        self._single = False
        self.streams = [torch.cuda.Stream(first.device) for _ in range(self.nstreams)] if self.nstreams > 1 else []
        self.distributed = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.world = dist.get_world_size(group) if self.distributed else 1
        # optional per-step timing (CUDA events on the step's main stream): set .timing = True, run steps back to back
        # (nothing waits on the host), then read .timings()
        self.timing = False
        self._marks = None
        self._history = []

    def _mark(self, i):
        if self.timing and self._marks is not None:
            self._marks[i].record(torch.cuda.current_stream())

    def timings(self):
        """After steps run with .timing = True: list of dict(kernel_ms, comm_ms, step_ms) per step on this rank —
        kernel_ms from the start of the step to the point where the local gradient bucket is complete, comm_ms the
        exchange after it, step_ms from this step's start to the next step's start.  Clears the history."""
        torch.cuda.synchronize()
        h, self._history = self._history, []
        out = []
        for i, m in enumerate(h):
            d = {"kernel_ms": m[0].elapsed_time(m[1]), "comm_ms": m[1].elapsed_time(m[2])}
            if i + 1 < len(h):
                d["step_ms"] = m[0].elapsed_time(h[i + 1][0])
            out.append(d)
        return out

    def _targets(self, i, first=False):
        """Accumulators of stream i.  `first`: this is the first frame written into them in this step, so the
        kernels may overwrite (only the opacity slice, which the blend backward adds into, was zeroed)."""
        if not self.direct_targets:
            return None
        views = self.bucket.views
        return dict(views, _overwrite=True) if (first and (self.overwrite_first or self._single)) else views

    def _call_forward(self, kf, i, handle, first=False):
        args = (self.params, kf) + ((self._targets(i, first),) if self.direct_targets else ())
        if self.prepare_fn is not None:
            args = args + (handle,)
        return self.forward_fn(*args)

    def _forwards(self, mine):
        n = self.nstreams
        if n == 1:
            handles = [self.prepare_fn(self.params, kf) for kf in mine] if self.prepare_fn else [None] * len(mine)
            return [self._call_forward(kf, 0, h, first=(j == 0)) for j, (kf, h) in enumerate(zip(mine, handles))]
        main = torch.cuda.current_stream()
        for st in self.streams:
            st.wait_stream(main)
        handles = [None] * len(mine)
        if self.prepare_fn is not None:
            for j, kf in enumerate(mine):
                with torch.cuda.stream(self.streams[j % n]):
                    handles[j] = self.prepare_fn(self.params, kf)
        outs = []
        for j, kf in enumerate(mine):
            with torch.cuda.stream(self.streams[j % n]):
                outs.append(self._call_forward(kf, j % n, handles[j], first=(j == 0)))
        return outs

    def my_keyframes(self, keyframes: Sequence):
        return [keyframes[i] for i in shard_keyframes(len(keyframes), self.rank, self.world)]

    def step(self, keyframes: Sequence, params: Optional[Dict[str, torch.Tensor]] = None):
        """Sum of per-keyframe gradients in self.bucket.flat on every rank. Returns the local losses.
        `params`: optional replacement parameter tensors for this step (same shapes) — used when the caller
        double-buffers parameter uploads; only valid with direct_targets (gradients do not go through .grad)."""
        if params is not None and not self.direct_targets:
            raise ValueError("per-step params need direct_targets=True")
        saved = self.params
        if params is not None:
            self.params = params
        try:
            return self._step(keyframes)
        finally:
            self.params = saved

    def _step(self, keyframes: Sequence):
        mine = self.my_keyframes(keyframes)
        if self.timing and next(iter(self.params.values())).is_cuda:
            self._marks = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            self._history.append(self._marks)
        else:
            self._marks = None
        self._mark(0)
        # this rank owns ONE keyframe (K keyframes on K GPUs): the bucket has exactly one writer, so its kernels may
        # overwrite instead of add — no 56 B/Gaussian zero fill, no read-modify-write
        self._single = bool(self.direct_targets and self.forward_fn is not None and len(mine) == 1
                            and self.single_frame_overwrite)
        if (self.overwrite_first or self._single) and self.forward_fn is not None and mine:
            self.bucket.views["opacities"].zero_()
        else:
            self.bucket.zero_()
        if not self.direct_targets:
            self.bucket.attach()
        if self.forward_fn is not None:
            outs = self._forwards(mine)
            if outs and self.overwrite_first and not self._single:
                # the frame that overwrites a bucket must be back-propagated before the frames that add to it:
                # one backward call per frame, in forward order (a joint call leaves the order to the engine)
                for o, g in outs:
                    torch.autograd.backward(o, g)
            elif outs:
                torch.autograd.backward([o for o, _ in outs], [g for _, g in outs])
            if self.nstreams > 1:
                main = torch.cuda.current_stream()
                for st in self.streams:
                    main.wait_stream(st)
                for o, _ in outs:
                    o.record_stream(main)
            losses = [o.detach() for o, _ in outs]
        else:
            losses = [self.frame_fn(self.params, kf) for kf in mine]
        self._mark(1)
        if self.world > 1:
            if self.peer is not None:
                self.peer.all_reduce()
            else:
                dist.all_reduce(self.bucket.flat, op=dist.ReduceOp.SUM, group=self.group)
        self._mark(2)
        return losses
