"""Row-strip partitioning of a frame across GPUs.

One process per GPU.  Every rank holds full-frame per-camera buffers, computes rows [y0, y1) of each
pass and, before every *gathering* pass, receives the halo rows that pass reads from the ranks that own
them (NCCL send/recv over NVLink through torch.distributed).  RNG streams and buffer indices are keyed on
absolute pixel coordinates, so a strip-partitioned run reproduces the single-GPU frame bit for bit.

`plan_frame` (pure Python, no GPU) turns a frame's pass schedule into exchange points; `StripRunner`
executes it.  The transport is pluggable so the plan can be exercised with gloo on CPU tensors and with
several engines inside one process on a single GPU (tests).
"""
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import numpy as np

# pass ids (strolle_b200/csrc/st_types.h PassId)
P_PRIM_GBUFFER, P_DI_SAMPLING, P_DI_TEMPORAL, P_DI_SPATIAL_PICK, P_DI_SPATIAL_TRACE, P_DI_SPATIAL_SAMPLE = 0, 1, 2, 3, 4, 5
P_DI_RESOLVING, P_GI_REPROJECTION, P_GI_SAMPLING_A, P_GI_SAMPLING_B, P_GI_TEMPORAL, P_GI_SPATIAL_PICK = 6, 7, 8, 9, 10, 11
P_GI_SPATIAL_TRACE, P_GI_SPATIAL_SAMPLE, P_GI_PREVIEW, P_GI_RESOLVING, P_FRAME_REPROJECTION = 12, 13, 14, 15, 16
P_DENOISE_REPROJECT, P_DENOISE_VARIANCE, P_DENOISE_WAVELET, P_COMPOSITION = 17, 18, 19, 20

SPATIAL_REACH = 128      # ReSTIR spatial taps: radius <= 128 px (di_spatial_resampling.rs:55-56)
PREVIEW2_REACH = 64      # second preview pass (gi_preview_resampling.rs:64-70)
VARIANCE_REACH = 3       # estimate_variance window (frame_denoising.rs:128-190)
WAVELET_REACH = [1, 2, 4, 9, 19]   # stride s plus jitter trunc((s-1)/4) (frame_denoising.rs:269-286)

# float4s per pixel of every exchangeable buffer
VEC4_PER_PIXEL = {"di_reservoirs_0": 2, "di_reservoirs_1": 2, "di_reservoirs_2": 2,
                  "gi_reservoirs_0": 4, "gi_reservoirs_1": 4, "gi_reservoirs_2": 4, "gi_reservoirs_3": 4}


STRIP_SIDE_ROWS = 36   # what one neighbour costs a strip, in rows of its own (engine.cu::kStripSideRows)


def strip_bounds(height: int, world: int) -> List[Tuple[int, int]]:
    """Rows [y0, y1) of every rank, contiguous.  From three ranks on the two outer strips (one neighbour each) get STRIP_SIDE_ROWS rows
    more than the inner ones (two neighbours: twice the recomputed / mirrored halo rows), unless the inner strips would get shorter
    than 160 rows; equal strips otherwise.  Same arithmetic as engine.cu::strip_bounds (st_strip_bounds)."""
    k = STRIP_SIDE_ROWS
    if world < 3 or (height + k * (2 * world - 2)) // world - 2 * k < 160:
        k = 0
    total = height + k * (2 * world - 2)
    edges = [0] + [total * r // world - k * (2 * r - 1) for r in range(1, world)] + [height]
    return [(edges[r], edges[r + 1]) for r in range(world)]


@dataclass
class Exchange:
    before_step: int                       # index into the frame schedule; len(schedule) = after the last pass
    buffers: List[Tuple[str, int]]         # (buffer name, reach in rows)


def plan_frame(schedule: Sequence[int], frame: int, temporal_reach: int = 16) -> List[Exchange]:
    """Exchange points of one frame.  `schedule` = pass ids in launch order (st_frame_schedule)."""
    cur = "b" if frame % 2 == 1 else "a"
    prv = "a" if cur == "b" else "b"
    plan: List[Exchange] = []
    have_gbuffer = False
    nth_preview = 0
    nth_wavelet = 0
    # ping-pong of the wavelet passes (strolle/src/camera_controller/passes/frame_denoising.rs:78-108)
    wavelet_inputs = ["stash", "prev_colors", "stash", "curr_colors", "stash"]
    gi_source = None
    if P_GI_PREVIEW in schedule:
        gi_source = 2 if P_GI_SPATIAL_PICK in schedule else 1
    for i, p in enumerate(schedule):
        bufs: List[Tuple[str, int]] = []
        if i == 0 and temporal_reach > 0:
            # last frame's outputs that this frame gathers at reprojected positions (K4, K6, K11, K14, K20)
            bufs += [(f"prim_surface_map_{prv}", temporal_reach), (f"prim_gbuffer_d0_{prv}", temporal_reach), (f"prim_gbuffer_d1_{prv}", temporal_reach),
                     ("di_reservoirs_0", temporal_reach), ("gi_reservoirs_0", temporal_reach), ("di_diff_prev_colors", temporal_reach),
                     ("gi_diff_prev_colors", temporal_reach), (f"di_diff_moments_{prv}", temporal_reach), (f"gi_diff_moments_{prv}", temporal_reach)]
        if p in (P_DI_SPATIAL_PICK, P_GI_SPATIAL_PICK):
            if not have_gbuffer:
                bufs += [(f"prim_gbuffer_d0_{cur}", SPATIAL_REACH), (f"prim_gbuffer_d1_{cur}", SPATIAL_REACH), ("surface_nd", SPATIAL_REACH)]
                have_gbuffer = True
            bufs.append(("di_reservoirs_1" if p == P_DI_SPATIAL_PICK else "gi_reservoirs_1", SPATIAL_REACH))
        elif p == P_GI_PREVIEW:
            if nth_preview == 0:
                bufs += [(f"prim_surface_map_{cur}", SPATIAL_REACH), (f"gi_reservoirs_{gi_source}", SPATIAL_REACH)]
                if not have_gbuffer:
                    bufs.append(("surface_nd", SPATIAL_REACH))
                    have_gbuffer = True
            else:
                bufs.append(("gi_reservoirs_3", PREVIEW2_REACH))
            nth_preview += 1
        elif p == P_DENOISE_VARIANCE:
            bufs += [("di_diff_curr_colors", VARIANCE_REACH), ("gi_diff_curr_colors", VARIANCE_REACH)]
            if not have_gbuffer:
                bufs.append(("surface_nd", WAVELET_REACH[-1]))
        elif p == P_DENOISE_WAVELET:
            src = wavelet_inputs[nth_wavelet]
            bufs += [(f"di_diff_{src}", WAVELET_REACH[nth_wavelet]), (f"gi_diff_{src}", WAVELET_REACH[nth_wavelet])]
            nth_wavelet += 1
        if bufs:
            plan.append(Exchange(i, bufs))
    return plan


def halo_transfers(bounds: Sequence[Tuple[int, int]], height: int, reach: int) -> List[Tuple[int, int, int, int]]:
    """(src_rank, dst_rank, row0, row1) for every block of rows rank `dst` needs (its strip grown by `reach`)
    and rank `src` owns.  Deterministic order shared by all ranks."""
    out = []
    for dst, (d0, d1) in enumerate(bounds):
        need0, need1 = max(0, d0 - reach), min(height, d1 + reach)
        for src, (s0, s1) in enumerate(bounds):
            if src == dst:
                continue
            a, b = max(need0, s0), min(need1, s1)
            if a < b:
                out.append((src, dst, a, b))
    return out


class _DevArray:
    """Exposes an engine-owned device allocation through __cuda_array_interface__ (zero-copy torch view)."""

    def __init__(self, ptr, nfloats):
        self.__cuda_array_interface__ = {"shape": (nfloats,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}


class TorchDistTransport:
    """NCCL (or gloo) point-to-point over torch.distributed; ordered against torch's current stream."""

    def __init__(self, rank):
        import torch.distributed as dist
        self.dist = dist
        self.rank = rank

    def run(self, ops):
        """ops: list of (src, dst, tensor_view) where tensor_view is this rank's view of the rows."""
        dist = self.dist
        reqs = []
        for src, dst, view in ops:
            if src == self.rank:
                reqs.append(dist.P2POp(dist.isend, view, dst))
            elif dst == self.rank:
                reqs.append(dist.P2POp(dist.irecv, view, src))
        if reqs:
            for w in dist.batch_isend_irecv(reqs):
                w.wait()


class StripRunner:
    """Renders one camera on one rank of a strip-partitioned (or single-GPU) run."""

    def __init__(self, engine, cam, width, height, rank=0, world=1, transport=None, temporal_reach=16, native=None, peer=True):
        """`native` (default when no transport is injected): the whole frame, halo exchanges included, is enqueued
        by one st_render_strips call over the engine's own NCCL communicator; otherwise the exchanges go through
        `transport` (torch.distributed P2P, or an in-process emulation in tests) between st_render_range calls."""
        self.engine, self.cam, self.w, self.h, self.rank, self.world = engine, cam, width, height, rank, world
        self.bounds = strip_bounds(height, world)
        self.y0, self.y1 = self.bounds[rank]
        self.temporal_reach = temporal_reach
        self.transport = transport
        self._views: Dict[str, object] = {}
        self.halo_bytes_last_frame = 0
        self.native = False
        self.peer = False
        if world > 1:
            import torch
            engine.set_strip(cam, self.y0, self.y1)
            self.native = (transport is None) if native is None else native
            if self.native:
                import torch.distributed as dist
                self.peer = bool(peer)
                if self.peer:   # map every rank's camera buffers (CUDA IPC): rows then travel as peer stores / loads on the engine's own stream
                    handles = [None] * world
                    dist.all_gather_object(handles, engine.peer_export(cam))
                    engine.peer_import(cam, handles, rank, world)
                    dist.barrier()
                else:           # engine-owned NCCL communicator, ordered against torch's stream
                    from .engine import nccl_unique_id
                    engine.set_stream(torch.cuda.current_stream().cuda_stream)
                    box = [nccl_unique_id() if rank == 0 else None]
                    dist.broadcast_object_list(box, src=0)
                    engine.nccl_init(box[0], rank, world)
            else:
                engine.set_stream(torch.cuda.current_stream().cuda_stream)
                if transport is None:
                    self.transport = TorchDistTransport(rank)

    def _view(self, name):
        """(H, floats_per_row) torch view of an engine buffer."""
        import torch
        if name not in self._views:
            ptr, nbytes = self.engine.buffer_device_ptr(self.cam, name)
            t = torch.as_tensor(_DevArray(ptr, nbytes // 4), device="cuda")
            self._views[name] = t.view(self.h, -1)
        return self._views[name]

    def _exchange(self, ex: Exchange):
        ops = []
        nbytes = 0
        for name, reach in ex.buffers:
            v = self._view(name)
            for src, dst, a, b in halo_transfers(self.bounds, self.h, reach):
                if src == self.rank or dst == self.rank:
                    ops.append((src, dst, v[a:b]))
                    if dst == self.rank:
                        nbytes += v[a:b].numel() * 4
        self.halo_bytes_last_frame += nbytes
        self.transport.run(ops)

    def transport_name(self):
        if self.world == 1:
            return "single GPU"
        if self.native and self.peer:
            return ("fused peer-memory transport over NVLink (CUDA IPC): DI / preview / SVGF boundary rows stored into the neighbours' buffers by "
                    "the kernels that produce them, the 128-row GI reservoir halos pushed by the copy engines on side streams, neighbour-only "
                    "sequence flags, G-buffer / SVGF halo rows recomputed, temporal rows pulled on demand; no NCCL on the data path")
        return "engine-owned NCCL send/recv per exchange point" if self.native else "torch.distributed P2P between st_render_range calls"

    def render(self, out=None, fmt=0, gather=None, moving=False):
        """`out` (world > 1): a host frame shared by all ranks (gather 2, default: every rank copies its own rows), or rank 0's
        private buffer with gather=1 (strips assembled on rank 0 first)."""
        eng, cam = self.engine, self.cam
        if self.world == 1:
            eng.render_camera(cam, out, fmt)
            return
        if self.native:
            g = 0 if out is None else (2 if gather is None else gather)
            eng.render_strips(cam, out, fmt, self.temporal_reach, gather=g)
            self.halo_bytes_last_frame = eng.halo_bytes()
            return
        schedule = eng.frame_schedule(cam)
        frame = eng.frame() - 1   # tick() already advanced the engine's counter; the camera renders frame-1
        # a fixed temporal halo is only enough while nothing moves: under motion the whole of last frame's buffers is exchanged
        plan = plan_frame(schedule, frame, self.h if moving else self.temporal_reach)
        self.halo_bytes_last_frame = 0
        first = 0
        for ex in plan:
            if ex.before_step > first:
                eng.render_range(cam, first, ex.before_step - 1)
                first = ex.before_step
            self._exchange(ex)
        eng.render_range(cam, first, len(schedule) - 1)
        if out is not None:
            self.gather_output(out, fmt)

    def gather_output(self, out, fmt):
        """Assembles the full composed frame on every rank (all strips), then copies it to `out` on rank 0."""
        v = self._view("output")
        ops = []
        for src, (s0, s1) in enumerate(self.bounds):
            if src != 0 and (self.rank == 0 or self.rank == src):
                ops.append((src, 0, v[s0:s1]))
        self.transport.run(ops)
        if self.rank == 0:
            self.engine.set_strip(self.cam, 0, self.h)
            n = len(self.engine.frame_schedule(self.cam))
            # re-run only the output conversion + copy on the assembled frame
            self.engine.copy_output(self.cam, out, fmt)
            self.engine.set_strip(self.cam, self.y0, self.y1)


class ReferenceAccumulator:
    """Sample-parallel path-traced reference mode (BASELINE config C5).

    `CameraMode::Reference` accumulates one path-traced sample per frame into `ref_colors`
    (strolle-shaders/src/ref_shading.rs:53-66).  Seeds depend on the frame id only, so rank g renders
    accumulations g+1, g+1+N, ... on its own full frame; the partial sums are then added with one NCCL
    reduce (f32 sum of 16 B/pixel) and rank 0 composes `rgb / w`.  The result equals the single-GPU sum up to
    f32 addition order.
    """

    def __init__(self, engine, cam, rank=0, world=1):
        self.engine, self.cam, self.rank, self.world = engine, cam, rank, world
        if world > 1:
            import torch
            engine.set_stream(torch.cuda.current_stream().cuda_stream)

    def accumulate(self, total):
        eng, cam = self.engine, self.cam
        for k in range(self.rank, total, self.world):
            eng.set_frame(k + 1)
            eng.tick()
            n = len(eng.frame_schedule(cam))
            eng.render_range(cam, 0, n - 2)      # every pass but the final composition

    def reduce_and_compose(self):
        """Sums ref_colors onto rank 0 and composes there; returns nothing (read "output" on rank 0)."""
        eng, cam = self.engine, self.cam
        if self.world > 1:
            import torch
            import torch.distributed as dist
            ptr, nbytes = eng.buffer_device_ptr(cam, "ref_colors")
            t = torch.as_tensor(_DevArray(ptr, nbytes // 4), device="cuda")
            dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
        if self.rank == 0:
            n = len(eng.frame_schedule(cam))
            eng.render_range(cam, n - 1, n - 1)
