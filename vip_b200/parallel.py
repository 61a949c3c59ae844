"""Single-cube PCA sharded over the GPUs of one node (one process per GPU, ``torch.distributed``).

SURVEY.md 8(e): the path shards with two real exchange steps and no other communication.

    pixel shards  M[:, p_g]   --  G_g = M_g M_g^T  --all-reduce(sum, fp64 n x n)-->  G
                                  top-k eigenpairs of G (replicated, broadcast from rank 0)
                                  V_g = Wt M_g ,  R_g = M_g - C V_g                 (local)
    all-to-all #1 (pixel -> frame shards):   R[f_g, :]
                                  derotate own frames                               (local)
    all-to-all #2 (frame -> pixel shards):   D[:, p_g]
                                  collapse own pixels, gather the frame on rank 0
    (collapse 'mean' / 'sum' are reducible: instead of all-to-all #2, partial sums over the own frames and
     one NCCL reduce of the (H,W) frame to rank 0)

The reference has no distributed mode (its ``nproc`` forks processes over frames,
``preproc/derotation.py:392-397``); this module is new functionality behind the same ``pca`` semantics
(3-d ADI cube, integer ``ncomp``, deterministic ``svd_mode``).

The arithmetic goes through an ``ops`` object: :class:`CudaOps` (the CUDA kernels of this package) by
default.  The CPU tests of the communication logic (``tests/test_parallel_gloo.py``, gloo backend,
world_size 2) pass a numpy test double instead -- the product never computes on the CPU.
"""
import os

import numpy as np
import torch
import torch.distributed as dist

from . import kernels
from .preproc.derotation import derotate_device
from .preproc.parangles import check_pa_vector
from .preproc.subsampling import collapse_device


def shard_bounds(total, world):
    """Contiguous, near-equal shards: returns world+1 offsets."""
    base, rem = divmod(total, world)
    sizes = [base + (1 if r < rem else 0) for r in range(world)]
    return np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)


class CudaOps:
    """The CUDA kernels behind the distributed driver (all tensors are CUDA tensors)."""

    name = "cuda"

    def upload_pixels(self, host2d, c0, c1, device):
        if host2d.dtype == np.float32 and host2d.flags["C_CONTIGUOUS"]:
            return kernels.upload_columns(host2d, c0, c1, device)          # one strided DMA
        return torch.from_numpy(np.ascontiguousarray(host2d[:, c0:c1], dtype=np.float32)).to(device)

    def gram(self, M):
        return kernels.gram(M)

    def leading_eig(self, G, k):
        n = G.shape[0]
        if kernels.topk_supported(n, k):
            evals, evecs, info = kernels.eigh_topk(G, k)
            if info["converged"]:
                return evals, evecs
        evals, evecs, _ = kernels.eigh(G)
        return evals[:k].contiguous(), evecs[:k].contiguous()

    def leading_eig_async(self, G, k):
        """Non-synchronising variant: (evals, evecs, record) with record = pinned int32 {iterations, converged},
        valid after the stream has been synchronised; None when the subspace solver does not apply."""
        if not kernels.topk_supported(G.shape[0], k):
            return None
        return kernels.eigh_topk_async(G, k)

    # The principal components travel as ONE stacked tensor [Vhi; Vlo] (2k, p_g): the error-free fp32 pair of the
    # high-precision projection (kernels.project_subtract_hp; why: csrc/proj.cu).  The driver treats it as opaque.
    high_precision = True

    def pcs(self, Wt, M):
        return torch.cat(kernels.pcs_hilo(Wt, M))

    def randomized_pcs(self, M, ncomp, omega, reduce):
        """(V stacked, C): the coefficients M V^T come for free from the sketch (psfsub.svd.randomized_pcs)."""
        from .psfsub.svd import randomized_pcs
        V, Cm = randomized_pcs(M, ncomp, omega, reduce=reduce, hilo=True, coeffs=True)
        return torch.cat(V), Cm

    def coeffs(self, M, V, reduce):
        """C (n,k) fp64 = M (Vhi + Vlo)^T, the pixel axis summed over the shards."""
        k = V.shape[0] // 2
        C2 = reduce(kernels.cross_gram(M, V))                      # (n, 2k)
        return (C2[:, :k] + C2[:, k:]).contiguous()

    def project_subtract(self, M, Cm, V):
        k = V.shape[0] // 2
        return kernels.project_subtract_hp(M, Cm, V[:k], V[k:])

    def derotate(self, cube, angles):
        return derotate_device(cube, angles)

    # ---- exchanges fused into the kernels over peer memory (NVLink), see PeerExchange below
    def project_subtract_rows(self, M, Cm, V, row_ptrs):
        k = V.shape[0] // 2
        kernels.project_subtract_hp_rows(M, Cm, V[:k], V[k:], row_ptrs)

    def derotate_scatter(self, cube, angles, out_bases, rows_per_shard, frame_stride, frame_offset):
        from .preproc.derotation import rotation_geometry, rotation_scalars
        S = cube.shape[-1]
        N, y0 = rotation_geometry(S)
        krot, a, b = rotation_scalars(angles)
        kernels.derotate_scatter(cube.contiguous(), krot, a, b, S, N, y0, out_bases, rows_per_shard, frame_stride,
                                 frame_offset)

    def sdi_stage1(self, cube4d, frames, scale_list, ncomp_ifs, collapse_ifs, device):
        """Stage 1 of the ADI+mSDI double PCA for the ADI frames ``frames`` of a host (z,n,H,W) cube:
        only those frames are uploaded.  Returns (len(frames), H, W) fp32."""
        from ._device import to_device_f32
        from .psfsub.sdi import RescaleOps, _stage1_frames, _CHUNK_BYTES
        z, n, H, W = cube4d.shape
        if (len(frames) > 0 and frames == list(range(frames[0], frames[-1] + 1)) and cube4d.dtype == np.float32
                and cube4d.flags["C_CONTIGUOUS"]):
            # own frames of channel c are one contiguous block of the host cube: z staged copies straight into the
            # device tensor, no host-side gather
            from . import _cabi
            from ._device import stream_ptr
            sub = torch.empty((z, len(frames), H, W), dtype=torch.float32, device=device)
            for c in range(z):
                blk = cube4d[c, frames[0]:frames[-1] + 1]
                _cabi.check(_cabi.lib().vb_memcpy_h2d_staged(int(sub[c].data_ptr()), int(blk.ctypes.data),
                                                              int(blk.nbytes), stream_ptr()), "vb_memcpy_h2d_staged")
        else:
            sub = to_device_f32(np.ascontiguousarray(cube4d[:, frames]), device)      # (z, F, H, W)
        rops = RescaleOps(scale_list, H, device)
        per_frame = 4 * z * rops.big * rops.big * 4 * 2
        chunk = max(1, int(_CHUNK_BYTES // per_frame))
        parts = [_stage1_frames(sub, list(range(c0, min(len(frames), c0 + chunk))), rops, ncomp_ifs, None, None,
                                "lapack", collapse_ifs, (0, z)) for c0 in range(0, len(frames), chunk)]
        return torch.cat(parts)

    def project_subtract_cube(self, cube_dev, ncomp):
        from .psfsub.pca_fullfr import project_subtract_device
        return project_subtract_device(cube_dev, ncomp)

    def collapse(self, cube2d, mode):
        n, p = cube2d.shape
        return collapse_device(cube2d.reshape(n, 1, p), mode).reshape(p)


class StageTimer:
    """CUDA-event marks between the stages of a sharded call (``stage_ms`` of bench.py for N > 1).

    ``mark(name)`` records an event on the current stream AFTER the stage ``name`` has been enqueued; NCCL
    collectives issued through ``torch.distributed`` order the current stream behind them, so the interval between
    two marks is the device time of the stage in between (including its exchange).  ``summary()`` synchronises."""

    def __init__(self, device):
        self.cuda = device.type == "cuda"
        self.marks = []
        self.mark("start")

    def mark(self, name):
        if self.cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append((name, ev))
        else:
            import time
            self.marks.append((name, time.perf_counter()))

    def summary(self):
        out = {}
        if self.cuda:
            torch.cuda.synchronize()
        for (_, a), (name, b) in zip(self.marks[:-1], self.marks[1:]):
            ms = a.elapsed_time(b) if self.cuda else (b - a) * 1e3
            out[name + "_ms"] = out.get(name + "_ms", 0.0) + ms
        return out


def _mark(timer, name):
    if timer is not None:
        timer.mark(name)


def _exchange_single(send_flat, send_counts, recv_counts, group, async_op=False):
    """ONE preallocated all-to-all: ``send_flat`` (1-d, contiguous) holds the blocks for peers 0..world-1 back to
    back (``send_counts`` elements each); returns (recv_flat, handle) with the blocks of the peers back to back
    (``recv_counts``).  Uneven splits are fine (1-d element counts).  No per-peer tensors, no repacking."""
    recv_flat = torch.empty(int(sum(recv_counts)), dtype=send_flat.dtype, device=send_flat.device)
    work = dist.all_to_all_single(recv_flat, send_flat, output_split_sizes=[int(c) for c in recv_counts],
                                  input_split_sizes=[int(c) for c in send_counts], group=group, async_op=async_op)
    return recv_flat, work


def _single_ok():
    """all_to_all_single is NCCL-native; the gloo backend of the CPU tests lacks it on some builds."""
    return os.environ.get("VIP_B200_SHARD_A2A", "single") == "single"


class PeerExchange:
    """Symmetric (peer-mapped) buffers for the two exchanges of ``pca_sharded``, so that they can be FUSED into the
    kernels that produce the data instead of running as separate NCCL all-to-alls:

      * ``frames`` (f_max, p) per rank -- every rank's projection kernel writes residual row i of its pixel shard
        directly into the frame shard of the rank that derotates frame i (``vb_project_subtract_hp_rows_f32``);
      * ``slab`` (n, p_g) per rank -- the last shear pass of every rank writes each derotated image row directly into
        the pixel-shard slab of the rank that takes that pixel's temporal median (``vb_derotate_scatter_f32``).

    The buffers come from ``torch.distributed._symmetric_memory`` (CUDA VMM allocations exchanged between the
    processes of the node, peer access over NVLink / NVSwitch) and are cached per geometry: allocation and rendezvous
    are collective and slow, the steps reuse them.  ``barrier()`` is a device-side barrier on the current stream
    (signal pads in the same allocations): peers have finished writing before a rank reads its buffer, and have
    finished reading before the next step overwrites it.  Measured on 2 x B200 (tools/symm_probe.py): plain peer
    writes move data at 2x the rate of ``all_to_all_single``; fused, the transfer also overlaps the kernel's math."""

    _cache = {}

    def __init__(self, n, H, W, world, rank, fb, pb, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        g = group if group is not None else dist.group.WORLD
        p = H * W
        self.world, self.rank = world, rank
        fmax = int(np.max(np.diff(fb)))
        pg = p // world
        self.frames = symm_mem.empty((fmax, p), dtype=torch.float32, device=device)
        self.frames_h = symm_mem.rendezvous(self.frames, g)
        self.slab = symm_mem.empty((n, pg), dtype=torch.float32, device=device)
        self.slab_h = symm_mem.rendezvous(self.slab, g)
        p0 = int(pb[rank])
        ptrs = np.empty(n, dtype=np.int64)
        for h in range(world):
            base = int(self.frames_h.buffer_ptrs[h])
            i = np.arange(int(fb[h]), int(fb[h + 1]), dtype=np.int64)
            ptrs[int(fb[h]):int(fb[h + 1])] = base + ((i - int(fb[h])) * p + p0) * 4
        self.row_ptrs = torch.from_numpy(ptrs).to(device)
        self.slab_bases = [int(self.slab_h.buffer_ptrs[h]) for h in range(world)]

    def barrier(self):
        self.frames_h.barrier()

    @classmethod
    def eligible(cls, ops, world, n, H, W, collapse, full_output):
        if os.environ.get("VIP_B200_SHARD_FUSED", "1") != "1" or getattr(ops, "name", "") != "cuda":
            return False
        if world < 2 or world > 8 or full_output or H != W:
            return False
        N = 4 * H
        pow2 = (N & (N - 1)) == 0 and 512 <= N <= 4096
        return pow2 and H % world == 0 and (H // world) % 2 == 0 and (H * W) % world == 0 and (H * W // world) % 2 == 0

    @classmethod
    def get(cls, n, H, W, world, rank, fb, pb, device, group):
        key = (n, H, W, world, rank, str(device), id(group))
        if key not in cls._cache:
            cls._cache.clear()               # one geometry at a time: the buffers are as large as the cube shard
            try:
                cls._cache[key] = cls(n, H, W, world, rank, fb, pb, device, group)
            except Exception as exc:                                       # noqa: BLE001
                cls._cache[key] = None
                import warnings
                warnings.warn(f"vip_b200: symmetric memory unavailable ({exc!r}); NCCL exchanges are used")
        return cls._cache[key]


def _all_to_all(send, recv, group):
    """Exchange lists of tensors (uneven sizes allowed)."""
    try:
        dist.all_to_all(recv, send, group=group)
    except (RuntimeError, NotImplementedError):
        # backends without all_to_all (older gloo): pairwise exchange
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        recv[rank].copy_(send[rank])
        reqs = []
        for peer in range(world):
            if peer != rank:
                reqs.append(dist.isend(send[peer], dst=dist.get_global_rank(group, peer) if group else peer,
                                       group=group))
                reqs.append(dist.irecv(recv[peer], src=dist.get_global_rank(group, peer) if group else peer,
                                       group=group))
        for r in reqs:
            r.wait()


class _Exchange:
    """Handle of an all-to-all in flight (``wait()`` orders the caller's stream after it)."""

    def __init__(self, works):
        self.works = works

    def wait(self):
        for w in self.works:
            w.wait()


def _all_to_all_async(send, recv, group):
    """``_all_to_all`` that returns while the exchange is in flight: NCCL runs it on its own stream, ordered after
    the work already enqueued on the caller's stream, so kernels launched afterwards overlap with the transfer."""
    try:
        return _Exchange([dist.all_to_all(recv, send, group=group, async_op=True)])
    except (RuntimeError, NotImplementedError):
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        recv[rank].copy_(send[rank])
        reqs = []
        for peer in range(world):
            if peer != rank:
                g = dist.get_global_rank(group, peer) if group else peer
                reqs.append(dist.isend(send[peer], dst=g, group=group))
                reqs.append(dist.irecv(recv[peer], src=g, group=group))
        return _Exchange(reqs)


def _rows_to_frames(block, fb, pb, rank, group, device, timer=None, async_op=False):
    """Exchange 1: ``block`` (n, p_g) -- every frame, this rank's pixel shard, row-major -- to this rank's frame
    shard with all pixels, (f_g, p).  Row blocks of ``block`` are already contiguous per destination, so the send
    side is zero-copy; the receive side is one flat buffer [peer h][f_g][p_h] that is stitched into (f_g, p) with
    one strided copy per peer (no list of temporaries, no ``cat``).  Returns a callable producing the tensor
    (it waits for the exchange when ``async_op``)."""
    world = dist.get_world_size(group)
    n, pg = block.shape
    f0, f1 = int(fb[rank]), int(fb[rank + 1])
    fg = f1 - f0
    p = int(pb[-1])
    widths = [int(pb[h + 1] - pb[h]) for h in range(world)]
    if _single_ok() and block.is_contiguous():
        try:
            recv_flat, work = _exchange_single(block.reshape(-1), [int(fb[h + 1] - fb[h]) * pg for h in range(world)],
                                               [fg * w for w in widths], group, async_op=async_op)
        except (RuntimeError, NotImplementedError):
            recv_flat = None
        if recv_flat is not None:
            def finish():
                if async_op and work is not None:
                    work.wait()
                if world == 1:
                    return recv_flat.reshape(fg, p)
                out = torch.empty((fg, p), dtype=block.dtype, device=device)
                off = 0
                for h in range(world):
                    out[:, int(pb[h]):int(pb[h + 1])] = recv_flat[off:off + fg * widths[h]].reshape(fg, widths[h])
                    off += fg * widths[h]
                return out
            return finish
    send = [block[int(fb[h]):int(fb[h + 1])].contiguous() for h in range(world)]
    recv = [torch.empty((fg, widths[h]), dtype=block.dtype, device=device) for h in range(world)]
    if async_op:
        ex = _all_to_all_async(send, recv, group)

        def finish_list():
            ex.wait()
            return torch.cat(recv, dim=1)
        return finish_list
    _all_to_all(send, recv, group)
    return lambda: torch.cat(recv, dim=1)


def _frames_to_pixels(der2, fb, pb, rank, group, device):
    """Exchange 2: this rank's derotated frames (f_g, p) -> every frame of this rank's pixel shard, (n, p_g).
    The send buffer is packed per destination with one strided copy per peer into ONE flat buffer; the flat
    receive buffer [peer h][f_h][p_g] IS the (n, p_g) matrix (frame shards are contiguous, in rank order)."""
    world = dist.get_world_size(group)
    fg, p = der2.shape
    n = int(fb[-1])
    p0, p1 = int(pb[rank]), int(pb[rank + 1])
    pg = p1 - p0
    widths = [int(pb[h + 1] - pb[h]) for h in range(world)]
    if _single_ok():
        if world == 1:
            send_flat = der2.reshape(-1)
        else:
            send_flat = torch.empty(fg * p, dtype=der2.dtype, device=device)
            off = 0
            for h in range(world):
                send_flat[off:off + fg * widths[h]].reshape(fg, widths[h]).copy_(der2[:, int(pb[h]):int(pb[h + 1])])
                off += fg * widths[h]
        try:
            recv_flat, _ = _exchange_single(send_flat, [fg * w for w in widths],
                                            [int(fb[h + 1] - fb[h]) * pg for h in range(world)], group)
            return recv_flat.reshape(n, pg)
        except (RuntimeError, NotImplementedError):
            pass
    send = [der2[:, int(pb[h]):int(pb[h + 1])].contiguous() for h in range(world)]
    recv = [torch.empty((int(fb[h + 1] - fb[h]), pg), dtype=der2.dtype, device=device) for h in range(world)]
    _all_to_all(send, recv, group)
    return torch.cat(recv, dim=0)


def _derotate_collapse_frame_shards(mine, angle_list, fb, pb, collapse, group, ops, device, timer=None):
    """Tail of the sharded paths: this rank's residual frames ``mine`` (f1-f0, H, W) -> derotation (local) ->
    temporal collapse across the ranks.  'mean' / 'sum' are reducible (partial sums + one NCCL reduce of an
    (H,W) frame); the other modes need every frame of a pixel on one rank: all-to-all to pixel shards, local
    collapse, gather.  Returns (frame on rank 0 / None elsewhere, own derotated frames)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    src = dist.get_global_rank(group, 0) if group is not None else 0
    n = int(fb[-1])
    p = int(pb[-1])
    H, W = mine.shape[1], mine.shape[2]
    p0, p1 = int(pb[rank]), int(pb[rank + 1])
    f0, f1 = int(fb[rank]), int(fb[rank + 1])
    der = ops.derotate(mine, -angle_list[f0:f1]) if f1 > f0 else mine
    _mark(timer, "derotate")

    if collapse in ("mean", "sum"):
        # ---- reducible collapse: partial sums over the own frames, one NCCL reduce of an (H,W) frame ------
        # (PCA residuals hold no NaNs -- a NaN in the cube would already have poisoned the Gramian -- so
        # nanmean == sum / n here; the non-reducible modes take the all-to-all below)
        part = ops.collapse(der.reshape(f1 - f0, p), "sum") if f1 > f0 else torch.zeros(p, device=device)
        part = part.to(torch.float32).contiguous()
        dist.reduce(part, dst=src, op=dist.ReduceOp.SUM, group=group)
        _mark(timer, "collapse_reduce")
        frame = None
        if rank == 0:
            if collapse == "mean":
                part = part / n
            frame = part.reshape(H, W).cpu().numpy()
        return frame, der

    # ---- exchange 2: frame shards -> pixel shards, collapse, gather --------------------------------
    slab_in = _frames_to_pixels(der.reshape(f1 - f0, p), fb, pb, rank, group, device)      # (n, p_g)
    _mark(timer, "exchange2")
    slab = ops.collapse(slab_in, collapse)                                                   # (p_g,)
    _mark(timer, "collapse")

    frame = None
    if p % world == 0 and _single_ok():
        # equal shards: the slabs land back to back in one (H*W) buffer, no padding / repacking
        full = torch.empty(p, dtype=slab.dtype, device=device)
        try:
            dist.all_gather_into_tensor(full, slab.contiguous(), group=group)
        except (RuntimeError, NotImplementedError, AttributeError):
            full = None
        if full is not None:
            _mark(timer, "gather")
            if rank == 0:
                frame = full.reshape(H, W).cpu().numpy()
            return frame, der
    pmax = int(np.max(np.diff(pb)))
    padded = torch.zeros(pmax, dtype=slab.dtype, device=device)
    padded[: p1 - p0] = slab
    gathered = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded, group=group)
    _mark(timer, "gather")
    if rank == 0:
        frame = torch.cat([gathered[h][: int(pb[h + 1] - pb[h])] for h in range(world)]).reshape(H, W)
        frame = frame.cpu().numpy()
    return frame, der


def pca_sharded(cube, angle_list, ncomp, collapse="median", group=None, ops=None, device=None,
                full_output=False, resident_shard=None, svd_mode="lapack", random_state=None, _defer_check=True,
                host_shard=None, shape=None, overlap_exchange=None, timer=None):
    """Full-frame ADI PCA of ONE cube, sharded over the ranks of ``group``.

    ``cube`` (n,H,W) is the host array, visible on every rank (each rank uploads only its pixel
    shard).  Returns the final frame (H,W) as a numpy array on rank 0 and ``None`` elsewhere; with
    ``full_output`` every rank additionally returns its own (frames, H, W) derotated residual shard
    as a device tensor and the frame offsets.  ``resident_shard``: this rank's pixel shard already on
    the device (skips the upload; used by bench.py for the device-resident number).

    ``svd_mode``: any deterministic mode (exact path: all-reduce of the n x n Gramian) or 'randsvd'
    (BASELINE config 5): scikit-learn's randomized SVD on pixel shards, whose only communication is
    the all-reduce of the (ncomp+10) x n sketches and (ncomp+10)^2 Gramians
    (``psfsub.svd.randomized_pcs``); the Gaussian test matrix is drawn on rank 0 from
    ``random_state`` (default: numpy's global RandomState, as the reference does) and broadcast.

    ``host_shard`` + ``shape``: for cubes that no single host buffer should hold (BASELINE config 5 is 16.8 GB),
    every rank may pass ONLY its own pixel shard -- a C-contiguous fp32 (n, p_g) host array with
    p_g = ``shard_bounds(H*W, world)`` of this rank -- together with ``shape=(n, H, W)``; ``cube`` is then ignored
    (pass None) and the upload is one contiguous copy.

    ``timer``: a :class:`StageTimer`; when given, a CUDA event is recorded after every stage (bench.py ``stage_ms``)."""
    if host_shard is not None and (shape is None or len(shape) != 3):
        raise TypeError("pca_sharded: `host_shard` needs `shape=(n, H, W)`")
    if shape is not None and (host_shard is not None or resident_shard is not None):
        n, H, W = (int(v) for v in shape)
    else:
        if cube.ndim != 3:
            raise TypeError("Input array is not a cube or 3d array")
        n, H, W = cube.shape
    ops = ops or CudaOps()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    p = H * W
    angle_list = check_pa_vector(np.asarray(angle_list))
    if n != angle_list.shape[0]:
        raise ValueError("`angle_list` vector has wrong length. It must equal the number of frames in the cube")
    if not isinstance(ncomp, (int, np.integer)) or ncomp <= 0:
        raise ValueError("pca_sharded needs an integer ncomp > 0")
    ncomp = min(int(ncomp), n)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if ops.name == "cuda" else torch.device("cpu")

    pb = shard_bounds(p, world)      # pixel shards
    fb = shard_bounds(n, world)      # frame shards
    p0, p1 = int(pb[rank]), int(pb[rank + 1])
    f0, f1 = int(fb[rank]), int(fb[rank + 1])

    # ---- pixel-sharded PCA ------------------------------------------------------------------
    if resident_shard is not None:
        M = resident_shard
    elif host_shard is not None:
        if host_shard.shape != (n, p1 - p0) or host_shard.dtype != np.float32 or not host_shard.flags["C_CONTIGUOUS"]:
            raise ValueError(f"pca_sharded: rank {rank} needs a C-contiguous fp32 host shard of shape {(n, p1 - p0)}")
        M = torch.from_numpy(host_shard).to(device, non_blocking=True)
    else:
        M = ops.upload_pixels(cube.reshape(n, p), p0, p1, device)
    src = dist.get_global_rank(group, 0) if group is not None else 0
    _mark(timer, "upload")
    record = None
    raw_frames = None
    peer = None
    if PeerExchange.eligible(ops, world, n, H, W, collapse, full_output):
        peer = PeerExchange.get(n, H, W, world, rank, fb, pb, device, group)
    if overlap_exchange is None:
        # default ON since round 2 (measured on NVLink: profiles/r02_scaling.md); VIP_B200_SHARD_OVERLAP=0 restores
        # the subtract-then-exchange order
        overlap_exchange = os.environ.get("VIP_B200_SHARD_OVERLAP", "1") == "1"
    mode = str(getattr(svd_mode, "value", svd_mode))
    if mode in ("randsvd", "randcupy", "randpytorch"):
        def reduce(t):                                                   # exchange step 0: sketches
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            return t
        ell = ncomp + 10
        if rank == 0:
            rs = random_state
            if rs is None:
                rs = np.random.mtrand._rand
            elif not isinstance(rs, np.random.RandomState):
                rs = np.random.RandomState(rs)
            omega = torch.from_numpy(rs.normal(size=(n, ell))).to(device)
        else:
            omega = torch.empty((n, ell), dtype=torch.float64, device=device)
        dist.broadcast(omega, src=src, group=group)
        out = ops.randomized_pcs(M, ncomp, omega if getattr(ops, "name", "") == "cuda" else omega.cpu().numpy(), reduce)
        _mark(timer, "randsvd")
        if isinstance(out, tuple):
            V, Cm = out                                                  # coefficients for free (B^T Wt^T)
        else:
            V = out
            Cm = ops.coeffs(M, V, reduce)
            _mark(timer, "coeffs")
    else:
        G = ops.gram(M)
        _mark(timer, "gram")
        dist.all_reduce(G, op=dist.ReduceOp.SUM, group=group)           # exchange step 0: n x n fp64
        _mark(timer, "allreduce_gram")
        if overlap_exchange and world > 1 and peer is None:
            # R = M - C V is row-local once V is known everywhere: the all-to-all to frame shards moves the RAW
            # cube while the (replicated, latency-bound) eigensolver runs; V (k x p, 21 MB at config 2) is
            # all-gathered afterwards and the subtraction happens on the frame shards.  Same kernels on the same
            # values: bit-identical to the default order.
            raw_frames = _rows_to_frames(M, fb, pb, rank, group, device, async_op=True)
        # the convergence record of the subspace solver is read once, after the whole pipeline has been
        # enqueued (no host stall behind the eigensolver); see the check before the return
        deferred = ops.leading_eig_async(G, ncomp) if (_defer_check and hasattr(ops, "leading_eig_async")) else None
        if deferred is not None:
            evals, evecs, record = deferred
        else:
            evals, evecs = ops.leading_eig(G, ncomp)
        _mark(timer, "eigensolver")
        if os.environ.get("VIP_B200_SHARD_BCAST", "0") == "1":
            # optional: bit-identical eigenpairs on every rank.  Not needed for the residuals: every rank projects
            # its pixel shard on ITS OWN (E, S) pair consistently (R_g = M_g - E E^T M_g is invariant to the sign
            # and order of the vectors), and all ranks start from the same all-reduced G, so the subspaces agree
            # to the solver's tolerance (1e-9 of lambda_k) -- two latency-bound broadcasts saved per call.
            dist.broadcast(evals, src=src, group=group)
            dist.broadcast(evecs, src=src, group=group)
        S = torch.sqrt(torch.clamp(evals, min=0.0))
        Wt = (evecs / S[:, None]).contiguous()
        Cm = (evecs * S[:, None]).t().contiguous()                     # fp64 (n, k)
        if not getattr(ops, "high_precision", False):
            Cm = Cm.to(torch.float32)
        V = ops.pcs(Wt, M)
        _mark(timer, "pcs")
    if peer is not None:
        # ---- fused exchanges over peer memory: the projection writes every residual row into the frame shard of its
        # owner, the last shear pass writes every derotated image row into the pixel slab of its owner
        peer.barrier()                                   # nobody still reads the buffers of the previous step
        ops.project_subtract_rows(M, Cm, V, peer.row_ptrs)
        peer.barrier()                                   # every rank's rows have landed in my frame shard
        _mark(timer, "subtract_scatter")
        mine = peer.frames[: f1 - f0].reshape(f1 - f0, H, W)
        if collapse in ("mean", "sum"):
            # reducible collapse: partial sums over the own frames + one NCCL reduce (collective: every rank calls it)
            frame, _ = _derotate_collapse_frame_shards(mine, angle_list, fb, pb, collapse, group, ops, device, timer)
        else:
            if f1 > f0:
                ops.derotate_scatter(mine, -angle_list[f0:f1], peer.slab_bases, H // world, p // world, f0)
            peer.barrier()                               # every rank's image rows have landed in my pixel slab
            _mark(timer, "derotate_scatter")
            slab = ops.collapse(peer.slab, collapse)
            _mark(timer, "collapse")
            full = torch.empty(p, dtype=slab.dtype, device=device)
            dist.all_gather_into_tensor(full, slab.contiguous(), group=group)
            _mark(timer, "gather")
            frame = full.reshape(H, W).cpu().numpy() if rank == 0 else None
        der = None
    elif raw_frames is not None:
        # all-gather of the PCs
        k = V.shape[0]
        if p % world == 0 and _single_ok():
            parts_flat = torch.empty((world * k, p // world), dtype=V.dtype, device=device)
            dist.all_gather_into_tensor(parts_flat, V.contiguous(), group=group)
            V_full = parts_flat.reshape(world, k, p // world).permute(1, 0, 2).reshape(k, p)
        else:
            pmax = int(np.max(np.diff(pb)))
            padded = torch.zeros((k, pmax), dtype=V.dtype, device=device)
            padded[:, : p1 - p0] = V
            parts = [torch.empty_like(padded) for _ in range(world)]
            dist.all_gather(parts, padded, group=group)
            V_full = torch.cat([parts[h][:, : int(pb[h + 1] - pb[h])] for h in range(world)], dim=1)
        V_full = V_full.contiguous()
        _mark(timer, "allgather_pcs")
        mine_raw = raw_frames()                                             # my frames, all pixels, raw
        _mark(timer, "exchange1_wait")
        if f1 > f0:
            mine = ops.project_subtract(mine_raw, Cm[f0:f1].contiguous(), V_full).reshape(f1 - f0, H, W)
        else:
            mine = mine_raw.reshape(0, H, W)
        _mark(timer, "subtract")
    else:
        R = ops.project_subtract(M, Cm, V)                                   # (n, p_g)
        _mark(timer, "subtract")
        # ---- exchange 1: pixel shards -> frame shards ---------------------------------------------
        mine = _rows_to_frames(R, fb, pb, rank, group, device)().reshape(f1 - f0, H, W)   # my frames, all pixels
        _mark(timer, "exchange1")
    if peer is None:
        frame, der = _derotate_collapse_frame_shards(mine, angle_list, fb, pb, collapse, group, ops, device, timer)
    if record is not None:
        if device.type == "cuda":
            torch.cuda.current_stream().synchronize()
        ok = torch.tensor([int(record[1])], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)            # one decision for all ranks
        if int(ok.item()) == 0:
            # rare: the subspace iteration stalled -> redo with the synchronous solver (Jacobi fallback)
            return pca_sharded(cube, angle_list, ncomp, collapse=collapse, group=group, ops=ops, device=device,
                               full_output=full_output, resident_shard=resident_shard, svd_mode=svd_mode,
                               random_state=random_state, _defer_check=False, host_shard=host_shard, shape=shape,
                               overlap_exchange=overlap_exchange, timer=timer)
    if full_output:
        return frame, der, (f0, f1)
    return frame


def pca_adimsdi_double_sharded(cube, angle_list, scale_list, ncomp, collapse="median", collapse_ifs="mean",
                               group=None, ops=None, device=None, full_output=False):
    """ADI+mSDI double-pass PCA of ONE IFS cube (z,n,H,W), sharded BY ADI FRAME over the ranks (BASELINE
    config 4): ``pca(cube, angs, scale_list=s, adimsdi='double', ncomp=(k_ifs, k_adi))``.

        stage 1 (per ADI frame: rescale the z channels, PCA across them, descale, collapse)    local, own frames
        all-gather of the stage-1 frames (n x H x W, 78 MB at config 4)                          NCCL
        stage 2 ADI PCA on the gathered (n, H*W) matrix (300 x 65 536: tiny)                     replicated
        derotation of the own frames, temporal collapse                          reduce / all-to-all + gather

    Each rank uploads only its own frames of the host cube.  Returns the final frame on rank 0 (None
    elsewhere); with ``full_output`` also the (n,H,W) stage-1 cube and the own derotated frames."""
    if cube.ndim != 4:
        raise TypeError("Input cube is not a 4d array (required with `scale_list`)")
    if not isinstance(ncomp, tuple) or len(ncomp) != 2:
        raise TypeError("`ncomp` must be a tuple when a double pass PCA is performed")
    ops = ops or CudaOps()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    z, n, H, W = cube.shape
    k_ifs, k_adi = ncomp
    angle_list = check_pa_vector(np.asarray(angle_list))
    if angle_list.shape[0] != n:
        raise ValueError("Angle list vector has wrong length. It must equal the number frames in the cube")
    scale_list = np.asarray(scale_list)
    if scale_list.ndim != 1 or scale_list.shape[0] != z:
        raise ValueError("Scaling factors vector has wrong length")
    if k_ifs is not None:
        k_ifs = min(int(k_ifs), z)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if ops.name == "cuda" else torch.device("cpu")
    fb = shard_bounds(n, world)
    pb = shard_bounds(H * W, world)
    f0, f1 = int(fb[rank]), int(fb[rank + 1])

    own = ops.sdi_stage1(cube, list(range(f0, f1)), scale_list, k_ifs, collapse_ifs, device) if f1 > f0 \
        else torch.zeros((0, H, W), dtype=torch.float32, device=device)
    fmax = int(np.max(np.diff(fb)))
    padded = torch.zeros((fmax, H, W), dtype=torch.float32, device=device)
    padded[: f1 - f0] = own
    gathered = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded, group=group)                              # exchange: stage-1 frames
    res_channels = torch.cat([gathered[h][: int(fb[h + 1] - fb[h])] for h in range(world)])     # (n, H, W)

    if k_adi is None:
        res2 = res_channels
    else:
        res2 = ops.project_subtract_cube(res_channels, min(int(k_adi), n))       # replicated, 300 x 65 536
    frame, der = _derotate_collapse_frame_shards(res2[f0:f1].contiguous(), angle_list, fb, pb, collapse, group,
                                                 ops, device)
    if full_output:
        return frame, res_channels, der
    return frame
