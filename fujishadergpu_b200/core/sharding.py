"""Row-band sharding of one raster over the GPUs of a node (one process per GPU, torch.distributed).

The reference scales out with Dask ``map_overlap`` halos shipped over TCP between worker processes
(reference: algorithms/_impl_topousm_fast.py:205-214, core/dask_cluster.py:88-101).  Here the raster is
cut into contiguous row bands (boundaries on multiples of 16 rows, so no decimation cell straddles
two GPUs) and neighbouring bands exchange only what each stage needs, point to point over NVLink:

    stage                 exchanged rows (per side)                      bytes at W = 65536
    fused full-res pass   max fused radius (<= 41) DEM rows              8 MiB
    level-f box mean      r/f + 1 rows of the f x f decimated grid       2-4 MiB (level 4 / 16)
    stats pre-pass        window rows gathered to the rank that owns     270 MiB per 8256^2 window
                          the window (9 windows, round-robin)

The pipeline is semantically the single-block one (global decimation anchors, global zoom mapping,
edge rules at the global raster edges): the result is bit-identical to one GPU processing the whole
raster.  No collective touches image data except the point-to-point halo copies; the percentile of
the statistics pre-pass is an exact distributed radix select (all-reduce of 2048-bin histograms).

`backend` abstracts the compute stages so the host logic can be exercised on CPU (gloo) by tests that
inject their own stand-in (tests/test_sharding_gloo.py); the product backend is CudaBackend
(libfsg_b200) and there is no other one in this package.
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

ALIGN = 16  # band boundaries are multiples of the largest decimation factor


# ------------------------------------------------------------------------------------------------
# geometry
# ------------------------------------------------------------------------------------------------
def band_bounds(H: int, world: int, align: int = ALIGN) -> List[Tuple[int, int]]:
    """Contiguous [r0, r1) per rank, boundaries on multiples of `align`, sizes as even as possible."""
    cells = (int(H) + align - 1) // align
    base, rem = divmod(cells, int(world))
    out, c = [], 0
    for i in range(int(world)):
        n = base + (1 if i < rem else 0)
        r0, r1 = min(H, c * align), min(H, (c + n) * align)
        out.append((r0, r1))
        c += n
    return out


def mirror_need(lo: int, hi: int, n: int, reflect: bool) -> Tuple[int, int]:
    """Rows of a length-n axis touched by indices [lo, hi] after mirroring (edge-inclusive 'reflect') or clamping."""
    if n <= 0:
        return (0, 0)
    a, b = max(lo, 0), min(hi, n - 1)
    if reflect:
        if lo < 0:
            b = max(b, min(n - 1, -lo - 1))
            if -lo - 1 > n - 1:
                a, b = 0, n - 1        # window re-reflects: everything
        if hi > n - 1:
            a = min(a, max(0, 2 * n - 1 - hi))
            if 2 * n - 1 - hi < 0:
                a, b = 0, n - 1
    return (a, b + 1)  # half-open


def _overlap(a, b):
    lo, hi = max(a[0], b[0]), min(a[1], b[1])
    return (lo, hi) if hi > lo else None


def plan_exchange(local: torch.Tensor, own: Sequence[Tuple[int, int]], need: Sequence[Tuple[int, int]], rank: int,
                  dist=None, out: Optional[torch.Tensor] = None):
    """First half of exchange_rows: allocates / fills the own part of the result and returns
    (out, p2p ops, buffers to keep alive).  Several plans can be sent as ONE batch (run_exchanges): a group
    launch costs ~0.1 ms, which dominated the nine window gathers of the statistics pre-pass."""
    lo, hi = need[rank]
    world = len(own)
    if out is None and hi > lo and own[rank][0] <= lo and hi <= own[rank][1] and not any(
            _overlap(own[rank], need[q]) for q in range(world) if q != rank):
        # everything this rank wants is already here and nobody else wants its rows: a view, no copy, no traffic
        return local[lo - own[rank][0]:hi - own[rank][0]], [], []
    if out is None:
        out = torch.empty((max(0, hi - lo),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    else:
        assert tuple(out.shape) == (max(0, hi - lo),) + tuple(local.shape[1:]) and out.dtype == local.dtype
    mine = _overlap(own[rank], need[rank])
    if mine:
        dst = out[mine[0] - lo:mine[1] - lo]
        src = local[mine[0] - own[rank][0]:mine[1] - own[rank][0]]
        if not (dst.data_ptr() == src.data_ptr() and dst.stride() == src.stride()):
            dst.copy_(src)
    ops, keep = [], []
    if dist is not None and world > 1:
        for q in range(world):
            if q == rank:
                continue
            snd = _overlap(own[rank], need[q])
            if snd:
                buf = local[snd[0] - own[rank][0]:snd[1] - own[rank][0]].contiguous()
                keep.append(buf)
                ops.append(dist.P2POp(dist.isend, buf, q))
            rcv = _overlap(own[q], need[rank])
            if rcv:
                view = out[rcv[0] - lo:rcv[1] - lo]
                if not view.is_contiguous():
                    raise ValueError("exchange_rows: the receive view must be contiguous")
                ops.append(dist.P2POp(dist.irecv, view, q))
    return out, ops, keep


EXCHANGE_LOG: Optional[dict] = None   # accounting (config 5 report): set to {"sent": 0, "received": 0, "events": []}


def run_exchanges(plans, dist=None) -> list:
    """Issue the point-to-point operations of several plan_exchange() results as one batch; returns the outs.
    Every rank must pass the plans in the same order (sends and receives between a pair match in issue order)."""
    ops = [op for (_out, o, _keep) in plans for op in o]
    if ops and dist is not None:
        log = EXCHANGE_LOG
        if log is not None:
            for op in ops:
                nbytes = op.tensor.numel() * op.tensor.element_size()
                log["sent" if op.op is dist.isend else "received"] += nbytes
            if op.tensor.is_cuda:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(torch.cuda.current_stream(op.tensor.device))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        if log is not None and ops[-1].tensor.is_cuda:
            e1.record(torch.cuda.current_stream(ops[-1].tensor.device))
            log["events"].append((e0, e1))
    return [out for (out, _o, _keep) in plans]


def exchange_rows(local: torch.Tensor, own: Sequence[Tuple[int, int]], need: Sequence[Tuple[int, int]], rank: int,
                  dist=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Every rank holds rows own[rank] of a global 2-D array and wants rows need[rank]; returns the
    tensor covering need[rank].  Point-to-point only (NCCL send/recv over NVLink, or gloo on CPU).
    `out`: caller's buffer for need[rank]; when `local` already is the matching slice of `out` (a band kept
    inside its halo buffer, see haloed_band) the own rows are not copied, only the halo rows arrive."""
    return run_exchanges([plan_exchange(local, own, need, rank, dist, out)], dist)[0]


# ------------------------------------------------------------------------------------------------
# compute backend (product): libfsg_b200
# ------------------------------------------------------------------------------------------------
class CudaBackend:
    def plan(self, radii, pixel_size):
        from .. import kernels as k
        return k.topousm_plan(radii, pixel_size)

    def pyramid(self, band, factors):
        from .. import kernels as k
        return k.pyramid_band(band, factors)

    def grid_mean(self, src, src_row0, gh, size, out_row0, out_rows):
        from .. import kernels as k
        return k.grid_mean_band(src, src_row0, gh, size, out_row0, out_rows)

    def void_fill(self, grid, inplace=False):
        from .. import kernels as k
        return k.grid_void_fill(grid, inplace=inplace)

    def fused(self, dem_ext, dem_row0, H, out_row0, out_rows, **kw):
        from .. import kernels as k
        return k.topousm_fused_band(dem_ext, dem_row0, H, out_row0, out_rows, **kw)


# ------------------------------------------------------------------------------------------------
# sharded topousm_fast
# ------------------------------------------------------------------------------------------------
def dem_halo_rows(H: int, world: int, rank: int, radii, pixel_size=1.0, backend=None) -> Tuple[int, int]:
    """Global rows [lo, hi) of the DEM the fused pass of `rank` reads (own band + mirrored halo)."""
    backend = backend or CudaBackend()
    plan = backend.plan(radii, pixel_size)
    kinds, sizes = plan["kind"], plan["size"]
    halo = plan["fused_halo"]
    for i in range(len(kinds)):
        if kinds[i] == 2:
            halo = max(halo, 4 if sizes[i] == 0 else sizes[i] // 2)
    a, b = band_bounds(H, world)[rank]
    return mirror_need(a - halo, b - 1 + halo, H, reflect=True) if b > a else (0, 0)


def haloed_band(H: int, W: int, world: int, rank: int, radii, pixel_size=1.0, device="cuda", backend=None,
                peer_group=None):
    """(ext, band): `ext` holds the rows dem_halo_rows() names, `band` is the view of this rank's own rows inside
    it.  A loader that writes its rows into `band` and passes `dem_ext=ext` to topousm_fast_sharded spares the
    per-call copy of the band into a halo buffer (2 x band bytes of HBM traffic per step).
    peer_group (a NCCL process group, every rank calls): `ext` is allocated in symmetric memory, so that the other
    ranks of the node can read this band directly over NVLink (PeerBands; the window gather of the statistics
    pre-pass then needs no send / receive pairs).  Falls back to a private buffer when symmetric memory is not
    available."""
    lo, hi = dem_halo_rows(H, world, rank, radii, pixel_size, backend)
    a, b = band_bounds(H, world)[rank]
    if peer_group is not None and world > 1:
        pb = PeerBands.create(H, W, world, rank, radii, pixel_size, device, peer_group, backend)
        if pb is not None:
            return pb.ext, pb.ext[a - lo:b - lo]
    ext = torch.empty((hi - lo, W), dtype=torch.float32, device=device)
    return ext, ext[a - lo:b - lo]


class PeerBands:
    """The DEM row bands of all ranks of one node, each in symmetric memory (torch.distributed._symmetric_memory:
    CUDA virtual-memory handles exchanged at rendezvous, every rank maps every band).  band_of(q) is rank q's band as
    a tensor on THIS device whose loads travel over NVLink / NVSwitch; barrier() orders the ranks on the stream
    (8 us).  Found again from the halo buffer by of()."""

    _by_ptr: dict = {}

    def __init__(self, ext, hdl, views, rank, world):
        self.ext, self.hdl, self.views, self.rank, self.world = ext, hdl, views, rank, world

    @classmethod
    def create(cls, H, W, world, rank, radii, pixel_size, device, group, backend=None):
        try:
            import torch.distributed._symmetric_memory as symm
            spans = [dem_halo_rows(H, world, q, radii, pixel_size, backend) for q in range(world)]
            rows = max(hi - lo for lo, hi in spans)          # symmetric allocations have one size on every rank
            buf = symm.empty((rows, W), dtype=torch.float32, device=device)
            hdl = symm.rendezvous(buf, group)
            own = band_bounds(H, world)
            views = []
            for q in range(world):
                lo_q, _hi_q = spans[q]
                whole = buf if q == rank else hdl.get_buffer(q, (rows, W), torch.float32)
                views.append(whole[own[q][0] - lo_q:own[q][1] - lo_q])
            lo, hi = spans[rank]
            pb = cls(buf[: hi - lo], hdl, views, rank, world)
            cls._by_ptr[pb.ext.data_ptr()] = pb
            return pb
        except Exception as e:   # no symmetric memory on this system: the point-to-point path is used instead
            import warnings
            warnings.warn(f"symmetric memory unavailable ({e!r}); window gather falls back to send / receive")
            return None

    @classmethod
    def of(cls, dem_ext):
        return cls._by_ptr.get(dem_ext.data_ptr()) if isinstance(dem_ext, torch.Tensor) and dem_ext.is_cuda else None

    def band_of(self, q: int) -> torch.Tensor:
        return self.views[q]

    def barrier(self):
        self.hdl.barrier()


def topousm_fast_sharded(band: torch.Tensor, H: int, rank: int, world: int, *, radii, weights=None, pixel_size=1.0,
                         norm_scale=None, output_dtype="float32", qp=None, dist=None, backend=None,
                         out: Optional[torch.Tensor] = None, dem_ext: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`band` = this rank's rows band_bounds(H, world)[rank] of the H x W raster (f32, NaN = NoData).
    Returns the same rows of the topousm_fast result (normalised by norm_scale when given).
    dem_ext: optional halo buffer from haloed_band() that already contains `band` (no copy of the own rows)."""
    prep = topousm_sharded_prepare(band, H, rank, world, radii=radii, pixel_size=pixel_size, dist=dist, backend=backend,
                                   dem_ext=dem_ext)
    return topousm_sharded_finish(prep, weights=weights, norm_scale=norm_scale, output_dtype=output_dtype, qp=qp, out=out)


def topousm_sharded_prepare(band: torch.Tensor, H: int, rank: int, world: int, *, radii, pixel_size=1.0, dist=None,
                            backend=None, dem_ext: Optional[torch.Tensor] = None, spec: Optional["Speculation"] = None) -> dict:
    """Everything of the sharded main pass that does not depend on the normalisation scale: DEM halo rows,
    pyramid levels, their halo rows, the coarse means.  topousm_sharded_finish() runs the fused pass.
    spec: plan with the remembered "level holds an all-NoData cell" flags instead of reading this step's flags
    back (the read-back is checked after the step; see Speculation)."""
    backend = backend or CudaBackend()
    W = int(band.shape[1])
    own = band_bounds(H, world)
    r0, r1 = own[rank]
    assert int(band.shape[0]) == r1 - r0, "band does not match band_bounds()"
    plan = backend.plan(radii, pixel_size)
    kinds, factors, sizes, R = plan["kind"], plan["factor"], plan["size"], plan["fused_halo"]
    n = len(kinds)

    # ---- 1. DEM halo for the fused pass (+ the sigma-1 Gaussian / fallback boxes of full-res planes)
    halo = R
    for i in range(n):
        if kinds[i] == 2:
            halo = max(halo, 4 if sizes[i] == 0 else sizes[i] // 2)
    dem_need = []
    for (a, b) in own:
        if b > a:
            lo, hi = mirror_need(a - halo, b - 1 + halo, H, reflect=True)
            dem_need.append((lo, hi))
        else:
            dem_need.append((0, 0))
    dem_plan = plan_exchange(band, own, dem_need, rank, dist, out=dem_ext)
    dem_row0 = dem_need[rank][0]

    # ---- 2. pyramid levels of the own rows, halo rows of each level, coarse means.  The DEM halo and the halo
    # rows of every level travel in ONE point-to-point batch (a group launch costs ~0.1 ms).
    levels = sorted({factors[i] for i in range(n) if kinds[i] == 1})
    term_grids: List[Optional[torch.Tensor]] = [None] * n
    term_grow0: List[int] = [0] * n
    plans = [dem_plan]
    per_level = []
    if levels:
        if r1 > r0:
            grids, flags = backend.pyramid(band, levels)
        else:
            grids = [torch.empty((0, (W + f - 1) // f), dtype=torch.float32, device=band.device) for f in levels]
            flags = torch.zeros(len(levels), dtype=torch.int32, device=band.device)
        single = not (dist is not None and world > 1)
        if single:
            # one rank holds the whole level: the void fill is gated by a device-side flag, no host round trip
            void = [1] * len(levels)
        else:
            void = flags.clone().to(torch.int32)
            dist.all_reduce(void, op=dist.ReduceOp.MAX)
            if spec is not None:
                void = spec.guess(("void", H, W, world, tuple(levels)), void, [0] * len(levels))
            else:
                void = void.cpu().tolist()
        for li, f in enumerate(levels):
            gh, gw = (H + f - 1) // f, (W + f - 1) // f
            g_own = [((a + f - 1) // f if b > a else 0, (b + f - 1) // f if b > a else 0) for (a, b) in own]
            rscale = (gh - 1) / (H - 1) if H > 1 else 1.0
            mean_rows = []   # rows of the MEAN grid each rank's bilinear taps touch
            for (a, b) in own:
                if b > a:
                    lo = int(math.floor(a * rscale))
                    hi = min(gh - 1, int(math.floor((b - 1) * rscale)) + 1)
                    mean_rows.append((lo, hi + 1))
                else:
                    mean_rows.append((0, 0))
            terms = [i for i in range(n) if kinds[i] == 1 and factors[i] == f]
            if void[li]:
                # a coarse cell is entirely NoData: the enclosed-void fill is a Gaussian over ~1/16 of the
                # raster side, so gather the (small) level once and fill it whole on every rank
                g_need = [(0, gh)] * world
            else:
                reach = max((4 if sizes[i] == 0 else sizes[i] // 2) for i in terms)
                g_need = [mirror_need(lo - reach, hi - 1 + reach, gh, reflect=True) if hi > lo else (0, 0)
                          for (lo, hi) in mean_rows]
            plans.append(plan_exchange(grids[li], g_own, g_need, rank, dist))
            per_level.append((gh, terms, mean_rows[rank], g_need[rank][0], bool(void[li])))
    outs = run_exchanges(plans, dist)
    dem_ext = outs[0]
    for (gh, terms, (lo, hi), g_row0, is_void), g_ext in zip(per_level, outs[1:]):
        if is_void:
            g_ext = backend.void_fill(g_ext, True) if isinstance(backend, CudaBackend) and world == 1 else backend.void_fill(g_ext)
        for i in terms:
            if hi > lo:
                term_grids[i] = backend.grid_mean(g_ext, g_row0, gh, sizes[i], lo, hi - lo)
                term_grow0[i] = lo
    # full-resolution planes (radius <= 1 -> sigma-1 Gaussian; general fallback boxes)
    for i in range(n):
        if kinds[i] == 2 and r1 > r0:
            term_grids[i] = backend.grid_mean(dem_ext, dem_row0, H, sizes[i], r0, r1 - r0)
            term_grow0[i] = r0

    return dict(backend=backend, band=band, dem_ext=dem_ext, dem_row0=dem_row0, H=H, W=W, r0=r0, r1=r1, radii=radii,
                pixel_size=pixel_size, term_grids=term_grids, term_grow0=term_grow0)


def topousm_sharded_finish(prep: dict, *, weights=None, norm_scale=None, output_dtype="float32", qp=None,
                           out: Optional[torch.Tensor] = None, norm_scale_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """---- 3. fused pass over the own rows (the only part that needs the scale)."""
    r0, r1, W, band = prep["r0"], prep["r1"], prep["W"], prep["band"]
    if r1 <= r0:
        return torch.empty((0, W), dtype=band.dtype if output_dtype == "float32" else getattr(torch, output_dtype),
                           device=band.device)
    return prep["backend"].fused(prep["dem_ext"], prep["dem_row0"], prep["H"], r0, r1 - r0, radii=prep["radii"],
                                 weights=weights, pixel_size=prep["pixel_size"], term_grids=prep["term_grids"],
                                 term_grow0=prep["term_grow0"], norm_scale=norm_scale, output_dtype=output_dtype,
                                 qp=qp, out=out, **({"norm_scale_dev": norm_scale_dev} if norm_scale_dev is not None else {}))


def topousm_fast_sharded_step(band: torch.Tensor, H: int, rank: int, world: int, *, radii, weights=None,
                              pixel_size=1.0, output_dtype="float32", qp=None, dist=None,
                              out: Optional[torch.Tensor] = None, dem_ext: Optional[torch.Tensor] = None):
    """One step (statistics pre-pass + main pass) enqueued WITHOUT any host synchronisation: the p99 scale stays on
    the device (the fused pass reads it there), the data-dependent planning values are speculated (Speculation).
    -> (out, scale_dev, spec): scale_dev is one f32 (NaN = no valid statistics, output not normalised); the caller
    must call spec.ok() once the step is enqueued and repeat the step when it returns False."""
    dev = band.device
    spec = Speculation()
    scale_dev = torch.empty(1, dtype=torch.float32, device=dev)
    cur = torch.cuda.current_stream(dev)
    key = (dev.index, "prep")
    if key not in _PREP_STREAMS:
        _PREP_STREAMS[key] = torch.cuda.Stream(device=dev)
    side = _PREP_STREAMS[key]
    trace = STEP_TRACE is not None
    if trace:
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        evs[0].record(cur)
    side.wait_stream(cur)
    box = {}

    def start_prep():
        with torch.cuda.stream(side):
            box["prep"] = topousm_sharded_prepare(band, H, rank, world, radii=radii, pixel_size=pixel_size, dist=dist,
                                                  dem_ext=dem_ext, spec=spec)
            if trace:
                evs[1].record(side)

    peers = PeerBands.of(dem_ext) if (dist is not None and world > 1) else None
    start_prep.side_stream = side
    sharded_topousm_scale(band, H, rank, world, radii=radii, weights=weights, pixel_size=pixel_size, dist=dist,
                          spec=spec, scale_out=scale_dev, after_gather=start_prep, peers=peers)
    prep = box["prep"]
    if trace:
        evs[2].record(cur)
    cur.wait_stream(side)
    for t in [prep["dem_ext"]] + [g for g in prep["term_grids"] if g is not None]:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            t.record_stream(cur)
    res = topousm_sharded_finish(prep, weights=weights, norm_scale=None, norm_scale_dev=scale_dev,
                                 output_dtype=output_dtype, qp=qp, out=out)
    if peers is not None:
        peers.barrier()   # every rank has read what it needed of the other bands: they may be refilled now
    if trace:
        evs[3].record(cur)
        STEP_TRACE.append(evs)
    return res, scale_dev, spec


PREPASS_TRACE: Optional[list] = None   # debugging: [(phase name, event)] of the statistics pre-pass
STEP_TRACE: Optional[list] = None   # debugging: set to [] to collect (start, prep done, scale done, fused done) events


def step_trace_summary() -> dict:
    """Milliseconds from the start of a step to: preparation done (side stream), scale done, fused pass done."""
    torch.cuda.synchronize()
    rows = [[e[0].elapsed_time(x) for x in e[1:]] for e in (STEP_TRACE or [])]
    if not rows:
        return {}
    n = len(rows)
    return {"steps": n, "prep_done_ms": sum(r[0] for r in rows) / n, "scale_done_ms": sum(r[1] for r in rows) / n,
            "fused_done_ms": sum(r[2] for r in rows) / n}


def topousm_fast_sharded_with_stats(band: torch.Tensor, H: int, rank: int, world: int, *, radii, weights=None,
                                    pixel_size=1.0, output_dtype="float32", qp=None, dist=None,
                                    out: Optional[torch.Tensor] = None, dem_ext: Optional[torch.Tensor] = None,
                                    backend=None, block_fn=None, select_fns=None):
    """Statistics pre-pass + main pass of one step.  The scale-independent part of the main pass (halo exchange,
    pyramid, coarse means) is enqueued on a side stream FIRST, so it fills the device while the pre-pass sits in
    its latency-bound stages (bounding box, window gather, the selection's collectives); the fused pass starts
    when both are done.  Every rank issues its NCCL calls in the same program order.  -> (out, scale)"""
    dev = band.device
    if dev.type != "cuda":   # host-logic tests (gloo, stand-in compute stages): same stages, one after the other
        scale = sharded_topousm_scale(band, H, rank, world, radii=radii, weights=weights, pixel_size=pixel_size, dist=dist,
                                      block_fn=block_fn, select_fns=select_fns)
        return topousm_fast_sharded(band, H, rank, world, radii=radii, weights=weights, pixel_size=pixel_size,
                                    norm_scale=scale, output_dtype=output_dtype, qp=qp, dist=dist, out=out,
                                    dem_ext=dem_ext, backend=backend), scale
    for _attempt in range(3):
        res, scale_dev, spec = topousm_fast_sharded_step(band, H, rank, world, radii=radii, weights=weights,
                                                         pixel_size=pixel_size, output_dtype=output_dtype, qp=qp,
                                                         dist=dist, out=out, dem_ext=dem_ext)
        if spec.ok():
            break
    scale = float(scale_dev.item())
    return res, (scale if scale == scale else None)


_PREP_STREAMS: dict = {}


def _side_stream(device, name, priority: int = 0):
    key = (device.index, name)
    if key not in _PREP_STREAMS:
        _PREP_STREAMS[key] = torch.cuda.Stream(device=device, priority=priority)
    return _PREP_STREAMS[key]


# ------------------------------------------------------------------------------------------------
# distributed exact percentile (np.percentile, method 'linear', f32 sample spread over the ranks)
# ------------------------------------------------------------------------------------------------
def _np_lerp_percentile(lo: float, hi: float, n: int, q: float) -> Tuple[int, object]:
    q32 = np.true_divide(q, np.float32(100))
    vi = (n - 1) * q32
    prev = min(max(int(np.floor(vi)), 0), n - 1)
    gamma = np.asanyarray(vi - np.floor(vi), dtype=np.asanyarray(vi).dtype)[()]
    return prev, gamma


class PeerExchange:
    """The selection's exchange slots of every rank in symmetric memory (see fsg_select_peer_publish / _reduce)."""

    _by_dev: dict = {}

    def __init__(self, buf, hdl, world):
        self.buf, self.hdl, self.world = buf, hdl, world
        self.my_slots = int(buf.data_ptr())
        self.peer_slots_dev = int(hdl.buffer_ptrs_dev)

    @classmethod
    def get(cls, device, group, world):
        key = (device.index, id(group))
        if key not in cls._by_dev:
            try:
                import torch.distributed._symmetric_memory as symm
                from .. import _lib
                words = int(_lib.load().fsg_select_peer_slot_words())
                buf = symm.empty((words,), dtype=torch.int64, device=device)
                buf.zero_()
                hdl = symm.rendezvous(buf, group)
                cls._by_dev[key] = cls(buf, hdl, world)
            except Exception as e:
                import warnings
                warnings.warn(f"symmetric memory unavailable ({e!r}); the selection uses all-reduce calls")
                cls._by_dev[key] = None
        return cls._by_dev[key]

    def barrier(self):
        self.hdl.barrier()


def distributed_percentile(chunks, q: float, *, take_abs: bool, finite_only: bool, device, dist=None,
                           hist_fn=None, rank_info_fn=None, key_to_float=None, scale_out=None, peer_exchange=None):
    """np.percentile over the union of every rank's `chunks`; all ranks return the same float.
    Exact: 3-level radix select on order-preserving keys, histograms summed with all_reduce."""
    multi = dist is not None and dist.is_initialized() and dist.get_world_size() > 1
    if hist_fn is None:
        # product path: the selection state stays on the device, the host only enqueues the stages and the
        # all-reduces of the exchange area (one host synchronisation per percentile instead of six)
        from .. import kernels as k

        def all_reduce(t, op):
            dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MIN)

        return k.staged_percentile(chunks, q, take_abs=take_abs, finite_only=finite_only, device=device,
                                   all_reduce=all_reduce if multi else None, scale_out=scale_out,
                                   peer_exchange=peer_exchange if multi else None)

    def allsum(t):
        if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    prefix, mask, rank_k, n = 0, 0, None, None
    for level, (shift, bins) in enumerate(((21, 2048), (10, 2048), (0, 1024))):
        hist, cnt = hist_fn(level, prefix, mask)
        both = allsum(torch.cat([hist, cnt.to(hist.dtype)]))
        h = both[:2048].cpu().numpy()
        if level == 0:
            n = int(both[2048].item())
            if n == 0:
                return float("nan")
            rank_k, gamma = _np_lerp_percentile(0.0, 0.0, n, q)
            remaining = rank_k
        csum = np.cumsum(h[:bins])
        b = int(np.searchsorted(csum, remaining, side="right"))
        b = min(b, bins - 1)
        remaining -= int(csum[b - 1]) if b > 0 else 0
        prefix |= b << shift
        mask |= (bins - 1) << shift
    key_k = prefix
    info = rank_info_fn(key_k)
    le = allsum(info[:1].clone())
    nxt = info[1:2].clone()
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(nxt, op=dist.ReduceOp.MIN)
    a_k = np.float32(key_to_float(key_k))
    a_k1 = a_k
    if rank_k + 1 < n and int(le.item()) < rank_k + 2:
        a_k1 = np.float32(key_to_float(int(nxt.item())))
    diff = a_k1 - a_k
    out = a_k + diff * gamma
    if gamma >= 0.5:
        out = a_k1 - diff * (1 - gamma)
    return float(out)


# ------------------------------------------------------------------------------------------------
# planning without host round trips
# ------------------------------------------------------------------------------------------------
class Speculation:
    """A few small, data-dependent integers steer the host-side planning of a step (the valid-data bounding box
    picks the statistics windows, the all-NoData-cell flags pick the halo geometry of a pyramid level).  Reading
    them back costs a device synchronisation in the middle of the step.  Instead the step is planned with the values
    the previous step saw (first step: the dense-raster values), the device values are copied to pinned memory
    asynchronously, and ok() compares them after everything has been enqueued: a mismatch updates the memory and the
    caller repeats the step (at most once per change of the data).  The compared values are all-reduced, so every
    rank takes the same decision."""

    _memory: dict = {}
    _pinned: dict = {}

    def __init__(self):
        self.pending = []

    def guess(self, key, dev_values: torch.Tensor, default):
        val = self.peek(key, default)
        self.verify(key, dev_values, val)
        return val

    def peek(self, key, default):
        """The value the planning uses (host only)."""
        return list(Speculation._memory.get(key, default))

    def verify(self, key, dev_values: torch.Tensor, val):
        """Enqueue the copy of the device values that ok() compares with `val` (on the current stream)."""
        dev_values = dev_values.to(torch.int32).contiguous()
        n = int(dev_values.numel())
        slot = (key, len(self.pending))
        pool = Speculation._pinned.setdefault(slot, [])
        host = pool.pop() if pool else torch.empty(n, dtype=torch.int32, pin_memory=True)
        host.copy_(dev_values, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev_values.device))
        self.pending.append((key, slot, host, ev, list(val)))

    def ok(self) -> bool:
        good = True
        for key, slot, host, ev, val in self.pending:
            ev.synchronize()
            seen = [int(v) for v in host.tolist()]
            Speculation._pinned.setdefault(slot, []).append(host)
            if seen != [int(v) for v in val]:
                Speculation._memory[key] = seen
                good = False
        self.pending = []
        return good


# ------------------------------------------------------------------------------------------------
# sharded statistics pre-pass (reference: algorithms/_norm_stats.py:176-298)
# ------------------------------------------------------------------------------------------------
_OWNER_MEMO: dict = {}


def assign_window_owners(wins, own) -> List[int]:
    """Owner rank of every statistics window: the assignment that lets the slowest rank finish first and, among
    those, moves the fewest rows.  Cost of a window on a rank = one window evaluation + the time to fetch the rows
    the rank lacks (measured on B200 / NVLink 5: 0.56 ms per 8256^2 window, ~500 GB/s per reader), in steps of 5 % of
    a window so that bands differing by a few rows tie.  Exact (branch and bound: a window goes to a rank holding
    part of it or to the least loaded of the others -- those are interchangeable).  Deterministic, identical on
    every rank; the result of the pre-pass does not depend on it."""
    key = (tuple(tuple(w) for w in wins), tuple(tuple(o) for o in own))
    if key in _OWNER_MEMO:
        return list(_OWNER_MEMO[key])
    world, n = len(own), len(wins)
    T_WIN_PER_PX = 0.56e-3 / (8256.0 * 8256.0)
    BYTES_PER_S = 500e9

    def cost(wi, q):
        wy0, _wx0, tw, th = wins[wi]
        ov = _overlap(own[q], (wy0, wy0 + th))
        have = (ov[1] - ov[0]) if ov else 0
        t_win = T_WIN_PER_PX * tw * th
        return int(round((t_win + (th - have) * tw * 4 / BYTES_PER_S) / (0.05 * t_win)))

    holders = [[q for q in range(world) if _overlap(own[q], (w[0], w[0] + w[3]))] for w in wins]
    costs = [[cost(wi, q) for q in range(world)] for wi in range(n)]
    # start: cheapest rank each (upper bound)
    best_owner = [min(range(world), key=lambda q: (costs[wi][q], q)) for wi in range(n)]
    load0 = [0] * world
    for wi, q in enumerate(best_owner):
        load0[q] += costs[wi][q]
    best = [max(load0), sum(costs[wi][q] for wi, q in enumerate(best_owner))]
    load = [0] * world
    cur = [0] * n
    rem_min = [0] * (n + 1)          # cheapest possible cost of the windows wi..n-1
    for wi in range(n - 1, -1, -1):
        rem_min[wi] = rem_min[wi + 1] + min(costs[wi])
    budget = [200000]                # nodes; the best assignment found so far is used beyond (same on every rank)

    def dfs(wi, total, top):
        if wi == n:
            if (top, total) < (best[0], best[1]):
                best[0], best[1] = top, total
                best_owner[:] = cur
            return
        cands = list(holders[wi])
        others = [q for q in range(world) if q not in holders[wi]]
        if others:
            cands.append(min(others, key=lambda q: (load[q], q)))
        for q in sorted(cands, key=lambda r: (costs[wi][r], r)):
            c = costs[wi][q]
            ntop = max(top, load[q] + c)
            if (ntop, total + c + rem_min[wi + 1]) >= (best[0], best[1]) or budget[0] <= 0:
                continue
            budget[0] -= 1
            load[q] += c
            cur[wi] = q
            dfs(wi + 1, total + c, ntop)
            load[q] -= c

    if n <= 32:
        dfs(0, 0, 0)
    _OWNER_MEMO[key] = list(best_owner)
    return list(best_owner)


def sharded_topousm_scale(band: torch.Tensor, H: int, rank: int, world: int, *, radii, weights, pixel_size=1.0,
                          dist=None, grid: int = 3, block_fn=None, select_fns=None, spec: Optional[Speculation] = None,
                          scale_out: Optional[torch.Tensor] = None, after_gather=None,
                          peers: Optional["PeerBands"] = None):
    """p99(|raw topousm_fast|) over the reference's stratified full-resolution windows.  Each window is
    evaluated whole by ONE rank (assign_window_owners) after gathering its rows from the owning bands; the
    percentile over all windows is an exact distributed selection.
    spec + scale_out (product path): no host synchronisation -- the bounding box is a guess checked later
    (Speculation) and the scale stays on the device (scale_out, NaN = no scale); returns scale_out.
    after_gather: called once the window gather is enqueued (the step starts the main-pass preparation there: the
    communication library runs its operations in issue order, so the bounding-box all-reduce and the window gather
    go first and the halo exchange, which has to wait for the pyramid kernel anyway, last).
    peers (PeerBands): the bands live in symmetric memory -- a window owner copies the rows it lacks straight out of
    the other ranks' bands (strided loads over NVLink, ~480 GB/s per reader, no packing, no send / receive pairing)
    after one stream-ordered barrier; the caller issues the closing barrier before any band may change."""
    from ..algorithms._norm_stats import _norm_stat_window_geometry, stratified_windows
    W = int(band.shape[1])
    own = band_bounds(H, world)
    r0, r1 = own[rank]
    margin, tile = _norm_stat_window_geometry("topousm_fast", {"radii": list(radii)})

    def _mark(name):
        if PREPASS_TRACE is not None and band.is_cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(band.device))
            PREPASS_TRACE.append((name, ev))
    _mark("start")
    # valid-data bounding box from a <=512 px nearest overview of the own rows
    cov = max(1, max(W, H) // 512)
    big = 1 << 40
    first = ((r0 + cov - 1) // cov) * cov
    if spec is not None:
        from .. import kernels as k
        n_rows = 0
        if r1 > first and first // cov < max(1, H // cov):
            n_rows = max(0, min((r1 - first + cov - 1) // cov, max(1, H // cov) - first // cov))
        n_cols = min((W + cov - 1) // cov, max(1, W // cov))
        boxd = k.valid_bbox(band, first - r0, cov, n_rows, n_cols, first // cov)
        box_ready = torch.cuda.Event()
        box_ready.record(torch.cuda.current_stream(band.device))
        full = [0, max(1, H // cov) - 1, 0, n_cols - 1]
        bkey = ("bbox", H, W, world)
        g = spec.peek(bkey, [-full[0], full[1], -full[2], full[3]])
        ymin, ymax, xmin, xmax = -g[0], g[1], -g[2], g[3]
        box = None
        bbox_pending = [True]

        def check_bbox():
            # the all-reduced box only CONFIRMS the guess the step is planned with, so its all-reduce is kept off the
            # critical path: it is issued behind the window gather (every rank at the same point of its program), on
            # the side stream of the main-pass preparation when there is one
            if bbox_pending[0]:
                bbox_pending[0] = False
                if dist is not None and world > 1:
                    dist.all_reduce(boxd, op=dist.ReduceOp.MAX)
                spec.verify(bkey, boxd, g)
    else:
        box = torch.tensor([big, -1, big, -1], dtype=torch.int64, device=band.device)  # ymin, ymax, xmin, xmax
    if box is not None and r1 > first and first // cov < max(1, H // cov):
        ov = band[first - r0::cov, ::cov][:, : max(1, W // cov)]
        ov = ov[: max(0, min(ov.shape[0], max(1, H // cov) - first // cov))]
        ok = torch.isfinite(ov)
        if bool(ok.any()):
            rows = torch.nonzero(ok.any(dim=1)).flatten() + first // cov
            cols = torch.nonzero(ok.any(dim=0)).flatten()
            box = torch.stack([rows.min(), rows.max(), cols.min(), cols.max()]).to(torch.int64)
    if box is not None:
        if dist is not None and world > 1:
            mn = box[[0, 2]].clone(); mx = box[[1, 3]].clone()
            dist.all_reduce(mn, op=dist.ReduceOp.MIN)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            box = torch.stack([mn[0], mx[0], mn[1], mx[1]])
        ymin, ymax, xmin, xmax = [int(v) for v in box.cpu().tolist()]
    if ymax < 0:
        if spec is not None:
            check_bbox()
        if after_gather is not None:
            after_gather()
        if scale_out is not None:
            scale_out.fill_(float("nan"))
            return scale_out
        return None
    by0, by1 = ymin * cov, min(H, (ymax + 1) * cov)
    bx0, bx1 = xmin * cov, min(W, (xmax + 1) * cov)
    wins = stratified_windows(W, H, by0, by1, bx0, bx1, grid=grid, tile=min(tile, max(W, H)))
    trimmed_fn = None
    if block_fn is None:
        from .. import kernels as k

        def trimmed_fn(a, m):   # the kernel is asked only for the region the trim keeps (same values)
            raw = k.topousm_fast(a, radii=radii, weights=weights, pixel_size=pixel_size,
                                 roi=(m, int(a.shape[0]) - 2 * m, m, int(a.shape[1]) - 2 * m) if m > 0 else None)
            return raw[m:-m, m:-m] if m > 0 else raw
    _mark("bbox")
    # 1. move every window's rows to its owner ...
    owners = assign_window_owners(wins, own)
    if peers is not None:
        from .. import kernels as k
        peers.barrier()          # every rank's band is in place
        cur_s = torch.cuda.current_stream(band.device)
        ce_s = _side_stream(band.device, "gather")
        ce_s.wait_stream(cur_s)
        mine, used_ce = [], False
        for wi, (wy0, wx0, tw, th) in enumerate(wins):
            if owners[wi] != rank:
                continue
            if r0 <= wy0 and wy0 + th <= r1:
                mine.append(band[wy0 - r0:wy0 - r0 + th, wx0:wx0 + tw])      # a view, nothing moves
                continue
            win = torch.empty((th, tw), dtype=band.dtype, device=band.device)
            for q in range(world):
                ov = _overlap(own[q], (wy0, wy0 + th))
                if not ov:
                    continue
                dst = win[ov[0] - wy0:ov[1] - wy0]
                if q == rank:    # own rows: a copy kernel on this stream ...
                    dst.copy_(band[ov[0] - r0:ov[1] - r0, wx0:wx0 + tw])
                    continue
                # ... rows of another band: the copy engines pull them over NVLink meanwhile
                with torch.cuda.stream(ce_s):
                    k.copy_rect(dst, peers.band_of(q)[ov[0] - own[q][0]:ov[1] - own[q][0], wx0:wx0 + tw])
                win.record_stream(ce_s)
                used_ce = True
                if EXCHANGE_LOG is not None:
                    EXCHANGE_LOG["received"] += (ov[1] - ov[0]) * tw * 4
                    EXCHANGE_LOG["peer_read"] = EXCHANGE_LOG.get("peer_read", 0) + (ov[1] - ov[0]) * tw * 4
            mine.append(win)
        if used_ce:
            cur_s.wait_stream(ce_s)
    else:
        plans, mine_idx = [], []
        for wi, (wy0, wx0, tw, th) in enumerate(wins):
            owner = owners[wi]
            need = [(0, 0)] * world
            need[owner] = (wy0, wy0 + th)
            cols = band[:, wx0:wx0 + tw]
            plans.append(plan_exchange(cols, own, need, rank, dist))
            if rank == owner:
                mine_idx.append(wi)
        outs = run_exchanges(plans, dist)          # all nine gathers in one point-to-point batch
        mine = [outs[wi] for wi in mine_idx]
    _mark("gather")
    if spec is not None:
        if after_gather is not None and getattr(after_gather, "side_stream", None) is not None:
            side_s = after_gather.side_stream
            side_s.wait_event(box_ready)   # (the box kernel ran on the caller's stream)
            with torch.cuda.stream(side_s):
                boxd.record_stream(side_s)
                check_bbox()
        else:
            check_bbox()
    if after_gather is not None:
        after_gather()
    # 2. ... then every rank evaluates its own windows, all ranks at the same time
    def job_for(win):
        def job():
            m = int(min(margin, win.shape[0] // 3, win.shape[1] // 3))
            if trimmed_fn is not None:
                return trimmed_fn(win, m)
            raw = block_fn(win)
            return raw[m:-m, m:-m] if m > 0 else raw
        return job

    from .. import _device as _dev
    pooled = [r for r in _dev.run_concurrently([job_for(w) for w in mine], band.device) if r.numel()]
    _mark("windows")
    kw = select_fns(pooled) if select_fns is not None else {}   # tests inject stand-ins for the kernels
    if scale_out is not None:
        px = None
        if peers is not None and dist is not None and world > 1:   # (every rank: same decision, same rendezvous order)
            px = PeerExchange.get(band.device, dist.group.WORLD, world)
        return distributed_percentile(pooled, 99.0, take_abs=True, finite_only=False, device=band.device, dist=dist,
                                      scale_out=scale_out, peer_exchange=px)
    s = distributed_percentile(pooled, 99.0, take_abs=True, finite_only=False, device=band.device, dist=dist, **kw)
    if not (s == s) or s <= 1e-9:
        return None
    return s


# ------------------------------------------------------------------------------------------------
# bench driver for N > 1 (called from bench.py under torchrun)
# ------------------------------------------------------------------------------------------------
def bind_host_to_gpu(device) -> dict:
    """Pin this process to the CPUs NVML reports as local to `device` (same NUMA node / PCIe root) BEFORE the pinned
    staging buffers are allocated, so that their pages are local and the upload / download DMA does not cross the
    socket interconnect (one process per GPU: eight ranks share the host's memory system).  Best effort: returns what
    was done, never raises."""
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device)
        bus = "%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        try:
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        local = {i for i in range(ncpu) if (int(mask[i // 64]) >> (i % 64)) & 1}
        allowed = set(os.sched_getaffinity(0))
        cpus = sorted(local & allowed)
        info.update(local_cpus=len(local), allowed_cpus=len(allowed), used_cpus=len(cpus))
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
    except Exception as e:   # NVML missing, cpuset restrictions, ...: stay where we are
        info["error"] = repr(e)
    return info


def _breakdown(prof, steps):
    names = {1: "fused_main", 2: "pyramid", 3: "coarse_means", 6: "fused_stats_windows"}
    out = {}
    for tag, ms in prof:
        if tag in names:
            out[names[tag]] = out.get(names[tag], 0.0) + ms / max(1, steps)
    return {k_: round(v, 3) for k_, v in out.items()}


def bench_sharded(a, dist, dev, metric, unit, radii, weights, clock_sampler=None, peak=None):
    import json
    import time
    from .. import kernels as k
    world, rank = dist.get_world_size(), dist.get_rank()
    S = int(a.size)
    H = W = S
    own = band_bounds(H, world)
    r0, r1 = own[rank]
    # the band lives inside its halo buffer, in symmetric memory (window gather by peer reads; FSG_NO_PEER=1: NCCL)
    pg = None if os.environ.get("FSG_NO_PEER") else dist.group.WORLD
    ext, band = haloed_band(H, W, world, rank, radii, device=dev, peer_group=pg)
    k.synth_dem((r1 - r0, W), seed=20261017 + 2, device=dev, row0=r0, h_global=H, out=band)
    out = torch.empty((r1 - r0, W), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()

    redo = [0]
    host_t = [0.0, 0.0]   # seconds spent enqueueing / waiting for the speculated values

    def step():   # no host synchronisation inside (see topousm_fast_sharded_step); a wrong guess repeats the step
        for _attempt in range(3):
            t_a = time.perf_counter()
            _res, scale_dev, spec = topousm_fast_sharded_step(band, H, rank, world, radii=radii, weights=weights, dist=dist,
                                                              out=out, dem_ext=ext)
            t_b = time.perf_counter()
            good = spec.ok()
            host_t[0] += t_b - t_a
            host_t[1] += time.perf_counter() - t_b
            if good:
                break
            redo[0] += 1
        return scale_dev

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    redo[0] = 0
    k.reset_launch_count()
    sampler = clock_sampler(dev.index or 0) if (clock_sampler is not None and rank == 0) else None
    if sampler is not None:
        sampler.start()
    k.profile_enable(True)
    import os as _os
    global STEP_TRACE
    global PREPASS_TRACE
    if _os.environ.get("FSG_STEP_TRACE"):
        STEP_TRACE = []
        PREPASS_TRACE = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    host_t[0] = host_t[1] = 0.0
    for _ in range(a.steps):
        scale_dev = step()
    ev1.record()
    host_ms = {"enqueue": round(host_t[0] * 1e3 / a.steps, 3), "wait_speculation": round(host_t[1] * 1e3 / a.steps, 3)}
    torch.cuda.synchronize()
    if STEP_TRACE is not None:
        import sys as _sys
        print(f"[rank {rank}] step trace: {step_trace_summary()}", file=_sys.stderr)
        acc = {}
        tr = PREPASS_TRACE
        for i in range(1, len(tr)):
            if tr[i][0] != "start":
                acc.setdefault(tr[i][0], []).append(tr[i - 1][1].elapsed_time(tr[i][1]))
        print(f"[rank {rank}] pre-pass phases (ms): " + ", ".join(f"{k_} {sum(v) / len(v):.2f}" for k_, v in acc.items()), file=_sys.stderr)
        STEP_TRACE = None
        PREPASS_TRACE = None
    scale = float(scale_dev.item())
    csum = torch.zeros(1, dtype=torch.int64, device=dev)
    for r in range(0, r1 - r0, 8192):
        csum += out[r:r + 8192].view(torch.int32).sum(dtype=torch.int64)
    dist.all_reduce(csum, op=dist.ReduceOp.SUM)
    checksum = int(csum.item()) & ((1 << 64) - 1)
    prof = k.profile_read(max_records=4096)
    k.profile_enable(False)
    clocks = sampler.stop() if sampler is not None else None
    dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = torch.tensor([k.launch_count()], dtype=torch.int64, device=dev)
    dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    ms_step = float(ms.item()) / a.steps
    # main pass only
    torch.cuda.synchronize(); dist.barrier()
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m0.record()
    for _ in range(a.steps):
        topousm_fast_sharded(band, H, rank, world, radii=radii, weights=weights, norm_scale=scale, dist=dist, out=out,
                             dem_ext=ext)
    m1.record()
    torch.cuda.synchronize()
    mm = torch.tensor([m0.elapsed_time(m1)], dtype=torch.float64, device=dev)
    dist.all_reduce(mm, op=dist.ReduceOp.MAX)
    main_ms = float(mm.item()) / a.steps
    # e2e: pinned host band -> device -> uint8 band back to pinned host, every step
    from ..io.output_encoding import quantize_params, resolve_output_range
    qp = quantize_params(*resolve_output_range("topousm_fast"), "uint8")
    binding = None if os.environ.get("FSG_NO_NUMA_BIND") else bind_host_to_gpu(dev)
    hin = torch.empty((r1 - r0, W), dtype=torch.float32, pin_memory=True)
    hin.copy_(band)
    hout = torch.empty((r1 - r0, W), dtype=torch.uint8, pin_memory=True)
    dext, dband = haloed_band(H, W, world, rank, radii, device=dev, peer_group=pg)
    out8 = torch.empty((r1 - r0, W), dtype=torch.uint8, device=dev)

    def e2e_step():
        dband.copy_(hin, non_blocking=True)
        topousm_fast_sharded_with_stats(dband, H, rank, world, radii=radii, weights=weights, dist=dist,
                                        output_dtype="uint8", qp=qp, out=out8, dem_ext=dext)
        hout.copy_(out8, non_blocking=True)
        torch.cuda.synchronize()

    e2e_step()
    dist.barrier()
    t0 = time.perf_counter()
    n_e2e = max(1, min(a.steps, 3))
    for _ in range(n_e2e):
        e2e_step()
    dist.barrier()
    dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        px = H * W
        # roofline of the dominant kernel on rank 0: the fused full-resolution kernel over this rank's band
        roofline = None
        main_fused = [ms_ for tag, ms_ in prof if tag == 1]
        if main_fused and peak is not None:
            if len(main_fused) != a.steps + redo[0]:
                raise RuntimeError(f"profiler recorded {len(main_fused)} main fused passes for {a.steps} steps")
            fms = sum(main_fused) / len(main_fused)
            ach = 8.0 * (r1 - r0) * W / (fms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "achieved": ach, "peak": peak[0], "unit": "GB/s", "frac": ach / peak[0],
                        "traffic": None, "kernel": "fsg::fused_kernel_v8<3> + v6 borders (rank 0 band)", "peak_source": peak[1],
                        "algorithmic_bytes_per_px": 8.0, "fused_kernel_ms": fms}
        line = {
            "metric": metric, "value": px / (ms_step * 1e-3) / 1e6, "unit": unit, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 (f64 window sums)", "data": "synthetic",
            "config": {"workload": f"topousm_fast --mode spatial, radii 2,8,32,128,512,2048, 2^n weights, {S}x{S} f32 DEM "
                                   f"row-band sharded over {world} GPUs (halo rows exchanged point-to-point over NVLink), "
                                   "stats pre-pass + main pass per step", "radii": radii, "size": S,
                       "l2_policy": "per-GPU band far larger than the 126 MB L2", "main_pass_ms": main_ms,
                       "stats_prepass_ms": ms_step - main_ms, "main_pass_mpx_s": px / (main_ms * 1e-3) / 1e6,
                       "kernel_ms_per_step_rank0": _breakdown(prof, a.steps), "host_enqueue_ms_per_step": host_ms,
                       "scale_p99": scale, "band_rows": r1 - r0, "respeculated_steps": redo[0],
                       "out_checksum": f"{checksum:016x}"},
            "roofline": roofline, "cpu_baseline": None, "clocks": clocks,
            "e2e": {"value": px / float(dt.item()) / 1e6, "unit": unit, "h2d_bytes_per_step": px * 4,
                    "d2h_bytes_per_step": px, "ms_per_step": float(dt.item()) * 1e3, "output_dtype": "uint8",
                    "host_binding_rank0": binding},
            "gpu_launches": int(launches.item()),
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
