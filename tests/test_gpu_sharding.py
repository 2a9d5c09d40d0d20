"""GPU tests of the row-band shard stages (C band API) and, when >= 2 GPUs are visible, of the real
NCCL pipeline: bands must reproduce the whole-raster kernel bit for bit."""
import os
import sys
import tempfile

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import terrain_oracle as orc  # noqa: E402


class _SlicedExchange:
    """Single-process stand-in for the halo exchange: every virtual rank sees the whole arrays, so
    exchange_rows() is replaced by slicing (the orchestration itself is covered by the gloo tests)."""


def _virtual_bands(dem_t, world, radii, weights, norm_scale=None, output_dtype="float32", qp=None):
    from fujishadergpu_b200 import kernels as k
    from fujishadergpu_b200.core import sharding as sh
    H, W = dem_t.shape
    own = sh.band_bounds(H, world)
    plan = k.topousm_plan(radii, 1.0)
    kinds, factors, sizes, R = plan["kind"], plan["factor"], plan["size"], plan["fused_halo"]
    levels = sorted({factors[i] for i in range(len(radii)) if kinds[i] == 1})
    # per-band pyramids stitched into full level grids (what the exchange would move around)
    full = {}
    for f in levels:
        parts = [k.pyramid_band(dem_t[a:b], [f])[0][0] for (a, b) in own if b > a]
        full[f] = torch.cat(parts, dim=0)
        assert full[f].shape[0] == (H + f - 1) // f
    outs = []
    for (r0, r1) in own:
        if r1 <= r0:
            continue
        halo = max([R] + [4 if sizes[i] == 0 else sizes[i] // 2 for i in range(len(radii)) if kinds[i] == 2])
        lo, hi = sh.mirror_need(r0 - halo, r1 - 1 + halo, H, True)
        dem_ext = dem_t[lo:hi]
        grids, grow0 = [None] * len(radii), [0] * len(radii)
        for i in range(len(radii)):
            if kinds[i] == 1:
                f = factors[i]
                gh = (H + f - 1) // f
                rs = (gh - 1) / (H - 1) if H > 1 else 1.0
                mlo = int(np.floor(r0 * rs)); mhi = min(gh - 1, int(np.floor((r1 - 1) * rs)) + 1)
                reach = 4 if sizes[i] == 0 else sizes[i] // 2
                glo, ghi = sh.mirror_need(mlo - reach, mhi + reach, gh, True)
                grids[i] = k.grid_mean_band(full[f][glo:ghi], glo, gh, sizes[i], mlo, mhi - mlo + 1)
                grow0[i] = mlo
            elif kinds[i] == 2:
                grids[i] = k.grid_mean_band(dem_ext, lo, H, sizes[i], r0, r1 - r0)
                grow0[i] = r0
        outs.append(k.topousm_fused_band(dem_ext, lo, H, r0, r1 - r0, radii=radii, weights=weights, term_grids=grids,
                                         term_grow0=grow0, norm_scale=norm_scale, output_dtype=output_dtype, qp=qp))
    return torch.cat(outs, dim=0)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_band_stages_equal_whole_raster_bitwise(world):
    from fujishadergpu_b200 import kernels as k
    for nodata, radii in ((False, [2, 8, 32, 128, 512, 2048]), (False, [1, 41, 42, 100]), (True, [2, 8, 32, 128])):
        dem = orc.synth_dem(1200, 900, seed=90 + world, nodata=nodata)
        if nodata:
            # keep every coarse cell partially valid (the void-fill gather path is covered on CPU)
            dem = np.where(np.isnan(dem) & (np.add.outer(np.arange(1200), np.arange(900)) % 3 == 0), 400.0, dem).astype(np.float32)
        d = torch.from_numpy(dem).cuda()
        w = orc.pow2_weights(len(radii))
        whole = k.topousm_fast(d, radii=radii, weights=w, norm_scale=9.5)
        bands = _virtual_bands(d, world, radii, w, norm_scale=9.5)
        assert torch.equal(torch.nan_to_num(whole, nan=-7777.0), torch.nan_to_num(bands, nan=-7777.0)), (world, radii)


def test_single_rank_orchestration_equals_whole_raster():
    """world = 1 through the sharded orchestration (what bench.py times at N = 1: main-pass preparation on a side
    stream underneath the statistics pre-pass, device-gated void fill) == the sequential whole-raster calls,
    bit for bit, dense and with all-NaN coarse cells."""
    from fujishadergpu_b200 import kernels as k
    from fujishadergpu_b200.core import sharding as sh
    from fujishadergpu_b200.algorithms._norm_stats import compute_norm_stats_device
    radii, w = [2, 8, 32, 128, 512, 2048], orc.pow2_weights(6)
    for nodata in (False, True):
        dem = orc.synth_dem(2100, 1900, seed=33, nodata=nodata)
        if nodata:
            dem[640:800, 512:700] = np.nan        # whole 16 x 16 and 4 x 4 cells without data -> enclosed-void fill
        d = torch.from_numpy(dem).cuda()
        st = compute_norm_stats_device(d, "topousm_fast", {"radii": radii, "weights": w, "pixel_size": 1.0})
        want = k.topousm_fast(d, radii=radii, weights=w, norm_scale=st[0])
        for _ in range(2):   # second call: warm side stream, same result
            got, scale = sh.topousm_fast_sharded_with_stats(d, 2100, 0, 1, radii=radii, weights=w, dem_ext=d)
            assert scale == st[0]
            assert torch.equal(torch.nan_to_num(got, nan=-7777.0), torch.nan_to_num(want, nan=-7777.0)), nodata


def _nccl_worker(rank, world, port, path, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from fujishadergpu_b200.core import sharding as sh
    dem = np.load(path)
    H = dem.shape[0]
    r0, r1 = sh.band_bounds(H, world)[rank]
    band = torch.from_numpy(np.ascontiguousarray(dem[r0:r1])).cuda()
    radii, w = [2, 8, 32, 128, 512, 2048], orc.pow2_weights(6)
    scale = sh.sharded_topousm_scale(band, H, rank, world, radii=radii, weights=w, dist=dist)
    out = sh.topousm_fast_sharded(band, H, rank, world, radii=radii, weights=w, norm_scale=scale, dist=dist)
    # the same band kept inside its halo buffer (no copy of the own rows): identical result
    ext, view = sh.haloed_band(H, dem.shape[1], world, rank, radii, device=band.device)
    view.copy_(band)
    out2 = sh.topousm_fast_sharded(view, H, rank, world, radii=radii, weights=w, norm_scale=scale, dist=dist, dem_ext=ext)
    assert torch.equal(torch.nan_to_num(out, nan=-7777.0), torch.nan_to_num(out2, nan=-7777.0))
    # one-call form: the scale-independent part of the main pass overlaps the statistics pre-pass on a side stream
    out3, scale3 = sh.topousm_fast_sharded_with_stats(view, H, rank, world, radii=radii, weights=w, dist=dist, dem_ext=ext)
    assert scale3 == scale
    assert torch.equal(torch.nan_to_num(out, nan=-7777.0), torch.nan_to_num(out3, nan=-7777.0))
    # bands in symmetric memory: the window gather reads the other ranks' bands directly over NVLink (PeerBands)
    pext, pview = sh.haloed_band(H, dem.shape[1], world, rank, radii, device=band.device, peer_group=dist.group.WORLD)
    # (a node without symmetric-memory support falls back to send / receive inside haloed_band: same call, same result;
    #  the marker file tells the parent which path ran)
    if sh.PeerBands.of(pext) is not None and rank == 0:
        open(os.path.join(out_dir, "peer_path_ran"), "w").close()
    pview.copy_(band)
    for _ in range(2):   # twice: the closing barrier of a step lets the next one refill / reread the bands
        out4, scale4 = sh.topousm_fast_sharded_with_stats(pview, H, rank, world, radii=radii, weights=w, dist=dist, dem_ext=pext)
        assert scale4 == scale
        assert torch.equal(torch.nan_to_num(out, nan=-7777.0), torch.nan_to_num(out4, nan=-7777.0))
    np.save(os.path.join(out_dir, f"out_{rank}.npy"), out.cpu().numpy())
    if rank == 0:
        np.save(os.path.join(out_dir, "scale.npy"), np.array([scale]))
    dist.barrier()
    dist.destroy_process_group()


def test_nccl_pipeline_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    from fujishadergpu_b200 import kernels as k
    from fujishadergpu_b200.algorithms._norm_stats import compute_norm_stats_device
    world = min(torch.cuda.device_count(), 4)
    dem = orc.synth_dem(4128, 3000, seed=5, nodata=True)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "dem.npy")
        np.save(path, dem)
        mp.spawn(_nccl_worker, args=(world, 29650, path, td), nprocs=world, join=True)
        got = np.concatenate([np.load(os.path.join(td, f"out_{r}.npy")) for r in range(world)], axis=0)
        scale = float(np.load(os.path.join(td, "scale.npy"))[0])
        if not os.path.exists(os.path.join(td, "peer_path_ran")):
            import warnings
            warnings.warn("symmetric memory unavailable on this node: the peer-memory gather / exchange was not exercised")
    d = torch.from_numpy(dem).cuda()
    radii, w = [2, 8, 32, 128, 512, 2048], orc.pow2_weights(6)
    st = compute_norm_stats_device(d, "topousm_fast", {"radii": radii, "weights": w, "pixel_size": 1.0})
    assert scale == st[0]
    want = k.topousm_fast(d, radii=radii, weights=w, norm_scale=st[0]).cpu().numpy()
    assert np.array_equal(got, want, equal_nan=True)
