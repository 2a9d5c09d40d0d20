"""Multi-process (gloo, CPU) tests of the row-band sharding host logic.

The compute stages are replaced by a NumPy stand-in built from the oracle (global-coordinate band
versions of the pyramid, the coarse means and the fused pass), so what is under test is exactly the
product's orchestration in fujishadergpu_b200/core/sharding.py: band geometry, which halo rows are
exchanged, global decimation anchors / zoom mapping / edge rules, the void-fill fallback and the
distributed exact percentile.  Expected result: the assembled bands equal the oracle on the whole
raster bit for bit.
"""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import terrain_oracle as orc  # noqa: E402
from fujishadergpu_b200.core import sharding as sh  # noqa: E402


def _gather_rows(src, src_row0, lo, hi, n, reflect):
    """rows [lo, hi] (inclusive, may lie outside [0, n)) of the global array, via mirror / clamp."""
    idx = np.arange(lo, hi + 1)
    if reflect:
        p = 2 * n
        idx = np.mod(idx, p)
        idx = np.where(idx < n, idx, p - 1 - idx)
    else:
        idx = np.clip(idx, 0, n - 1)
    return src[idx - src_row0]


class NumpyBackend:
    """Oracle-based stand-in for CudaBackend (same interface, torch CPU tensors in/out)."""

    def plan(self, radii, pixel_size):
        kind, factor, size, halo, nf = [], [], [], 0, 0
        for (r, ds, k, param) in orc.topousm_plan(radii, pixel_size):
            if ds > 1:
                kind.append(1); factor.append(ds); size.append(0 if k == "gauss1" else int(param))
            elif r <= 1:
                kind.append(2); factor.append(1); size.append(0)
            elif r <= 41 and nf < 8:
                kind.append(0); factor.append(1); size.append(2 * r + 1); nf += 1; halo = max(halo, r)
            else:
                kind.append(2); factor.append(1); size.append(2 * r + 1)
        return {"kind": kind, "factor": factor, "size": size, "fused_halo": halo}

    def pyramid(self, band, factors):
        a = band.numpy()
        grids = [orc.decimate_cells(a, f) for f in factors]
        flags = torch.tensor([int(np.isnan(g).any()) for g in grids], dtype=torch.int32)
        return [torch.from_numpy(g) for g in grids], flags

    def grid_mean(self, src, src_row0, gh, size, out_row0, out_rows):
        s = src.numpy()
        reach = 4 if size == 0 else size // 2
        ext = _gather_rows(s, src_row0, out_row0 - reach, out_row0 + out_rows - 1 + reach, gh, reflect=size > 0)
        m = orc.gauss_mean(ext, 1.0, "nearest") if size == 0 else orc.box_mean(ext, size, "reflect")
        return torch.from_numpy(np.ascontiguousarray(m[reach:reach + out_rows]))

    def void_fill(self, grid):
        return torch.from_numpy(orc.fill_enclosed_voids(grid.numpy().copy()))

    def fused(self, dem_ext, dem_row0, H, out_row0, out_rows, *, radii, weights=None, pixel_size=1.0,
              term_grids=None, term_grow0=None, norm_scale=None, output_dtype="float32", qp=None, out=None):
        d = dem_ext.numpy()
        W = d.shape[1]
        x = d[out_row0 - dem_row0:out_row0 - dem_row0 + out_rows]
        plan = self.plan(radii, pixel_size)
        wv = (np.array([1.0 / len(radii)] * len(radii), np.float32) if weights is None
              else np.asarray(weights, np.float32))
        acc = None
        for i, r in enumerate(radii):
            if plan["kind"][i] == 0:
                ext = _gather_rows(d, dem_row0, out_row0 - r, out_row0 + out_rows - 1 + r, H, reflect=True)
                mean = orc.box_mean(ext, 2 * r + 1, "reflect")[r:r + out_rows]
            elif plan["kind"][i] == 2:
                mean = term_grids[i].numpy()[out_row0 - term_grow0[i]:out_row0 - term_grow0[i] + out_rows]
            else:
                g = term_grids[i].numpy().astype(np.float64)
                f = plan["factor"][i]
                gh, gw = (H + f - 1) // f, (W + f - 1) // f
                ri = np.arange(out_row0, out_row0 + out_rows, dtype=np.float64) * ((gh - 1) / (H - 1) if H > 1 else 1.0)
                ci = np.arange(W, dtype=np.float64) * ((gw - 1) / (W - 1) if W > 1 else 1.0)
                r0 = np.minimum(np.floor(ri).astype(np.int64), gh - 1)
                c0 = np.minimum(np.floor(ci).astype(np.int64), gw - 1)
                r1, c1 = np.minimum(r0 + 1, gh - 1), np.minimum(c0 + 1, gw - 1)
                fr, fc = (ri - r0)[:, None], (ci - c0)[None, :]
                g0, g1 = g[r0 - term_grow0[i]], g[r1 - term_grow0[i]]
                t = g0[:, c0] * (1 - fr) * (1 - fc)
                t = t + g0[:, c1] * (1 - fr) * fc
                t = t + g1[:, c0] * fr * (1 - fc)
                t = t + g1[:, c1] * fr * fc
                mean = t.astype(np.float32)
            term = wv[i] * (x - mean)
            acc = term if acc is None else acc + term
        acc[np.isnan(x)] = np.nan
        if norm_scale is not None:
            acc = orc.normalise_by_scale(acc, (norm_scale,))
        res = torch.from_numpy(np.ascontiguousarray(acc.astype(np.float32)))
        if out is not None:
            out.copy_(res)
            return out
        return res


def _numpy_select_fns(pooled):
    """NumPy stand-ins for key_histogram / key_rank_info / key_to_float (host-logic tests on CPU)."""
    vals = np.concatenate([np.abs(p.cpu().numpy().ravel()) for p in pooled]) if pooled else np.zeros(0, np.float32)
    vals = vals[~np.isnan(vals)].astype(np.float32)
    keys = vals.view(np.uint32).astype(np.int64)

    def hist_fn(level, prefix, mask):
        shift, bins = ((21, 2048), (10, 2048), (0, 1024))[level]
        sel = keys[(keys & mask) == prefix]
        h = np.bincount((sel >> shift) & (bins - 1), minlength=2048).astype(np.int64)
        return torch.from_numpy(h), torch.tensor([keys.size], dtype=torch.int64)

    def rank_info_fn(key):
        le = int((keys <= key).sum())
        gt = keys[keys > key]
        return torch.tensor([le, int(gt.min()) if gt.size else 0xffffffff], dtype=torch.int64)

    def key_to_float(key):
        return float(np.array([key], dtype=np.uint32).view(np.float32)[0])

    return dict(hist_fn=hist_fn, rank_info_fn=rank_info_fn, key_to_float=key_to_float)


def _worker(rank, world, port, dem_path, out_dir, radii, weights, with_stats, haloed=False):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dem = np.load(dem_path)
    H = dem.shape[0]
    r0, r1 = sh.band_bounds(H, world)[rank]
    band = torch.from_numpy(np.ascontiguousarray(dem[r0:r1]))
    be = NumpyBackend()
    ext = None
    if haloed:   # the band lives inside its halo buffer: only the halo rows are exchanged, nothing is copied
        ext, view = sh.haloed_band(H, dem.shape[1], world, rank, radii, device="cpu", backend=be)
        ext.fill_(float("nan"))
        view.copy_(band)
        band = view
    scale = None
    if with_stats:
        scale = sh.sharded_topousm_scale(band, H, rank, world, radii=radii, weights=weights, dist=dist,
                                         block_fn=lambda a: torch.from_numpy(orc.topousm_fast_block(a.numpy(), radii=radii, weights=weights)),
                                         select_fns=_numpy_select_fns)
    out = sh.topousm_fast_sharded(band, H, rank, world, radii=radii, weights=weights, norm_scale=scale, dist=dist, backend=be,
                                  dem_ext=ext)
    if with_stats:   # the one-call form bench.py times must give the same scale and rows
        out1, scale1 = sh.topousm_fast_sharded_with_stats(
            band, H, rank, world, radii=radii, weights=weights, dist=dist, dem_ext=ext, backend=be,
            block_fn=lambda a: torch.from_numpy(orc.topousm_fast_block(a.numpy(), radii=radii, weights=weights)),
            select_fns=_numpy_select_fns)
        assert scale1 == scale and torch.equal(torch.nan_to_num(out1, nan=-7.0), torch.nan_to_num(out, nan=-7.0))
    np.save(os.path.join(out_dir, f"out_{rank}.npy"), out.numpy())
    if rank == 0 and with_stats:
        np.save(os.path.join(out_dir, "scale.npy"), np.array([scale if scale is not None else np.nan]))
    dist.barrier()
    dist.destroy_process_group()


def _run(world, dem, radii, weights, with_stats=False, haloed=False):
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "dem.npy")
        np.save(path, dem)
        port = 29500 + (os.getpid() * 7 + world * 13 + dem.shape[0]) % 2000
        mp.spawn(_worker, args=(world, port, path, td, radii, weights, with_stats, haloed), nprocs=world, join=True)
        outs = [np.load(os.path.join(td, f"out_{r}.npy")) for r in range(world)]
        scale = float(np.load(os.path.join(td, "scale.npy"))[0]) if with_stats else None
    return np.concatenate(outs, axis=0), scale


def test_band_bounds_are_aligned_and_cover():
    for H, world in ((65536, 8), (1000, 3), (40, 8), (131072, 4), (17, 2)):
        b = sh.band_bounds(H, world)
        assert b[0][0] == 0 and b[-1][1] == H
        for (a0, a1), (b0, b1) in zip(b, b[1:]):
            assert a1 == b0
        for (r0, r1) in b:
            if r1 > r0:   # small rasters leave trailing ranks without rows
                assert r0 % 16 == 0 and (r1 % 16 == 0 or r1 == H)


def test_window_owner_assignment_is_balanced_and_local():
    """Statistics windows: the assignment that lets the slowest rank finish first and moves the fewest rows (cost model
    of assign_window_owners); no rank gets more than ceil(n / world) windows, never more traffic than round-robin,
    every rank computes the same assignment."""
    from fujishadergpu_b200.algorithms._norm_stats import stratified_windows
    H = W = 65536
    wins = stratified_windows(W, H, 0, H, 0, W, grid=3, tile=8256)
    for world in (1, 2, 3, 4, 8, 16):
        own = sh.band_bounds(H, world)
        owners = sh.assign_window_owners(wins, own)
        assert len(owners) == len(wins) and all(0 <= o < world for o in owners)
        cap = (len(wins) + world - 1) // world
        assert max(owners.count(q) for q in range(world)) <= cap
        local = sum((sh._overlap(own[o], (w[0], w[0] + w[3])) or (0, 0))[1] - (sh._overlap(own[o], (w[0], w[0] + w[3])) or (0, 0))[0]
                    for o, w in zip(owners, wins))
        rr = sum((sh._overlap(own[i % world], (w[0], w[0] + w[3])) or (0, 0))[1] - (sh._overlap(own[i % world], (w[0], w[0] + w[3])) or (0, 0))[0]
                 for i, w in enumerate(wins))
        assert local >= rr          # never more traffic than round-robin
    assert sh.assign_window_owners(wins, sh.band_bounds(H, 8)) == sh.assign_window_owners(list(wins), sh.band_bounds(H, 8))


def test_haloed_band_layout():
    be = NumpyBackend()
    H, W, radii = 400, 64, [2, 8, 32, 128]
    for world in (2, 3):
        own = sh.band_bounds(H, world)
        for rank in range(world):
            lo, hi = sh.dem_halo_rows(H, world, rank, radii, backend=be)
            a, b = own[rank]
            assert lo == max(0, a - 32) and hi == min(H, b + 32)
            ext, band = sh.haloed_band(H, W, world, rank, radii, device="cpu", backend=be)
            assert tuple(ext.shape) == (hi - lo, W) and tuple(band.shape) == (b - a, W)
            assert band.data_ptr() == ext[a - lo:].data_ptr()


def test_mirror_need():
    assert sh.mirror_need(-3, 5, 100, True) == (0, 6)
    assert sh.mirror_need(95, 104, 100, True) == (95, 100)
    assert sh.mirror_need(-8, 2, 100, True) == (0, 8)          # mirrored rows reach further than the window
    assert sh.mirror_need(-300, 2, 100, True) == (0, 100)      # re-reflection
    assert sh.mirror_need(-3, 5, 100, False) == (0, 6)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_topousm_equals_single_block_oracle(world):
    dem = orc.synth_dem(400, 210, seed=77)
    radii, w = [2, 8, 32, 128, 512], orc.pow2_weights(5)
    got, _ = _run(world, dem, radii, w)
    want = orc.topousm_fast_block(dem, radii=radii, weights=w)
    assert got.shape == want.shape
    assert np.array_equal(got, want, equal_nan=True), float(np.nanmax(np.abs(got - want)))
    if world == 3:
        got2, _ = _run(world, dem, radii, w, haloed=True)
        assert np.array_equal(got2, want, equal_nan=True)


def test_sharded_topousm_nodata_void_fill_and_stats():
    dem = orc.synth_dem(352, 190, seed=78, nodata=True)
    dem[96:176, 64:160] = np.nan        # all-NaN coarse cells -> enclosed-void fill path
    radii, w = [1, 3, 20, 64, 200], [0.3, 0.2, 0.2, 0.2, 0.1]
    got, scale = _run(2, dem, radii, w, with_stats=True)
    raw = orc.topousm_fast_block(dem, radii=radii, weights=w)
    # single-process statistics with the reference's window geometry
    margin, tile = orc.stats_window_geometry("topousm_fast", {"radii": radii})
    pooled = []
    ok = np.isfinite(dem)
    ys, xs = np.where(ok)
    for (wy0, wx0, tw, th) in orc.stats_windows(dem.shape[1], dem.shape[0], 0, dem.shape[0], 0, dem.shape[1],
                                                grid=3, tile=min(tile, max(dem.shape))):
        r = orc.topousm_fast_block(dem[wy0:wy0 + th, wx0:wx0 + tw], radii=radii, weights=w)
        m = int(min(margin, r.shape[0] // 3, r.shape[1] // 3))
        if m > 0:
            r = r[m:-m, m:-m]
        pooled.append(r[~np.isnan(r)])
    want_scale = orc.abs_p99_scale(np.concatenate(pooled))[0]
    assert scale == want_scale
    want = orc.normalise_by_scale(raw.copy(), (want_scale,))
    assert np.array_equal(got, want, equal_nan=True)


def test_distributed_percentile_single_process_matches_numpy():
    rng = np.random.default_rng(3)
    a = (rng.standard_normal(50001) * 5).astype(np.float32)
    a[::97] = np.nan
    chunks = [torch.from_numpy(a[:20000].reshape(100, 200)), torch.from_numpy(a[20000:].reshape(1, -1))]
    fns = _numpy_select_fns(chunks)
    got = sh.distributed_percentile(chunks, 99.0, take_abs=True, finite_only=False, device="cpu", **fns)
    assert got == float(np.percentile(np.abs(a[~np.isnan(a)]), 99.0))


def test_window_owner_assignment_is_optimal_for_its_cost_model():
    """assign_window_owners (branch and bound) against brute force over all world^9 assignments for small worlds:
    same (slowest rank, rows moved) under the documented cost model (0.56 ms per 8256^2 window, 500 GB/s per reader,
    steps of 5 % of a window)."""
    import itertools
    from fujishadergpu_b200.algorithms._norm_stats import stratified_windows
    t_px, bps = 0.56e-3 / (8256.0 * 8256.0), 500e9
    for H, world in ((65536, 2), (65536, 3), (20000, 3), (65536, 4)):
        wins = stratified_windows(H, H, 0, H, 0, H, grid=3, tile=min(8256, H))
        own = sh.band_bounds(H, world)

        def cost(wi, q):
            wy0, _wx0, tw, th = wins[wi]
            ov = sh._overlap(own[q], (wy0, wy0 + th))
            have = (ov[1] - ov[0]) if ov else 0
            t_win = t_px * tw * th
            return int(round((t_win + (th - have) * tw * 4 / bps) / (0.05 * t_win)))

        costs = [[cost(wi, q) for q in range(world)] for wi in range(len(wins))]

        def score(owners):
            load = [0] * world
            for wi, q in enumerate(owners):
                load[q] += costs[wi][q]
            return max(load), sum(costs[wi][q] for wi, q in enumerate(owners))

        best = min(score(o) for o in itertools.product(range(world), repeat=len(wins)))
        sh._OWNER_MEMO.clear()
        got = score(sh.assign_window_owners(wins, own))
        assert got == best, (H, world, got, best)
