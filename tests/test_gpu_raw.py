"""The fused VPint2 load (constructor + mean init + feature plane from the RAW rasters), the non-blocking run and the
unknown-cells-only result, against (1) the host-built constructor + the round-1 device path (bit for bit), (2) NumPy's
own np.nanmean (bit for bit, any grid size, also row-sharded), (3) the oracle / golden fixtures.

Reference: VPint/VPint2.py:13-55,122-150 (constructor), VPint/MRP.py:73-77 (mean init), VPint/VPint2.py:58-119 (run)."""

import warnings

import numpy as np
import pytest

from oracle import vpint_oracle as O
from vpint_b200 import VPint2_interpolator, VPintError
from vpint_b200.engine import Solver, buffer_clouds, cloud_mask
from vpint_b200.vpint2 import solve_bands, solve_raw
from vpint_b200.wp import wp_run_options

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore")


def smooth(rng, shape, lo=300.0, hi=2700.0, modes=5):
    axes = np.meshgrid(*[np.linspace(0, 1, n) for n in shape], indexing="ij")
    acc = np.zeros(shape)
    for _ in range(modes):
        term = np.ones(shape)
        for ax in axes:
            term = term * np.sin(2 * np.pi * (rng.uniform(0.3, 2.0) * ax + rng.uniform()))
        acc += rng.uniform(0.5, 1.0) * term
    acc = (acc - acc.min()) / (acc.max() - acc.min() + 1e-12)
    return lo + (hi - lo) * acc


def discs(rng, shape, n, rmin, rmax):
    yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
    m = np.zeros(shape, dtype=bool)
    for _ in range(n):
        cy, cx, r = rng.uniform(0, shape[0]), rng.uniform(0, shape[1]), rng.uniform(rmin, rmax)
        m |= (yy - cy) ** 2 + (xx - cx) ** 2 < r * r
    return m


def raw_stack(rng, n, H, W, B, zeros=False):
    """uint16 rasters + a cloud-probability mask (0 / 100), the C3 recipe."""
    t = np.empty((n, H, W, B), dtype=np.uint16)
    f = np.empty((n, H, W, B), dtype=np.uint16)
    m = np.zeros((n, H, W), dtype=np.uint8)
    for i in range(n):
        for b in range(B):
            base = smooth(rng, (H, W))
            t[i, :, :, b] = np.round(base + rng.normal(0, 20, (H, W))).clip(1, 65535)
            f[i, :, :, b] = np.round(1.1 * base + rng.normal(0, 10, (H, W))).clip(1, 65535)
        m[i][discs(rng, (H, W), 3, min(H, W) / 16, min(H, W) / 5)] = 100
    if zeros:
        f[rng.uniform(size=f.shape) < 0.01] = 0          # zero features become 0.01 (WP_MRP.py:143-145)
    return t, f, m


def host_reference(t, f, m, opts=None, **ctor):
    """The round-1 route: constructor on the host, float32 arrays to the device, blocking solve."""
    outs, sweeps = [], []
    for i in range(t.shape[0]):
        vp = VPint2_interpolator(t[i], f[i], mask=None if m is None else m[i], **ctor)
        o, s = solve_bands(vp.target, vp.features, opts or wp_run_options(), return_sweeps=True)
        outs.append(o)
        sweeps.append(s)
    return np.stack(outs), np.stack(sweeps)


@pytest.mark.parametrize("shape", [(256, 256, 13), (96, 80, 3), (64, 200, 4), (130, 67, 2)])
def test_raw_uint16_equals_host_constructor_path(shape):
    H, W, B = shape
    rng = np.random.default_rng(2001)
    t, f, m = raw_stack(rng, 5, H, W, B, zeros=True)
    ref, ref_sweeps = host_reference(t, f, m, threshold=10)
    out, sweeps = solve_raw(t, f, m, mask_threshold=10, return_sweeps=True, chunk_patches=2, streams=2)
    assert np.array_equal(sweeps.numpy(), ref_sweeps)
    assert np.array_equal(out.numpy(), ref), "fused load must reproduce the host-built path bit for bit"
    # run_serial's contract: DTYPE (N, H, W, B)
    out32 = solve_raw(t, f, m, mask_threshold=10, out_layout="hwb", out_dtype=np.float32, chunk_patches=3, streams=3)
    assert out32.dtype.is_floating_point and out32.shape == (5, H, W, B)
    assert np.array_equal(out32.numpy(), np.moveaxis(ref, 1, -1).astype(np.float32))
    # against the oracle as well (one patch)
    vp = VPint2_interpolator(t[3], f[3], mask=m[3], threshold=10)
    oref, osw = O.vpint2_run(vp.target, vp.features)
    assert list(osw) == list(sweeps[3].numpy())
    err = np.abs(np.moveaxis(out.numpy()[3], 0, -1) - oref) / np.abs(oref)
    assert err.max() <= 1e-9


def test_raw_unknown_only_result_in_place():
    """result='unknown' writes only the cloud cells: into pinned host memory through its device mapping, or into a
    CUDA tensor; everything else in `out` is left exactly as the caller had it."""
    import torch
    rng = np.random.default_rng(2002)
    t, f, m = raw_stack(rng, 6, 64, 64, 5)
    ref, ref_sweeps = host_reference(t, f, m, threshold=10)
    ref32 = np.moveaxis(ref, 1, -1).astype(np.float32)
    cloud = np.broadcast_to((m > 10)[..., None], t.shape)
    sentinel = np.float32(-7.5)
    for where in ("pinned", "cuda"):
        out = torch.full((6, 64, 64, 5), float(sentinel), dtype=torch.float32)
        out = out.pin_memory() if where == "pinned" else out.cuda()
        tt, ff, mm = (torch.from_numpy(a).pin_memory() for a in (t, f, m))
        res, sweeps = solve_raw(tt, ff, mm, mask_threshold=10, out=out, out_layout="hwb", out_dtype=np.float32,
                                result="unknown", return_sweeps=True, chunk_patches=4, streams=2)
        got = res.cpu().numpy()
        assert np.array_equal(sweeps.numpy(), ref_sweeps)
        assert np.array_equal(got[cloud], ref32[cloud])
        assert (got[~cloud] == sentinel).all(), "known cells must not be touched"
    # float64 (N, B, H, W), the run() contract
    out = torch.full((6, 5, 64, 64), -1.0, dtype=torch.float64).pin_memory()
    solve_raw(t, f, m, mask_threshold=10, out=out, result="unknown")
    cl = np.moveaxis(cloud, -1, 1)
    assert np.array_equal(out.numpy()[cl], ref[cl]) and (out.numpy()[~cl] == -1.0).all()
    with pytest.raises(VPintError):
        solve_raw(t, f, m, out=torch.empty((6, 5, 64, 64), dtype=torch.float64), result="unknown")   # pageable host memory


def test_raw_uint16_write_back():
    """out_dtype uint16 (N4 write-back): the float64 result clipped to [0, 65535] and truncated like astype(uint16);
    full result and cloud-cells-only in place."""
    import torch
    rng = np.random.default_rng(2010)
    t, f, m = raw_stack(rng, 3, 64, 96, 4)
    ref, _ = host_reference(t, f, m, threshold=10)
    want = np.clip(np.moveaxis(ref, 1, -1), 0, 65535).astype(np.uint16)
    out = solve_raw(t, f, m, mask_threshold=10, out_layout="hwb", out_dtype=np.uint16)
    assert out.dtype == torch.uint16 and np.array_equal(out.numpy(), want)
    cloud = np.broadcast_to((m > 10)[..., None], t.shape)
    assert np.array_equal(out.numpy()[~cloud], t[~cloud]), "known cells come back exactly"
    img = torch.from_numpy(t.copy()).pin_memory()                # the caller's own product: only its cloud cells change
    solve_raw(t, f, m, mask_threshold=10, out=img, out_layout="hwb", out_dtype=np.uint16, result="unknown")
    assert np.array_equal(img.numpy(), want)


@pytest.mark.parametrize("raster", [np.float32, np.float64])
def test_raw_float_rasters_nan_clip_and_buffer(raster):
    """float rasters with their own NaNs (per band), clip_target / clip_features, cloud buffering (incl. two passes and
    the reference's column bound on a wide grid), DTYPE float64."""
    rng = np.random.default_rng(2003)
    H, W, B = 48, 80, 3                                   # H < W: columns >= H are never buffered (VPint2.py:146)
    t, f, m = raw_stack(rng, 3, H, W, B)
    t = t.astype(raster)
    f = f.astype(raster)
    t[rng.uniform(size=t.shape) < 0.002] = np.nan         # isolated NaNs in single bands
    for ctor in (dict(threshold=10, clip_target=2500, clip_features=(400.0, 2600.0)),
                 dict(threshold=10, buffer_mask=True, mask_buffer_size=3),
                 dict(threshold=10, buffer_mask=True, mask_buffer_size=2, mask_buffer_passes=2),
                 dict(threshold=10, dtype=np.float64)):
        ref, ref_sweeps = host_reference(t, f, m, **ctor)
        kw = dict(ctor)
        out, sweeps = solve_raw(t, f, m, mask_threshold=kw.pop("threshold"), return_sweeps=True, chunk_patches=2, **kw)
        assert np.array_equal(sweeps.numpy(), ref_sweeps), ctor
        assert np.array_equal(out.numpy(), ref, equal_nan=True), ctor
    # no mask at all: the NaNs of the target alone, buffered
    ref, ref_sweeps = host_reference(t, f, None, buffer_mask=True, mask_buffer_size=2)
    out, sweeps = solve_raw(t, f, None, buffer_mask=True, mask_buffer_size=2, return_sweeps=True)
    assert np.array_equal(sweeps.numpy(), ref_sweeps) and np.array_equal(out.numpy(), ref, equal_nan=True)


def test_device_constructor_pieces_match_host():
    """cloud_mask / buffer_clouds kernels against VPint2_interpolator.apply_mask / buffer_clouds on the host."""
    import torch
    rng = np.random.default_rng(2004)
    for H, W in ((40, 64), (64, 40), (50, 50)):
        t, f, m = raw_stack(rng, 2, H, W, 3)
        tf = t.astype(np.float32)
        tf[rng.uniform(size=tf.shape) < 0.003] = np.nan
        mask = rng.uniform(0, 1, (2, H, W))
        cl = cloud_mask(torch.from_numpy(mask).cuda(), 0.8)
        assert np.array_equal(cl.cpu().numpy() != 0, mask > 0.8)
        for size, passes in ((5, 1), (1, 1), (3, 2), (0, 1)):
            got = buffer_clouds(torch.from_numpy(tf).cuda(), cl.clone(), size, passes).cpu().numpy() != 0
            for i in range(2):
                vp = VPint2_interpolator(tf[i], tf[i], mask=mask[i], threshold=0.8, buffer_mask=True,
                                         mask_buffer_size=size, mask_buffer_passes=passes)
                host = np.isnan(vp.target)
                dev = got[i][..., None] | np.isnan(tf[i])
                assert np.array_equal(dev, host), (H, W, size, passes)


@pytest.mark.parametrize("raster,value", [(np.uint16, "float32"), (np.float32, "float32"), (np.float64, "float64"),
                                          (np.float32, "float64"), (np.uint16, "float64")])
@pytest.mark.parametrize("shape", [(256, 256, 13), (100, 100, 1), (37, 91, 3), (300, 301, 2), (1000, 1003, 2), (16, 8, 4),
                                   (2049, 1025, 1)])
def test_device_mean_init_any_size_is_numpy_nanmean(shape, raster, value):
    """fill_leaf_kernel + fill_combine_kernel: np.nanmean of every band IN the interpolator's dtype, bit for bit."""
    import torch
    H, W, B = shape
    rng = np.random.default_rng(H * 31 + W)
    n_img = 2 if H * W < 200000 else 1
    t = rng.uniform(1, 6000, (n_img, H, W, B))
    t = np.round(t).astype(raster) if raster == np.uint16 else t.astype(raster)
    cloud = rng.uniform(size=(n_img, H, W)) < 0.2
    if raster != np.uint16:
        t[rng.uniform(size=t.shape) < 0.01] = np.nan
    vdt = np.float32 if value == "float32" else np.float64
    solver = Solver(2, n_img * B, H, W, nprob_inner=B, planes=True, k_block=1)
    try:
        td = torch.from_numpy(t).cuda()
        cd = torch.from_numpy(cloud.astype(np.uint8)).cuda()
        solver.ingest_bands(td, td, cd, value_dtype=value)
        got = solver.fill_tensor().cpu().numpy().reshape(n_img, B)
    finally:
        solver.close()
    for i in range(n_img):
        for b in range(B):
            band = t[i, :, :, b].astype(vdt)
            band[cloud[i]] = np.nan
            want = np.nanmean(band.copy())          # contiguous copy, as MRP.__init__ holds it
            assert got[i, b] == np.float64(want), (i, b, got[i, b], want)


@pytest.mark.parametrize("world", [2, 3, 5])
def test_device_mean_init_row_sharded(world):
    """Row shards sum the leaves of NumPy's tree they own; the all-reduced leaf sums give every rank the np.nanmean of
    the WHOLE band (here: shards in one process, the all-reduce is a tensor sum)."""
    import torch
    from vpint_b200.sharded import RowShard
    rng = np.random.default_rng(2005 + world)
    H, W, B = 203, 160, 3
    t = np.round(rng.uniform(1, 6000, (H, W, B))).astype(np.uint16)
    cloud = discs(rng, (H, W), 4, 10, 40)
    want = []
    for b in range(B):
        band = t[:, :, b].astype(np.float32)
        band[cloud] = np.nan
        want.append(np.float64(np.nanmean(band.copy())))
    solvers, scr = [], []
    try:
        for r in range(world):
            sh = RowShard(H, world, r, halo=1)
            s = Solver(2, B, sh.local_rows, W, nprob_inner=B, planes=True, k_block=1, shard=sh.desc())
            solvers.append(s)
            sc = s.fill_scratch("float32")
            td = torch.from_numpy(np.ascontiguousarray(t[sh.lo:sh.hi])[None]).cuda()
            cd = torch.from_numpy(np.ascontiguousarray(cloud[sh.lo:sh.hi]).astype(np.uint8)[None]).cuda()
            s.fill_leaves(td, cd, sc, value_dtype="float32")
            scr.append(sc)
        leaves = [s.fill_views(sc, "float32") for s, sc in zip(solvers, scr)]
        tot_l = sum(l[0] for l in leaves)
        tot_c = sum(l[1] for l in leaves)
        for s, sc, (l, c) in zip(solvers, scr, leaves):
            l.copy_(tot_l); c.copy_(tot_c)
            s.fill_combine(sc, "float32")
            assert list(s.fill_tensor().cpu().numpy()) == want
    finally:
        for s in solvers:
            s.close()


def test_interpolator_runs_from_raw_and_keeps_its_attributes(golden):
    """VPint2_interpolator keeps the raw rasters and run() / run_serial() construct on the device; .target / .features
    still materialise on the host on demand, and assigning to them switches to the host arrays."""
    g = golden("vpint2")
    vp = VPint2_interpolator(g["target"], g["features"], mask=g["mask"], threshold=10)
    assert vp._raw is not None and vp._target is None
    p = vp.run()
    assert vp._target is None, "run() must not have needed the host-built arrays"
    assert p.shape == g["pred_par"].shape and p.dtype == np.float64
    assert np.abs(p - g["pred_par"]).max() <= 1e-9 * np.abs(g["pred_par"]).max()
    ps = vp.run_serial()
    assert ps.dtype == np.float32 and ps.shape == g["pred_serial"].shape
    vp2 = VPint2_interpolator(g["target"], g["features"], mask=g["mask"], threshold=10)
    tgt = vp2.target                                    # host construction on demand
    assert tgt.dtype == np.float32 and np.isnan(tgt).any()
    p2 = vp2.run()                                      # still the raw route (nothing was assigned)
    assert np.array_equal(p, p2)
    vp2.target = tgt.copy()
    assert vp2._raw is None
    assert np.array_equal(vp2.run(), p), "host-built arrays and device constructor must agree bit for bit"


def test_single_large_scene_from_raw_through_the_tiled_path():
    """A scene too large for the cluster-resident kernel: general mean init + fused load + tiled sweeps, against the
    oracle (the lazy constructor never builds the float arrays on the host)."""
    rng = np.random.default_rng(2006)
    H, W, B = 300, 700, 2
    t, f, m = raw_stack(rng, 1, H, W, B)
    vp = VPint2_interpolator(t[0], f[0], mask=m[0], threshold=10)
    out = vp.run()
    assert vp._target is None
    ref, ref_sweeps = O.vpint2_run(vp.target, vp.features)
    err = np.abs(out - ref) / np.abs(ref)
    assert err.max() <= 1e-9
    _, sweeps = solve_raw(t, f, m, mask_threshold=10, return_sweeps=True)
    assert list(sweeps[0].numpy()) == list(ref_sweeps)


@pytest.mark.parametrize("zeros", [False, True], ids=["float32-exact", "zero-features"])
def test_scene_sweeps_copy_engine_kernel_equals_register_kernel(zeros, monkeypatch):
    """Rows beyond the cluster-resident kernel: ratio-direct load + the copy-engine-fed ratio stencil (csrc/vpint_step_bulk.cu),
    with the float32 copy of the feature plane when every feature is a float32 value and the fp64 plane when not (a zero
    feature becomes 0.01, VPint/WP_MRP.py:143-145) -- bit for bit what the register-fed kernel gives, and the oracle's
    sweep counts."""
    import torch
    rng = np.random.default_rng(515)
    H, W, B = 600, 300, 3
    t, f, m = raw_stack(rng, 1, H, W, B, zeros=zeros)
    td, fd = torch.from_numpy(t).cuda(), torch.from_numpy(f).cuda()
    cd = torch.from_numpy((m > 10).astype(np.uint8)).cuda()

    def solve():
        s = Solver(2, B, H, W, nprob_inner=B, planes=True, k_block=1)
        try:
            block = torch.zeros(s.peer_bytes(1), dtype=torch.uint8, device="cuda")
            s.peer_bind(1, 0, [block.data_ptr()])
            s.ingest_bands(td, fd, cd, value_dtype="float32")
            s.run_sharded(5000, auto_terminate=True, threshold=1e-4)
            out, sweeps = s.result()
            return out.cpu().numpy(), sweeps.cpu().numpy()
        finally:
            s.close()

    out5, sw5 = solve()
    monkeypatch.setenv("VPINT_NO_F32_PLANE", "1")
    out5d, sw5d = solve()
    monkeypatch.setenv("VPINT_RATIO_KERNEL", "4")
    out4, sw4 = solve()
    assert np.array_equal(sw5, sw4) and np.array_equal(sw5d, sw4)
    assert np.array_equal(out5, out4) and np.array_equal(out5d, out4)
    vp = VPint2_interpolator(t[0], f[0], mask=m[0], threshold=10)
    ref, ref_sweeps = O.vpint2_run(vp.target, vp.features)
    assert list(sw5) == list(ref_sweeps)
    err = np.abs(np.moveaxis(out5, 0, -1) - ref) / np.abs(ref)
    assert err.max() <= 1e-9


def test_run_async_then_finish_equals_run():
    """vpint_solver_run_async + run_finish == vpint_solver_run, also when the resident kernel declines (inf feature)."""
    import torch
    rng = np.random.default_rng(2007)
    t, f, m = raw_stack(rng, 2, 64, 64, 3)
    tf, ff = t.astype(np.float32), f.astype(np.float32)
    tf[np.broadcast_to((m > 10)[..., None], tf.shape)] = np.nan
    for poison in (False, True):
        if poison:
            ff[1, 10, 10, 2] = np.inf
        outs = []
        for mode in ("run", "async"):
            s = Solver(2, 6, 64, 64, nprob_inner=3, planes=True, k_block=0)
            try:
                td, fd = torch.from_numpy(tf).cuda(), torch.from_numpy(ff).cuda()
                s.load(td.permute(0, 3, 1, 2))
                s.weights(fd.permute(0, 3, 1, 2).unsqueeze(-1))
                if mode == "run":
                    s.run(5000)
                else:
                    s.run_async(5000)
                    assert s.run_finish() == poison
                o, n = s.result()
                outs.append((o.cpu().numpy(), n.cpu().numpy()))
            finally:
                s.close()
        assert np.array_equal(outs[0][1], outs[1][1])
        assert np.array_equal(outs[0][0], outs[1][0], equal_nan=True)
