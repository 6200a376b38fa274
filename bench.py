#!/usr/bin/env python
"""Benchmark of the VPint2 value-propagation path on B200 (BASELINE.json metric:
VPint2 Mpixel*iter/s and HBM roofline fraction at 1/2/4/8 GPUs vs host-CPU NumPy).

Workload (BASELINE config 3): VPint2 cloud removal on synthetic Sentinel-2-shaped
256x256x13 patches, `--patches` patches PER GPU (default 512: 6656 independent
65 536-pixel band problems, 436 Mpx, ~21 GB of fp64 state + weights, far larger
than the 126 MB L2).  Patches are independent, so ranks shard them with no
data-path collective (weak scaling).

A "step" is one complete solve of the batch: constructor steps (cloud mask -> unknown cells), per-band
NaN -> mean init, feature planes, Jacobi sweeps to every problem's own stopping sweep (auto-terminate,
threshold 1e-4), result.  px*it = sum over problems of H*W*sweeps.

  value : device time, inputs (float32 HWB target with NaN clouds / features, as VPint2_interpolator stores them)
          already resident in HBM, full float64 (N, B, H, W) result left in HBM
  e2e   : the same solve through the public host API (vpint_b200.vpint2.RawPipeline.solve == solve_raw) from pinned
          host buffers holding what a Sentinel-2 product holds -- uint16 target / features and a cloud mask -- with
          the result written where it changes anything: the cloud cells of the caller's float32 (N, H, W, B) image
          (run_serial's layout).  H2D + D2H inside the timed region.  `e2e_variants` times the full-result contracts.
  scene : BASELINE config 4 -- ONE 10980x10980x13 scene row-sharded over the ranks (halo exchange + all-reduce of
          the stopping sums), with an in-run parity check of the sharded path against the unsharded solve
  --impl reference : the reference's own NumPy implementation on the host cores
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W, B = 256, 256, 13
BYTES_WP_F64 = 48.0      # algorithmic bytes per px*it, WP 2-D fp64 (SURVEY 8d / BASELINE.md)
METRIC = "vpint2_mpixel_iter_per_s"
UNIT = "Mpx*it/s"
MASK_THRESHOLD = 10      # notebook cell 2: VPint2_interpolator(target, features, mask=mask, threshold=10)


# ---------------------------------------------------------------------------
# synthetic Sentinel-2-shaped patches (SURVEY 8d, config 3 recipe)
# ---------------------------------------------------------------------------
def make_patches_numpy(n, seed):
    """(target, features, mask): uint16 (n, H, W, B) reflectances and a uint8 (n, H, W) cloud mask (0 / 100, 4 discs)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    yy /= H
    xx /= W
    target = np.empty((n, H, W, B), dtype=np.uint16)
    feats = np.empty((n, H, W, B), dtype=np.uint16)
    mask = np.zeros((n, H, W), dtype=np.uint8)
    for i in range(n):
        for b in range(B):
            acc = np.zeros((H, W))
            for _ in range(6):
                fy, fx = rng.uniform(0.3, 2.0, 2)
                py, px = rng.uniform(0, 1, 2)
                acc += rng.uniform(0.5, 1.0) * np.sin(2 * np.pi * (fy * yy + py)) * np.cos(2 * np.pi * (fx * xx + px))
            base = 1500.0 + 1200.0 * acc / (np.abs(acc).max() + 1e-9)
            target[i, :, :, b] = np.round(base + rng.normal(0, 20, (H, W))).clip(1, 65535)
            feats[i, :, :, b] = np.round(1.1 * base + rng.normal(0, 10, (H, W))).clip(1, 65535)
        cloud = np.zeros((H, W), dtype=bool)
        for _ in range(4):
            cy, cx, r = rng.uniform(0, 1), rng.uniform(0, 1), rng.uniform(15, 50) / H
            cloud |= (yy - cy) ** 2 + (xx - cx) ** 2 < r * r
        mask[i][cloud] = 100
    return target, feats, mask


def make_patches_torch(n, seed, device):
    """Same recipe generated on the device (inputs are synthetic; generation is not timed).  Returns the float32
    arrays as VPint2_interpolator stores them (NaN = cloud) AND the raw form they come from: uint16 rasters (the
    cloudy pixels keep their measured value) + the uint8 cloud mask."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    yy = (torch.arange(H, device=device, dtype=torch.float64) / H).view(1, H, 1, 1)
    xx = (torch.arange(W, device=device, dtype=torch.float64) / W).view(1, 1, W, 1)
    target = torch.empty((n, H, W, B), dtype=torch.float32, device=device)
    feats = torch.empty((n, H, W, B), dtype=torch.float32, device=device)
    t_raw = torch.empty((n, H, W, B), dtype=torch.uint16, device=device)
    f_raw = torch.empty((n, H, W, B), dtype=torch.uint16, device=device)
    mask = torch.empty((n, H, W), dtype=torch.uint8, device=device)
    chunk = 32
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        acc = torch.zeros((m, H, W, B), dtype=torch.float64, device=device)
        for _ in range(6):
            u = torch.rand((5, m, 1, 1, B), generator=g, device=device, dtype=torch.float64)
            fy, fx = 0.3 + 1.7 * u[0], 0.3 + 1.7 * u[1]
            acc += (0.5 + 0.5 * u[4]) * torch.sin(2 * np.pi * (fy * yy + u[2])) * torch.cos(2 * np.pi * (fx * xx + u[3]))
        base = 1500.0 + 1200.0 * acc / (acc.abs().amax(dim=(1, 2), keepdim=True) + 1e-9)
        nt = torch.randn((m, H, W, B), generator=g, device=device, dtype=torch.float64)
        nf = torch.randn((m, H, W, B), generator=g, device=device, dtype=torch.float64)
        t = torch.round(base + 20.0 * nt).clamp(1, 65535)
        f = torch.round(1.1 * base + 10.0 * nf).clamp(1, 65535)
        cloud = torch.zeros((m, H, W, 1), dtype=torch.bool, device=device)
        for _ in range(4):
            c = torch.rand((3, m, 1, 1, 1), generator=g, device=device, dtype=torch.float64)
            r = (15.0 + 35.0 * c[2]) / H
            cloud |= (yy - c[0]) ** 2 + (xx - c[1]) ** 2 < r * r
        t_raw[s:s + m] = t.to(torch.int32).to(torch.uint16)
        f_raw[s:s + m] = f.to(torch.int32).to(torch.uint16)
        mask[s:s + m] = cloud[..., 0].to(torch.uint8) * 100
        t = torch.where(cloud, torch.full_like(t, float("nan")), t)
        target[s:s + m] = t.to(torch.float32)
        feats[s:s + m] = f.to(torch.float32)
    return target, feats, t_raw, f_raw, mask


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def ready(self):
        """True once nvidia-smi has written its first sample (or cannot run)."""
        if self.proc is None or self.proc.poll() is not None:
            return True
        try:
            return os.path.getsize(self.path) > 0
        except OSError:
            return False

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 6:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except ValueError:
                    continue
                for nm, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            # median over the busier half of the samples = "under load"
            hi = sorted(sm)[len(sm) // 2:]
            out["sm_mhz"] = statistics.median(hi)
            out["sm_max_mhz"] = max(mx)
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


# ---------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's NumPy path on host cores
# ---------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def reference_available():
    """oracle/_ref holds a copy of the reference package (made by __graft_entry__.build() where /root/reference
    exists; git-ignored, shipped to the GPU box like the built library)."""
    return os.path.isdir(os.path.join(REF_DIR, "VPint"))


def _reference_patch(args):
    """One patch through the reference itself: VPint2_interpolator's constructor, then every band's WP_SMRP.run --
    what VPint2_interpolator.run_serial does (VPint/VPint2.py:109-119), with track_delta so that the sweep counts
    are known.  Returns (pixel-sweeps, pred (H, W, B) float64, sweeps)."""
    import warnings
    warnings.filterwarnings("ignore")
    np.product = np.prod                      # the reference targets numpy 1.26 (VPint/MRP.py:71)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    from VPint.VPint2 import VPint2_interpolator
    from VPint.WP_MRP import WP_SMRP
    target, feats, mask = args
    vp = VPint2_interpolator(target, feats, mask=mask, threshold=MASK_THRESHOLD)
    pred = np.empty(target.shape, dtype=np.float64)
    sweeps = []
    for b in range(target.shape[2]):
        p, d = WP_SMRP(vp.target[:, :, b], vp.features[:, :, b]).run(track_delta=True)
        pred[:, :, b] = p
        sweeps.append(len(d))
    return int(sum(sweeps)) * H * W, pred, sweeps


def _oracle_patch(args):
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import vpint_oracle as O
    from vpint_b200 import VPint2_interpolator
    target, feats, mask = args
    vp = VPint2_interpolator(target, feats, mask=mask, threshold=MASK_THRESHOLD)
    pred, sweeps = O.vpint2_run(vp.target, vp.features)
    return int(sweeps.sum()) * H * W, pred, [int(s) for s in sweeps]


def cpu_rate(target, feats, mask, procs):
    """Mpx*it/s of the reference (or, without oracle/_ref, of the oracle port) over the given raw patches."""
    fn = _reference_patch if reference_available() else _oracle_patch
    jobs = [(target[i], feats[i], mask[i]) for i in range(target.shape[0])]
    t0 = time.perf_counter()
    if procs <= 1:
        res = [fn(j) for j in jobs]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(fn, jobs, chunksize=1)
    dt = time.perf_counter() - t0
    pxit = sum(r[0] for r in res)
    return pxit / dt / 1e6, pxit, dt, res


def cpu_kind():
    return ("reference", "VPint 0.2.4 itself (oracle/_ref): VPint2_interpolator constructor + WP_SMRP.run per band") \
        if reference_available() else ("port", "NumPy oracle port of VPint2_interpolator.run (oracle/vpint_oracle.py)")


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(key):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["k_block"][str(key)]
    except Exception:
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = args.ref_patches if args.ref_patches > 0 else max(cores, 2)
    target, feats, mask = make_patches_numpy(n, 4242)
    for _ in range(args.warmup and 1):
        cpu_rate(target[:1], feats[:1], mask[:1], 1)
    rates, times = [], []
    for _ in range(args.steps):
        r, pxit, dt, _ = cpu_rate(target, feats, mask, cores)
        rates.append(r)
        times.append(dt)
    val = statistics.median(rates)
    kind, what = cpu_kind()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * statistics.median(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3: VPint2 256x256x13 patches (bounded sample of %d patches per step)" % n,
                   "patches": n, "H": H, "W": W, "bands": B},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d patches x 13 bands from uint16 rasters + cloud mask, auto-terminate; %s, one forked "
                                   "worker per core" % (n, what)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def prefer_gpu_memory_node(pci):
    """Host memory allocated from now on (the pinned buffers) prefers the NUMA node the GPU hangs off, even when the
    CPUs this process may run on sit elsewhere: set_mempolicy(MPOL_PREFERRED).  Returns a note for the bench line."""
    import ctypes
    import glob
    nodes = sorted(int(d.rsplit("node", 1)[1]) for d in glob.glob("/sys/devices/system/node/node[0-9]*"))
    if len(nodes) < 2 or not pci:
        return "%d NUMA node(s) visible" % len(nodes)
    dom, rest = pci.split(":", 1)
    with open("/sys/bus/pci/devices/%s:%s/numa_node" % (dom[-4:], rest)) as fh:
        node = int(fh.read().strip())
    if node < 0:
        return "%d NUMA nodes, GPU node unknown" % len(nodes)
    mask = ctypes.c_ulong(1 << node)
    libc = ctypes.CDLL(None, use_errno=True)
    rc = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(65))      # x86-64 set_mempolicy, MPOL_PREFERRED
    if rc != 0:
        return "%d NUMA nodes, GPU on node %d, set_mempolicy refused (errno %d)" % (len(nodes), node, ctypes.get_errno())
    return "%d NUMA nodes, host buffers prefer node %d (the GPU's)" % (len(nodes), node)


def bind_to_gpu_node(index, pci=None, local_ranks=1, local_rank=0):
    """Pin this process (and the threads it starts later) to CPUs NVML reports as local to GPU `index`, so that the
    pinned host buffers of the e2e leg are allocated on the GPU's NUMA node (PCIe traffic across the socket link runs
    at half speed).  When several ranks share the same CPU set (one NUMA node for all GPUs of the box), every rank
    takes its own slice of it instead of all ranks piling onto the same cores.  Where the box shows several NUMA
    nodes, the memory policy is pointed at the GPU's node as well (prefer_gpu_memory_node)."""
    note = "unbound"
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByPciBusId(pci.encode()) if pci else pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            note = "%d CPUs local to GPU %d" % (len(cpus), index)
            if local_ranks > 1 and len(cpus) >= 2 * local_ranks:
                ordered = sorted(cpus)
                per = len(ordered) // local_ranks
                cpus = set(ordered[local_rank * per:(local_rank + 1) * per])
                note += ", slice of %d for this rank" % len(cpus)
            os.sched_setaffinity(0, cpus)
    except Exception as exc:                      # NVML missing or affinity not permitted: run unbound
        note = "unbound (%s)" % type(exc).__name__
    if os.environ.get("VPINT_BENCH_NO_MEMPOLICY") != "1":
        try:
            note += "; " + prefer_gpu_memory_node(pci)
        except Exception as exc:
            note += "; memory policy untouched (%s)" % type(exc).__name__
    return note


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from vpint_b200.engine import Solver
    from vpint_b200.vpint2 import RawPipeline, solve_bands
    from vpint_b200.wp import wp_run_options

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    try:
        pr = torch.cuda.get_device_properties(local)
        pci = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
    except Exception:
        pci = None
    affinity = bind_to_gpu_node(local, pci, local_world, local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = args.patches
    target, feats, t_raw, f_raw, mask = make_patches_torch(n, 1003 + rank, dev)
    torch.cuda.synchronize()
    opts = wp_run_options()
    nprob = n * B
    total_px = nprob * H * W

    # host side of the e2e leg (pinned): the raw product + the caller's float32 image, whose cloud cells get filled in
    t_host = torch.empty(t_raw.shape, dtype=torch.uint16).pin_memory()
    f_host = torch.empty(f_raw.shape, dtype=torch.uint16).pin_memory()
    m_host = torch.empty(mask.shape, dtype=torch.uint8).pin_memory()
    t_host.copy_(t_raw)
    f_host.copy_(f_raw)
    m_host.copy_(mask)
    img_host = torch.empty((n, H, W, B), dtype=torch.float32).pin_memory()
    img_host.copy_(t_raw.to(torch.float32))          # the caller's own copy of the scene (cloudy pixels included)
    torch.cuda.synchronize()
    cloud_cells = int((mask > MASK_THRESHOLD).sum().item()) * B

    pipe = RawPipeline(dev, chunk_patches=args.chunk_patches, streams=args.streams, k_block=args.k_block)
    out_dev = torch.empty((n, B, H, W), dtype=torch.float64, device=dev)

    def device_step():
        _, sweeps = pipe.solve(target, feats, None, opts, out=out_dev, out_layout="bhw", result="full", return_sweeps=True)
        return sweeps

    def e2e_step():
        _, sweeps = pipe.solve(t_host, f_host, m_host, opts, mask_threshold=MASK_THRESHOLD, out=img_host, out_layout="hwb",
                               out_dtype=np.float32, result="unknown", return_sweeps=True)
        return sweeps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, per_step=False):
        """Total device time of `steps` calls (CUDA events, barrier + sync on both sides, max over ranks)."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = []
        barrier()
        ev0.record()
        sweeps_total = 0
        for _ in range(steps):
            nit = fn()
            sweeps_total += int(torch.as_tensor(nit).sum().item())
            if per_step:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append(e)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        each = []
        if per_step:
            prev = ev0
            for e in marks:
                each.append(prev.elapsed_time(e))
                prev = e
        if world > 1:
            tm = torch.tensor([ms] + each, dtype=torch.float64, device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ms, each = float(tm[0].item()), [float(v) for v in tm[1:].tolist()]
            sw = torch.tensor([sweeps_total], dtype=torch.float64, device=dev)
            dist.all_reduce(sw, op=dist.ReduceOp.SUM)
            sweeps_total = float(sw.item())
        return ms, sweeps_total, each

    for _ in range(max(args.warmup, 0)):
        int(torch.as_tensor(device_step()).sum().item())     # the same host-side reduction a timed step ends with
    extra_warmup = 0
    sampler = ClockSampler(local)
    if rank == 0:
        # nvidia-smi takes a few hundred ms to initialise NVML; keep the GPU busy with untimed steps meanwhile so that
        # neither its start-up nor a clock ramp after an idle wait falls into the timed region (it then samples every
        # 100 ms during the timed region)
        sampler.start()
        t_wait = time.perf_counter()
        while not sampler.ready() and time.perf_counter() - t_wait < 5.0:
            device_step()
            extra_warmup += 1
    launches0 = pipe.total_launches()
    ms_dev, sweeps_dev, _ = timed(device_step, args.steps)
    launches = pipe.total_launches() - launches0
    clocks = sampler.stop() if rank == 0 else {}
    pxit_dev = sweeps_dev * H * W           # whole job (all ranks)
    value = pxit_dev / (ms_dev * 1e-3) / 1e6
    dev_head = out_dev[:args.cpu_patches].cpu().numpy() if rank == 0 else None       # for the in-run parity check
    sweeps_head = device_step()[:args.cpu_patches].numpy() if rank == 0 else None

    # e2e through the host API
    for _ in range(min(max(args.warmup, 0), 3)):
        int(torch.as_tensor(e2e_step()).sum().item())
    e2e_steps = max(1, args.e2e_steps)
    ms_e2e, sweeps_e2e, each_e2e = timed(e2e_step, e2e_steps, per_step=True)
    e2e_val = sweeps_e2e * H * W / (ms_e2e * 1e-3) / 1e6
    e2e_ok = None
    if rank == 0:
        # the cloud cells the e2e leg wrote == the device leg's result (float32 of it), the rest untouched
        chk = min(n, 32)
        got = img_host[:chk].numpy()
        want = np.moveaxis(out_dev[:chk].cpu().numpy(), 1, -1).astype(np.float32)
        cl = (m_host[:chk].numpy() > MASK_THRESHOLD)
        e2e_ok = bool(np.array_equal(got[cl], want[cl]) and np.array_equal(got[~cl], t_host[:chk].numpy()[~cl].astype(np.float32)))

    # the full-result contracts, a few steps each (reported beside the headline e2e, never instead of it)
    variants = {}
    if not args.no_variants and world == 1:         # they explain the one-GPU number; 12 GB more pinned memory per rank
        out_f32 = torch.empty((n, H, W, B), dtype=torch.float32).pin_memory()
        out_f64 = torch.empty((n, B, H, W), dtype=torch.float64).pin_memory()
        t32_host = torch.empty(target.shape, dtype=torch.float32).pin_memory()
        f32_host = torch.empty(feats.shape, dtype=torch.float32).pin_memory()
        t32_host.copy_(target)
        f32_host.copy_(feats)
        torch.cuda.synchronize()
        legs = {
            "uint16_in_full_float32_hwb_out": (
                lambda: pipe.solve(t_host, f_host, m_host, opts, mask_threshold=MASK_THRESHOLD, out=out_f32, out_layout="hwb",
                                   out_dtype=np.float32, result="full", return_sweeps=True)[1],
                t_host.numel() * 4 + m_host.numel(), out_f32.numel() * 4 + nprob * 4,
                "run_serial() contract: every cell of the float32 (N, H, W, B) result"),
            "uint16_in_full_float64_bhw_out": (
                lambda: pipe.solve(t_host, f_host, m_host, opts, mask_threshold=MASK_THRESHOLD, out=out_f64, out_layout="bhw",
                                   result="full", return_sweeps=True)[1],
                t_host.numel() * 4 + m_host.numel(), out_f64.numel() * 8 + nprob * 4,
                "run() contract: every cell of the float64 (N, B, H, W) result"),
            "float32_in_full_float64_bhw_out_round1_path": (
                lambda: solve_bands(t32_host, f32_host, opts, out_layout="bhw", return_sweeps=True, device=dev,
                                    k_block=args.k_block, out=out_f64)[1],
                t32_host.numel() * 8, out_f64.numel() * 8 + nprob * 4,
                "round 1's e2e leg: host-built float32 arrays in, float64 result out, six host threads"),
        }
        for name, (fn, h2d, d2h, what) in legs.items():
            int(torch.as_tensor(fn()).sum().item())
            ms_v, sw_v, _ = timed(fn, args.variant_steps)
            variants[name] = {"ms_per_step": ms_v / args.variant_steps, "value": sw_v * H * W / (ms_v * 1e-3) / 1e6, "unit": UNIT,
                              "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": args.variant_steps,
                              "what": what}
        del out_f32, out_f64, t32_host, f32_host

    # roofline of the dominant kernel, timed alone with CUDA events on the stream it is launched on
    masked_frac = float(torch.isnan(target).double().mean().item())
    masked_px = masked_frac * total_px
    unknown_pp = torch.isnan(target[..., 0]).sum(dim=(1, 2)).double()          # unknown cells per patch (all bands alike)
    peak, peak_src = peak_hbm()
    solver = Solver(2, nprob, H, W, nprob_inner=B, planes=True, k_block=args.k_block, device=dev)
    kb = solver.k_block
    resident = (args.k_block == 0)
    scratch = solver.fill_scratch("float32")
    roof = {}
    if resident:
        # one launch = the whole run(): a cluster per problem, all sweeps on chip (vpint_resident.cu)
        times, units, psw = [], 0.0, 0.0
        for _ in range(max(2, args.roofline_launches // 4)):
            solver.ingest_bands(target, feats, None, value_dtype="float32", scratch=scratch)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            solver.run(opts["iterations"], auto_terminate=True, threshold=opts["threshold"], clip_val=opts["clip_val"])
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
            _, nit = solver.result(out=out_dev)
            units = float((nit.view(n, B).double() * unknown_pp.view(n, 1)).sum().item())   # unknown cell x sweep
            psw = float(nit.double().sum().item())                                           # problem x sweep
        k_ms = statistics.mean(times)
        kernel_name = "resident2d_kernel<ratio> (one launch = run() of every problem)"
        unit_def = "one unknown cell updated for one sweep (%.3e per launch: sum over problems of unknown cells x sweeps)" % units
        e = load_traffic("resident")
        onchip, traffic, traffic_src = None, None, None
        if e:
            # what bounds a kernel whose state never leaves the chip: shared-memory wavefronts and issue slots (ncu
            # capture of THIS configuration, counts per problem-sweep x the problem-sweeps of the timed launch)
            clk = 1e6 * (clocks.get("sm_mhz") or 1965.0)
            sms = e["sms_with_a_cta"]
            onchip = {"problem_sweeps_per_launch": psw,
                      "smem_wavefronts_per_s": e["smem_wavefronts_per_problem_sweep"] * psw / (k_ms * 1e-3),
                      "smem_peak_wavefronts_per_s": sms * clk,
                      "warp_instructions_per_s": e["warp_instructions_per_problem_sweep"] * psw / (k_ms * 1e-3),
                      "issue_peak_per_s": 4 * sms * clk, "sms_with_a_cta": sms, "source": e["source"]}
            onchip["smem_frac"] = onchip["smem_wavefronts_per_s"] / onchip["smem_peak_wavefronts_per_s"]
            onchip["issue_frac"] = onchip["warp_instructions_per_s"] / onchip["issue_peak_per_s"]
            traffic, traffic_src = e["dram_bytes_per_unit"] * nprob, e["source"]
        alg_bytes = BYTES_WP_F64 * units
        hbm_equiv = alg_bytes / (k_ms * 1e-3) / 1e9
        roof = {"bound": "smem", "kernel": kernel_name, "kernel_ms": k_ms,
                "achieved": onchip["smem_wavefronts_per_s"] / 1e9 if onchip else None,
                "peak": onchip["smem_peak_wavefronts_per_s"] / 1e9 if onchip else None, "unit": "Gwavefront/s",
                "frac": onchip["smem_frac"] if onchip else None,
                "traffic": traffic, "traffic_source": traffic_src,
                "onchip": onchip,
                "hbm_equiv": {"achieved_gbps": hbm_equiv, "peak_gbps": peak, "frac": hbm_equiv / peak, "peak_source": peak_src,
                              "alg_bytes_per_launch": alg_bytes, "bytes_per_unit": BYTES_WP_F64, "unit_def": unit_def,
                              "whole_grid_equiv_gbps": value * BYTES_WP_F64 / 1e3 / max(world, 1),
                              "note": "48 B x unknown-cell sweeps / kernel time: an EQUIVALENT bandwidth -- the state never "
                                      "leaves shared memory, so it may exceed the HBM peak; whole_grid_equiv = value x 48 B "
                                      "per GPU (what the reference's sweep would have touched)"},
                "note": "state stays in shared memory for all sweeps: DRAM traffic is one read + one write per problem; the "
                        "kernel is bound by shared-memory wavefronts / issue slots, which `frac` reports"}
    else:
        solver.ingest_bands(target, feats, None, value_dtype="float32", scratch=scratch)
        solver.begin_run(10 ** 6, auto_terminate=False)
        for _ in range(3):
            solver.step_begin(); solver.step_end()
        times = []
        for _ in range(args.roofline_launches):
            times.append(solver.step_profile())
            solver.step_end()
        torch.cuda.synchronize()
        k_ms = statistics.mean(times)
        units = masked_px * kb
        e = load_traffic(kb)
        alg_bytes = BYTES_WP_F64 * units
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "jacobi2d_%s (stencil launch, all problems active)" % ("k1" if kb == 1 else "tb"),
                "kernel_ms": k_ms, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": e["dram_bytes_per_unit"] * masked_px if e else None, "traffic_source": e["source"] if e else None,
                "peak_source": peak_src, "alg_bytes_per_launch": alg_bytes, "bytes_per_unit": BYTES_WP_F64,
                "unit_def": "one unknown cell updated for one sweep (%.0f cells x %d sweeps per launch)" % (masked_px, kb),
                "note": "streams state and weights from HBM every launch"}
    active_tiles, total_tiles = solver.active_tiles, solver.total_tiles
    solver.close()

    line = None
    if rank == 0:
        srt = sorted(each_e2e) if each_e2e else [ms_e2e / e2e_steps]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C3: VPint2 cloud removal, 256x256x13 Sentinel-2-shaped patches, %d patches per GPU "
                                   "(%d band problems, %.0f Mpx per GPU), auto-terminate 1e-4" % (n, nprob, total_px / 1e6),
                       "baseline_config": "BASELINE.json configs[2] (the metric is quoted on VPint2; configs[1] is a single "
                                          "WP_SMRP grid; configs[3], the row-sharded scene, is the `scene` record)",
                       "patches_per_gpu": n, "H": H, "W": W, "bands": B, "k_block": args.k_block,
                       "path": ("fused load + cluster-resident solve (k_block 0), %d-patch chunks on %d streams"
                                % (pipe.chunk, pipe.K)) if args.k_block == 0 else "tiled, %d sweeps per launch" % kb,
                       "cloud_fraction": round(masked_frac, 4),
                       "l2_note": "state+weights %.1f GB per GPU >> 126 MB L2; no flush needed" % (48.0 * total_px / 1e9),
                       "mean_sweeps_per_problem": sweeps_dev / args.steps / (nprob * world),
                       "extra_warmup_steps": extra_warmup},
            "roofline": roof,
            "e2e": {"value": e2e_val, "unit": UNIT,
                    "h2d_bytes_per_step": int(t_host.numel() * 2 + f_host.numel() * 2 + m_host.numel()),
                    "d2h_bytes_per_step": int(cloud_cells * 4 + nprob * 4), "steps": e2e_steps,
                    "ms_per_step": ms_e2e / e2e_steps, "ms_per_step_min": srt[0], "ms_per_step_median": statistics.median(srt),
                    "ms_per_step_max": srt[-1],
                    "api": "vpint_b200.vpint2.RawPipeline.solve (= solve_raw): pinned uint16 target + features and uint8 cloud "
                           "mask in; the cloud cells of the caller's pinned float32 (N, H, W, B) image are written in place "
                           "through its device mapping (result='unknown'); one host thread, %d streams" % pipe.K,
                    "result_check": e2e_ok, "cpu_affinity": affinity, "variants": variants},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
    pipe.close()

    if rank == 0 and not args.no_cpu_baseline and world == 1:
        cn = args.cpu_patches
        ct, cf, cm = t_host[:cn].numpy(), f_host[:cn].numpy(), m_host[:cn].numpy()
        r, pxit, dt, res = cpu_rate(ct, cf, cm, 1)
        kind, what = cpu_kind()
        line["cpu_baseline"] = {"value": r, "unit": UNIT, "cores": 1, "kind": kind,
                                "sample": "first %d patches (%d band problems) of the same workload from their uint16 rasters, "
                                          "%s, %.1f s" % (cn, cn * B, what, dt)}
        # the same patches on the GPU (device leg): values and stopping sweeps against the CPU run
        max_rel, sweeps_equal = 0.0, True
        for i in range(cn):
            ref = res[i][1]
            got = np.moveaxis(dev_head[i], 0, -1)
            max_rel = max(max_rel, float(np.max(np.abs(got - ref) / np.abs(ref))))
            sweeps_equal = sweeps_equal and list(res[i][2]) == [int(v) for v in sweeps_head[i]]
        line["parity_check"] = {"against": kind, "patches": cn, "max_rel": max_rel, "sweeps_equal": bool(sweeps_equal),
                                "tolerance": 1e-9, "ok": bool(max_rel <= 1e-9 and sweeps_equal)}
    elif rank == 0:
        line["cpu_baseline"] = None

    # BASELINE config 4 in the same run: one scene row-sharded over the ranks, with a parity check of the sharded path
    if not args.no_scene:
        import bench_scene
        rec = bench_scene.scene_record(world, rank, dev, size=args.scene_size, parity_size=args.scene_parity_size,
                                       steps=args.scene_steps)
        if rank == 0:
            line["scene"] = rec
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="vpint_b200", choices=["vpint_b200", "reference"])
    ap.add_argument("--patches", type=int, default=512, help="patches per GPU")
    ap.add_argument("--k-block", type=int, default=0)
    ap.add_argument("--chunk-patches", type=int, default=0, help="patches per pipeline chunk (0 = library default)")
    ap.add_argument("--streams", type=int, default=0, help="streams of the pipeline (0 = library default)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--variant-steps", type=int, default=3)
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--roofline-launches", type=int, default=20)
    ap.add_argument("--cpu-patches", type=int, default=3)
    ap.add_argument("--ref-patches", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scene", action="store_true")
    ap.add_argument("--scene-size", type=int, default=10980)
    ap.add_argument("--scene-parity-size", type=int, default=2048)
    ap.add_argument("--scene-steps", type=int, default=2)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
