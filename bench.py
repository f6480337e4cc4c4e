#!/usr/bin/env python
"""Benchmark of the plane-sweep hot path (BASELINE.json metric: cost-volume Gvoxels/s, ref views/s,
% of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg4|cfg1|cfg3|cfg5|fuse] [--impl ours|reference]

One "step" = one reference view of the workload: feature relayout (V launches), the fused warp+aggregate kernel (1 launch,
the dominant one) and the fused softmax/regression kernel (1 launch).  The CNN regulariser between the two is out of scope
on both arms, so the logit volume is a synthetic input resident in HBM.  Under torchrun every rank processes its own K
reference views (weak scaling by reference-view sharding, no collective on the data path); the only communication is the
barrier and the max/sum joins of the timing.

`value` is measured with inputs resident in HBM.  `e2e` runs the same step from pinned HOST buffers through the public API
(deep3d_aerial_b200.pipeline.ViewPipeline): every rank walks a scene block in order and a reference view uploads the
feature maps of the images that are not resident on its GPU yet (consecutive views share 4 of their 5 images), its
cameras and hypotheses, and downloads its depth + confidence maps.  `--impl reference` times the reference's CPU PyTorch
path (the oracle restatement, which calls the same ATen ops) on the host cores.

Workloads: cfg2 (default) is the configuration BASELINE.json's metric is quoted on; cfg4 its group-wise-correlation variant;
cfg3 the AdaMVS 3-stage cascade per view at 1856 x 2752; cfg5 ONE scene block of --block-views reference views dealt to
the ranks (strong scaling, views/s); fuse the depth-map fusion consistency check.  The default N = 1
run appends short cfg4 / cfg3 / cfg5 measurements to its line as `extra_workloads` (--no-extra skips them).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (V, C, D, H, W, mode, groups, description)
    "cfg1": (3, 8, 48, 128, 160, "variance", 0, "tiny variance V=3 C=8 D=48 128x160"),
    "cfg2": (5, 32, 384, 688, 464, "variance", 0, "WHU-OMVS V=5 C=32 D=384 688x464 (1/4 of 2752x1856) variance"),
    "cfg2v3": (3, 32, 384, 688, 464, "variance", 0, "WHU-OMVS shape with V=3 (two source views): C=32 D=384 688x464 variance"),
    "cfg4": (5, 32, 384, 688, 464, "gwc", 8, "WHU-OMVS V=5 C=32 D=384 688x464 group-wise correlation G=8"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS) + ["cfg3", "cfg5", "fuse"],
                    help="cfg2 = the configuration the metric is quoted on; cfg3 = the AdaMVS 3-stage cascade; "
                         "cfg5 = a scene block of --block-views reference views dealt to the ranks (strong scaling)")
    ap.add_argument("--block-views", type=int, default=64, help="cfg5: reference views in the scene block")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", type=int, default=0, help="kernel variant (A/B; 0 = production)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true",
                    help="default run only: skip the short cfg4 / cfg3 / cfg5 measurements appended as extra_workloads")
    ap.add_argument("--debug-flat-hyps", action="store_true",
                    help="DIAGNOSTIC: all depth planes equal, so no footprint ever moves (isolates the re-fetch cost)")
    return ap.parse_args()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            js = json.load(f)
        for key in ("hbm_gbs", "hbm_gb_s", "hbm_copy_gbs"):
            if key in js:
                return float(js[key]), "measured (MEASURED_PEAKS.json:%s)" % key
    except (OSError, ValueError):
        pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def measured_traffic(wl):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/r2_traffic.json, else
    round 1's); None when there is no capture for this workload."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                rec = json.load(f).get(wl)
            if rec:
                return int(rec["dram_bytes_read"]) + int(rec["dram_bytes_write"])
        except (OSError, ValueError, KeyError):
            pass
    return None


class ClockSampler(threading.Thread):
    """Samples the SM clock and the throttle reasons of one GPU WHILE the timed region runs: NVML in-process
    every few milliseconds (the timed region of the default run is ~0.3 s, too short for `nvidia-smi -lms`,
    which needs most of that to start); falls back to nvidia-smi when NVML cannot be loaded."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = self._physical_index(index)
        self.sm, self.mx, self.seen = [], [], set()
        self.proc = None
        self.halt = threading.Event()
        self.how = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if index < len(ids) and ids[index].isdigit():
                return int(ids[index])
        return index

    def _run_nvml(self):
        import pynvml as nv

        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        bits = {"hw_slowdown": int(getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                "hw_thermal_slowdown": int(getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                "sw_thermal_slowdown": int(getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                "sw_power_cap": int(getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4))}
        try:
            reasons = nv.nvmlDeviceGetCurrentClocksEventReasons
        except AttributeError:
            reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        self.how = "nvml"
        while not self.halt.is_set():
            self.sm.append(int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            self.mx.append(int(mx))
            r = int(reasons(h))
            for name, bit in bits.items():
                if r & bit:
                    self.seen.add(name)
            time.sleep(0.004)

    def _run_smi(self):
        self.how = "nvidia-smi"
        self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits", "-lms", "100"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            if r and r[0].isdigit():
                self.sm.append(int(r[0]))
            if len(r) > 1 and r[1].isdigit():
                self.mx.append(int(r[1]))
            for i, name in enumerate(self.NAMES):
                if len(r) > 2 + i and r[2 + i].lower().startswith("active"):
                    self.seen.add(name)

    def run(self):
        try:
            self._run_nvml()
        except Exception:  # noqa: BLE001 - no NVML binding / driver mismatch: use the CLI
            try:
                self._run_smi()
            except OSError:
                pass

    def stop(self):
        self.halt.set()
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        return {"sm_mhz": int(statistics.median(self.sm)) if self.sm else None,
                "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": [n for n in self.NAMES if n in self.seen], "samples": len(self.sm), "source": self.how}


def make_inputs(wl, seed, device):
    import numpy as np  # noqa: F401
    import torch

    from deep3d_aerial_b200 import synth

    v, c, d, h, w, mode, groups, _ = WORKLOADS[wl]
    if wl == "cfg1":
        rig = synth.tiny_rig(num_views=v, width=w * 4, height=h * 4)
    else:
        rig = synth.make_rig(num_views=v)
    g = torch.Generator(device="cpu").manual_seed(seed)
    feats = torch.randn(v, c, h, w, generator=g, dtype=torch.float32)
    proj = torch.from_numpy(rig.proj(4))
    hyps = synth.uniform_hypotheses(rig.dmin, rig.dmax, d)
    logits = 4.0 * torch.randn(d, h, w, generator=g, dtype=torch.float32)
    if device is not None:
        feats, proj, hyps, logits = (t.to(device) for t in (feats, proj, hyps, logits))
    return feats, proj, hyps, logits


def algorithmic_bytes(wl):
    """kernel 1 writes 4*Cout B/voxel and reads every feature map once (+ the hypotheses);
    kernel 2 reads 4 B/voxel and writes depth, conf, index maps."""
    v, c, d, h, w, mode, groups, _ = WORKLOADS[wl]
    vox = d * h * w
    cout = groups if mode == "gwc" else c
    k1 = 4 * cout * vox + 4 * v * c * h * w + 4 * d
    k2 = 4 * vox + 4 * d + 12 * h * w
    relayout = 2 * 4 * v * c * h * w
    return vox, k1, k2, relayout


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path (oracle restatement = the same ATen calls,
    pinned bit-exact to the live reference by tests/test_oracle.py), all host threads, on a bounded
    sample of the workload: a subset of the D planes at the full image size."""
    import torch

    from oracle import sweep_torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    torch.set_grad_enabled(False)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = args.workload
    v, c, d, h, w, mode, groups, desc = WORKLOADS[wl]
    feats, proj, hyps, logits = make_inputs(wl, 0, None)
    views = [feats[i:i + 1] for i in range(v)]
    projb = proj.unsqueeze(0)

    # Where a reference checkout is present (this authoring container, or DEEP3D_REFERENCE_ROOT; never the GPU box), the
    # variance workload runs the LIVE reference instead of the restatement: cas_mvsnet.DepthNet.forward (warp, variance,
    # softmax, regression, confidence; cas_mvsnet.py:35-78) per chunk of 4 planes, its regulariser argument handing back
    # the synthetic logits of the chunk.
    kind, live_net = "port", None
    if mode != "gwc":
        try:
            from oracle import ref_live
            if ref_live.available():
                live_net = ref_live.load().cas_mvsnet.DepthNet().eval()
                kind = "reference"
        except Exception:  # noqa: BLE001 -- no checkout, or it does not import here: the port is the same ATen calls
            live_net = None

    def step(planes):
        sub = hyps[:planes].unsqueeze(0)
        if mode == "gwc":
            vol = sweep_torch.groupwise_correlation_volume(views, projb, sub, groups)
            float(vol[..., ::7, ::5].sum())
            sweep_torch.regress_maxprob(logits[:planes].unsqueeze(0), sub)
        elif live_net is not None:
            for d0 in range(0, planes, 4):
                n = min(4, planes - d0)
                chunk = logits[d0:d0 + n].view(1, 1, n, h, w)
                out = live_net(views, projb, hyps[d0:d0 + n].view(1, n, 1, 1).expand(1, n, h, w), n, lambda vol, c=chunk: c)
                float(out["depth"][0, 0, 0])
        else:
            sweep_torch.cpu_step_variance(views, projb, sub, logits[:planes].unsqueeze(0), plane_chunk=4)

    # calibrate the sample so the whole run stays within ~150 s
    t0 = time.perf_counter()
    step(2)
    per_plane = (time.perf_counter() - t0) / 2
    budget = 150.0 / max(1, args.steps + args.warmup)
    planes = int(max(1, min(d, budget / max(per_plane, 1e-6))))
    for _ in range(args.warmup):
        step(planes)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(planes)
    dt = time.perf_counter() - t0
    vox = planes * h * w
    value = vox * args.steps / dt / 1e9
    sample = "%d of %d depth planes at full %dx%d, V=%d C=%d, plane chunks of 4" % (planes, d, h, w, v, c)
    line = {
        "impl": "reference", "metric": "cost-volume Gvoxels/s", "value": value, "unit": "Gvoxel/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Gvoxel/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Gvoxel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline(wl):
    """Bounded CPU sample beside the GPU number (rank 0, N=1): ~10-30 s of the reference's CPU path."""
    import torch

    from oracle import sweep_torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    v, c, d, h, w, mode, groups, _ = WORKLOADS[wl]
    feats, proj, hyps, logits = make_inputs(wl, 0, None)
    views = [feats[i:i + 1] for i in range(v)]
    projb = proj.unsqueeze(0)

    def step(planes):
        sub = hyps[:planes].unsqueeze(0)
        if mode == "gwc":
            sweep_torch.groupwise_correlation_volume(views, projb, sub, groups)
            sweep_torch.regress_maxprob(logits[:planes].unsqueeze(0), sub)
        else:
            sweep_torch.cpu_step_variance(views, projb, sub, logits[:planes].unsqueeze(0), plane_chunk=4)

    t0 = time.perf_counter()
    step(2)
    per_plane = (time.perf_counter() - t0) / 2
    planes = int(max(1, min(d, 6.0 / max(per_plane, 1e-6))))
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        step(planes)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": planes * h * w / best / 1e9, "unit": "Gvoxel/s", "cores": cores, "kind": "port",
            "sample": "%d of %d depth planes at full %dx%d (best of 2), torch %s CPU" % (planes, d, h, w, torch.__version__)}


def block_image_ids(i, n_block, v):
    """Image ids of reference view i of a scene block of n_block images (the viewpair structure of the reference's
    workspace: the V - 1 nearest neighbours along the strip, datasets/data_io.py viewpair lists): the reference image
    first, then the other images of the window of V consecutive images that contains it."""
    start = min(max(i - (v - 1) // 2, 0), max(n_block - v, 0))
    return [i] + [j for j in range(start, start + v) if j != i]


def e2e_block(args, wl, dev, rank, slots, views, n_block, warm_views=2):
    """Runs the reference views `views` (indices into a scene block of `n_block` images) through
    pipeline.ViewPipeline from pinned host memory, in order; returns the e2e record with the wall-clock `seconds`
    of the timed part (the caller joins the ranks).  A few views of ANOTHER block warm the pipeline up first, so
    nothing of the timed block is resident when the clock starts: its first view uploads all V images."""
    import torch

    from deep3d_aerial_b200 import shard, sweep
    from deep3d_aerial_b200.pipeline import ViewPipeline

    v, c, d, h, w, mode, groups, _ = WORKLOADS[wl]
    agg = sweep.AGG_GROUP_CORR if mode == "gwc" else sweep.AGG_VARIANCE
    pipe = ViewPipeline(v, c, h, w, d, dev, mode=agg, groups=groups, variant=args.variant, resident_images=3 * v)
    # host side of the block: a pool of distinct pinned per-image feature maps (image j -> pool[j % len(pool)])
    g = torch.Generator(device="cpu").manual_seed(4321 + rank)
    pool = [torch.randn(c, h, w, generator=g, dtype=torch.float32).pin_memory() for _ in range(2 * v + 1)]
    host = [tuple(t.cpu().pin_memory() for t in (sl[1], sl[2])) for sl in slots]
    logit_slots = [sl[3] for sl in slots]
    sink = 0.0

    def run(view_ids, block, tag):
        nonlocal sink
        for k, i in enumerate(view_ids):
            ids = block_image_ids(i, block, v)
            hp, hh = host[k % 2]
            lg = logit_slots[k % 2]
            pipe.submit([pool[j % len(pool)] for j in ids], hp, hh, lambda vol, lg=lg: lg,    # regulariser: out of scope
                        image_ids=[(tag, j) for j in ids])
            if k:
                dep, conf = pipe.collect()
                sink += float(dep[0, 0]) + float(conf[0, 0])      # the host touches every result
        if view_ids:
            dep, conf = pipe.collect()
            sink += float(dep[0, 0]) + float(conf[0, 0])

    run(list(range(warm_views)), warm_views + v, "warm-up block")
    views = list(views)
    b0, n0 = pipe.h2d_bytes_total, pipe.views_submitted
    shard.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(views, n_block, "block")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    done = max(1, pipe.views_submitted - n0)
    if not math.isfinite(sink):
        raise SystemExit("bench.py: the end-to-end pass produced non-finite maps")
    return {"seconds": dt, "h2d_bytes_per_step": (pipe.h2d_bytes_total - b0) // done, "d2h_bytes_per_step": pipe.d2h_bytes,
            "h2d_bytes_per_step_without_image_residency": pipe.h2d_bytes,
            "ms_per_step": dt / done * 1e3, "steps": len(views), "timer": "host wall clock around submit/collect",
            "stream": "a scene block walked in order (viewpair windows of %d consecutive images): a view uploads only "
                      "the images not yet resident on its GPU (LRU of %d per-image maps kept as channels-last texels: an "
                      "image is laid out once, when it arrives, and the sweep names its views by pool slot, whereas the "
                      "device-timed `value` step lays out all %d maps of its view); the first view of the timed block "
                      "uploads all %d" % (v, 3 * v, v, v),
            "image_hits": pipe.lru.hits, "image_misses": pipe.lru.misses}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch

    from deep3d_aerial_b200 import _lib, shard, sweep

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the sweep engine has no CPU path")
    _lib.load()
    torch.set_grad_enabled(False)
    rank, world, local = shard.init()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    wl = args.workload
    v, c, d, h, w, mode, groups, desc = WORKLOADS[wl]
    agg = sweep.AGG_GROUP_CORR if mode == "gwc" else sweep.AGG_VARIANCE
    cout = groups if mode == "gwc" else c
    vox, b1, b2, brel = algorithmic_bytes(wl)

    # this rank's reference views: distinct seeds per (rank, slot); two input slots alternate so
    # consecutive steps never reuse a resident input (and the 15.7 GB output is far beyond L2 anyway)
    slots = [make_inputs(wl, 1000 * rank + s, dev) for s in range(2)]
    if args.debug_flat_hyps:
        slots = [(f, pr, torch.full_like(hy, float(hy.mean())), lg) for f, pr, hy, lg in slots]
    texels = torch.empty((v, h, w, c), device=dev)
    volume = torch.empty((cout, d, h, w), device=dev)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]

    def step(i, events=None):
        feats, proj, hyps, logits = slots[i % 2]
        sweep.to_texels(feats, out=texels)
        pose = sweep.relative_poses(proj)                       # module.py:528, per source view
        rays = sweep.rays_for(pose, h, w)                       # module.py:538 where the kernel's order differs
        if events:
            events[0].record()
        sweep.cost_volume(texels, pose, hyps, agg, groups=groups, out=volume, variant=args.variant, rays=rays)
        if events:
            events[1].record()
            events[2].record()
        r = sweep.depth_regress(logits, hyps)
        if events:
            events[3].record()
        return r

    for i in range(args.warmup):
        step(i)
    sampler = ClockSampler(local)
    sampler.start()
    shard.barrier()
    torch.cuda.synchronize()
    launches0 = _lib.launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        step(i, ev[i])
    t_end.record()
    torch.cuda.synchronize()
    shard.barrier()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop()
    ms_total = shard.join_max(t_start.elapsed_time(t_end))
    ms_step = ms_total / args.steps
    total_vox = shard.join_sum(vox * args.steps)
    value = total_vox / (ms_total * 1e-3) / 1e9
    k1_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in ev)
    k2_ms = statistics.mean(e[2].elapsed_time(e[3]) for e in ev)
    peak, peak_src = measured_peak()
    achieved = b1 / (k1_ms * 1e-3) / 1e9

    # ---- end to end from pinned host buffers through the public API (deep3d_aerial_b200.pipeline):
    # every rank walks its own scene block in order; a reference view copies host->device the feature maps of the
    # images that are not on the device yet (consecutive views of a block share 4 of their 5 images, see
    # block_image_ids), its cameras and hypotheses, and device->host its depth + confidence maps; neighbouring
    # views' copies overlap the sweep (two input slots, copy + compute streams).  The timed block starts on images
    # the device has never seen: its first view uploads all V of them.
    e2e = None
    if not args.no_e2e:
        del texels, volume
        torch.cuda.empty_cache()
        n_e2e = max(3, args.steps)
        e2e = e2e_block(args, wl, dev, rank, slots, range(n_e2e), n_e2e + v - 1)
        e2e = dict({"value": shard.join_sum(vox * n_e2e) / shard.join_max(e2e.pop("seconds")) / 1e9, "unit": "Gvoxel/s"}, **e2e)

    total_launches = int(shard.join_sum(launches))
    if rank != 0:
        return None
    line = {
        "metric": "cost-volume Gvoxels/s", "value": value, "unit": "Gvoxel/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "views_per_step_per_gpu": 1, "sharding": "reference views, no collective",
                   "l2": "inputs alternate between two slots; features (%.0f MB) and volume (%.2f GB) exceed the 126 MB L2"
                         % (4 * v * c * h * w / 1e6, 4 * cout * vox / 1e9),
                   "variant": args.variant},
        "ref_views_per_s": world * 1e3 / ms_step,
        "kernel_ms": {"sweep": k1_ms, "regress": k2_ms},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": measured_traffic(wl), "kernel": "fused sweep (warp+aggregate)", "algorithmic_bytes": b1,
                     "peak_source": peak_src, "frac_of_8TBps_nominal": achieved / 8000.0},
        "clocks": clocks,
        "gpu_launches": total_launches,
    }
    if e2e:
        line["e2e"] = e2e
    if mode == "gwc":
        line["note"] = ("group-wise correlation has no body in the reference (a commented-out call, adamvs.py:271,295): parity "
                        "for this mode is UNPINNED by the reference; the oracle restates it by analogy with adamvs.py:473")
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(wl)
    return line


# ------------------------------------------------------------------------------------- AdaMVS cascade (cfg3)
CASCADE = [  # stage: scale, C, D, interval ratio, regress upsampling (adamvs.py:565-617, in_up per stage)
    (4, 32, 48, 4, 2),
    (2, 16, 32, 2, 2),
    (1, 8, 8, 1, 1),
]


def run_cascade(args):
    """BASELINE.json config 3: the hot path of one AdaMVS reference view at 1856x2752 -- per stage the pair
    volumes (stage 1), their softmax confidences, the view-weighted product volume in plane-major layout,
    the plane-at-a-time streaming soft-argmax on the (2x upsampled) regulariser output, and the hypothesis
    resampling between stages -- through the same `deep3d_aerial_b200.sweep` calls the drop-in
    `InferDepthNet.forward` makes (depthnets.py).  The CNN regularisers are out of scope on both arms: their
    outputs are synthetic tensors resident in HBM."""
    import torch
    import torch.nn.functional as F

    from deep3d_aerial_b200 import _lib, depthnets, shard, sweep, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the sweep engine has no CPU path")
    _lib.load()
    torch.set_grad_enabled(False)
    rank, world, local = shard.init()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    batch_planes = int(os.environ.get("D3D_STREAM_BATCH_PLANES", depthnets.STREAM_BATCH_PLANES))
    v, full_h, full_w, num_depth = 5, 2752, 1856, 384
    rig = synth.make_rig(num_views=v)
    base_interval = (rig.dmax - rig.dmin) / num_depth
    g = torch.Generator(device="cpu").manual_seed(7 + rank)
    stages = []
    for scale, c, d, ratio, up in CASCADE:
        h, w = full_h // scale, full_w // scale
        feats = torch.randn(v, c, h, w, generator=g, dtype=torch.float32).to(dev)
        proj = torch.from_numpy(rig.proj(scale)).to(dev)
        logits = (4.0 * torch.randn(d, h * up, w * up, generator=g, dtype=torch.float32)).clamp_(-20, 20).to(dev)
        stages.append({"c": c, "d": d, "h": h, "w": w, "up": up, "ratio": ratio, "feats": feats,
                       "pose": sweep.relative_poses(proj), "logits": logits,
                       "state": torch.zeros((3, h * up, w * up), device=dev)})
    s1 = stages[0]
    pair_logits = (4.0 * torch.randn(v - 1, s1["d"], s1["h"], s1["w"], generator=g, dtype=torch.float32)).to(dev)
    vox = sum(s["d"] * s["h"] * s["w"] for s in stages)
    # algorithmic bytes: volumes written once, features read once, logits read once, hypotheses
    bytes_stage = []
    for i, s in enumerate(stages):
        n = s["d"] * s["h"] * s["w"]
        b = 4 * s["c"] * n + 4 * v * s["c"] * s["h"] * s["w"] + 4 * n + 4 * s["d"] * s["h"] * s["up"] * s["w"] * s["up"]
        if i == 0:
            b += 2 * 4 * (v - 1) * n            # pair volumes written, pair logits read
        bytes_stage.append(b)
    ev = {}

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        ev.setdefault(name, []).append(e)

    # Block walk (FeatureCache + texel_pool.TexelPool, what predict.py --feature_cache runs): the views
    # of a scene block share 4 of their 5 images, whose texels stay resident in a pool of V + 1 slots per stage; a view
    # lays out only the image new to the walk and names its views by slot.  `walk` = None: every view lays out all V maps
    # into a dense block (no cache: what a lone reference view costs).
    for s in stages:
        s["pool"] = torch.empty((v + 1, s["h"], s["w"], s["c"]), device=dev)
        sweep.to_texels(s["feats"], out=s["pool"][:v])
        s["pool"][v].copy_(s["pool"][0])

    def step(walk=None, lay_out=True):
        depth = None
        conf = None
        slots = None if walk is None else [(walk - j) % (v + 1) for j in range(v)]   # newest image first (the reference)
        for i, s in enumerate(stages):
            mark("s%d_begin" % (i + 1))
            if walk is None:
                tex = sweep.to_texels(s["feats"])
            else:
                tex = s["pool"]
                if lay_out:
                    sweep.to_texels([s["feats"][walk % v]], out=tex[walk % (v + 1):walk % (v + 1) + 1])
            rays = sweep.rays_for(s["pose"], s["h"], s["w"])            # once per stage (depthnets._scene)
            if depth is None:
                hyps = sweep.depth_samples(sweep.SAMPLES_RANGE, s["d"], (s["h"], s["w"]), device=dev,
                                           dmin=rig.dmin, dmax=rig.dmax)
                pairs = sweep.cost_volume(tex, s["pose"], hyps, sweep.AGG_PAIR_MEAN, rays=rays, view_slots=slots)
                conf = torch.stack([sweep.depth_regress(pair_logits[k], hyps, want_index=False)["conf"]
                                    for k in range(v - 1)], 0)
                del pairs
            else:
                hyps = sweep.depth_samples(sweep.SAMPLES_AROUND, s["d"], (s["h"], s["w"]), cur=depth,
                                           interval=s["ratio"] * base_interval)
                conf = sweep.resize_bilinear(conf, (s["h"], s["w"]))       # adamvs.py:498-503, one launch for the V-1 maps
            mark("s%d_sweep" % (i + 1))
            sim = sweep.cost_volume(tex, s["pose"], hyps, sweep.AGG_WEIGHTED_PRODUCT, weights=conf.contiguous(),
                                    plane_major=True, rays=rays, view_slots=slots)
            mark("s%d_regress" % (i + 1))
            r = None
            held = []                          # planes arrive one at a time, as the GRU regulariser delivers them;
            for k in range(s["d"]):            # depthnets._Stream folds STREAM_BATCH_PLANES of them in per launch
                held.append(s["logits"][k])
                if len(held) == batch_planes or k == s["d"] - 1:
                    r = sweep.depth_regress(held, hyps, softmax_mode=sweep.SOFTMAX_RAW_EXP, d_begin=k + 1 - len(held),
                                            num_depth=s["d"], state=s["state"], finalize=(k == s["d"] - 1))
                    held = []
            depth = r["depth"]
            del sim
            mark("s%d_end" % (i + 1))
        return depth, r["conf"]

    # the dense-block form first (every view lays out all V maps), then the timed block walk
    for _ in range(args.warmup):
        step()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    d0.record()
    for _ in range(args.steps):
        step()
    d1.record()
    torch.cuda.synchronize()
    ms_dense = shard.join_max(d0.elapsed_time(d1)) / args.steps
    for k in range(args.warmup):
        step(walk=k)
    ev.clear()
    sampler = ClockSampler(local)
    sampler.start()
    shard.barrier()
    torch.cuda.synchronize()
    launches0 = _lib.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(args.steps):
        out = step(walk=args.warmup + k)
    t1.record()
    torch.cuda.synchronize()
    shard.barrier()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop()
    ms_total = shard.join_max(t0.elapsed_time(t1))
    ms_step = ms_total / args.steps
    value = shard.join_sum(vox * args.steps) / (ms_total * 1e-3) / 1e9

    def span(a, b):
        return statistics.mean(x.elapsed_time(y) for x, y in zip(ev[a], ev[b]))

    kernel_ms = {}
    for i in range(3):
        n = "s%d" % (i + 1)
        kernel_ms[n + "_prepare"] = span(n + "_begin", n + "_sweep")
        kernel_ms[n + "_weighted_product"] = span(n + "_sweep", n + "_regress")
        kernel_ms[n + "_streaming_regress"] = span(n + "_regress", n + "_end")
    peak, peak_src = measured_peak()
    dom = max(range(3), key=lambda i: kernel_ms["s%d_weighted_product" % (i + 1)])
    s = stages[dom]
    dom_bytes = 4 * s["c"] * s["d"] * s["h"] * s["w"] + 4 * v * s["c"] * s["h"] * s["w"] + 4 * s["d"] * s["h"] * s["w"]
    achieved = dom_bytes / (kernel_ms["s%d_weighted_product" % (dom + 1)] * 1e-3) / 1e9
    total_launches = int(shard.join_sum(launches))

    # ---- end to end: the reference views of a scene block in order; a view brings ONE image the device has not seen (the
    # other four are the previous view's), so per view the feature pyramid of one image (three maps, 286 MB) goes
    # host->device from pinned memory into a ring of V + 1 image slots -- on a copy stream, under the previous view's
    # sweeps -- and the full-resolution depth + confidence maps come back device->host.
    e2e = None
    if not args.no_e2e:
        n_e2e = max(3, args.steps)
        pyramid = [torch.randn(s["c"], s["h"], s["w"], generator=g, dtype=torch.float32).pin_memory() for s in stages]
        staging = [torch.empty((s["c"], s["h"], s["w"]), device=dev) for s in stages]   # the uploaded maps, before their relayout
        out_host = [torch.empty((2, full_h, full_w), dtype=torch.float32).pin_memory() for _ in range(2)]
        copy_stream, down_stream = torch.cuda.Stream(), torch.cuda.Stream()
        arrived, consumed = torch.cuda.Event(), torch.cuda.Event()
        computed, landed = torch.cuda.Event(), [torch.cuda.Event(), torch.cuda.Event()]
        consumed.record()
        h2d = sum(t.numel() * 4 for t in pyramid)
        sink = 0.0

        def collect(i):
            """The host's read of view i's maps (its device->host copy was queued a view ago)."""
            nonlocal sink
            landed[i & 1].synchronize()
            sink += float(out_host[i & 1][0, 0, 0]) + float(out_host[i & 1][1, 0, 0])

        def view(i):
            """View i reads texel-pool slots i, i-1 .. i-V+1 (mod V+1); the image of slot i+1 (the next view's new one) is
            uploaded AND laid out now, on the copy stream, under this view's sweeps; its maps go device->host on a third
            stream under the NEXT view's sweeps, and the host reads view i-1's maps meanwhile: nothing in the loop waits
            for the view it has just queued."""
            slot = (i + 1) % (v + 1)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed)                   # the view that last read this slot is done with it
                for k, s in enumerate(stages):
                    staging[k].copy_(pyramid[k], non_blocking=True)
                    sweep.to_texels([staging[k]], out=s["pool"][slot:slot + 1])
                arrived.record()
            dep, conf = step(walk=i, lay_out=False)
            consumed.record()
            computed.record()
            torch.cuda.current_stream().wait_event(arrived)        # the next view needs the image that just arrived
            with torch.cuda.stream(down_stream):
                down_stream.wait_event(computed)
                out_host[i & 1][0].copy_(dep, non_blocking=True)
                out_host[i & 1][1].copy_(conf, non_blocking=True)
                landed[i & 1].record()
            dep.record_stream(down_stream)
            conf.record_stream(down_stream)
            if i > 0:
                collect(i - 1)

        view(0)
        collect(0)
        shard.barrier()
        torch.cuda.synchronize()
        t_host = time.perf_counter()
        for i in range(1, 1 + n_e2e):
            view(i)
        collect(n_e2e)
        torch.cuda.synchronize()
        dt = shard.join_max(time.perf_counter() - t_host)
        if not math.isfinite(sink):
            raise SystemExit("bench.py: the cascade produced non-finite maps")
        e2e = {"value": shard.join_sum(vox * n_e2e) / dt / 1e9, "unit": "Gvoxel/s", "ms_per_step": dt / n_e2e * 1e3,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": out_host[0].numel() * 4, "steps": n_e2e,
               "timer": "host wall clock", "stream": "one new image (feature pyramid of 3 maps) per reference view, "
               "uploaded and laid out under the view's sweeps into texel pools of V + 1 resident images; the maps of view i "
               "come back under view i+1"}
    if rank != 0:
        return None
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import sweep_torch
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        s = stages[0]
        feats = [s["feats"][i:i + 1].cpu() for i in range(v)]
        proj = torch.from_numpy(rig.proj(4)).unsqueeze(0)
        planes = 2
        hy = torch.linspace(rig.dmin, rig.dmax, s["d"])[:planes].view(1, planes, 1, 1).repeat(1, 1, s["h"], s["w"])
        wts = torch.rand(1, v - 1, s["h"], s["w"])
        t_cpu = time.perf_counter()
        sweep_torch.pair_mean_volumes(feats, proj, hy)
        sweep_torch.weighted_product_volume(feats, proj, hy, [wts[:, i:i + 1] for i in range(v - 1)])
        dt_cpu = time.perf_counter() - t_cpu
        cpu = {"value": planes * s["h"] * s["w"] / dt_cpu / 1e9, "unit": "Gvoxel/s", "cores": cores, "kind": "port",
               "sample": "stage 1 only (C=32 at 688x464): pair volumes + weighted product of %d of %d planes, torch %s CPU"
                         % (planes, s["d"], torch.__version__)}
    line = {
        "metric": "cost-volume Gvoxels/s", "value": value, "unit": "Gvoxel/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "AdaMVS 3-stage cascade V=5 at 1856x2752: C/D = 32/48 @1/4, 16/32 @1/2, 8/8 @full; "
                               "pair volumes + weighted product + streaming soft-argmax + resampling",
                   "voxels_per_view": vox, "algorithmic_bytes_per_view": sum(bytes_stage),
                   "l2": "every volume (1.3-2.6 GB) exceeds the 126 MB L2", "views_per_step_per_gpu": 1,
                   "stream_batch_planes": batch_planes,
                   "views": "a block walk: the view's five images live as texels in a pool of V + 1 slots per stage and only "
                            "the image new to the walk is laid out (FeatureCache + TexelPool, predict.py --feature_cache); "
                            "ms_per_step_dense_block is the same view laying out all five maps (no cache)"},
        "ms_per_step_dense_block": ms_dense,
        "ref_views_per_s": world * 1e3 / ms_step, "kernel_ms": kernel_ms,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "weighted-product sweep, stage %d" % (dom + 1),
                     "algorithmic_bytes": dom_bytes, "peak_source": peak_src,
                     "whole_view_frac": sum(bytes_stage) / (ms_step * 1e-3) / 1e9 / peak},
        "clocks": clocks, "gpu_launches": total_launches,
        "e2e": e2e, "cpu_baseline": cpu,
        "note": "parity-test configuration measured for SURVEY.md 8d; the headline line is --workload cfg2",
    }
    float(out[0][0, 0])
    return line


# --------------------------------------------------------------------------- scene block across ranks (cfg5)
def run_scene_block(args):
    """BASELINE.json config 5: ONE scene block of --block-views reference views (cfg2 shape each) dealt to the ranks
    with shard.partition(..., "contiguous") -- neighbouring views, which share 4 of their 5 images, stay on one GPU --
    and run end to end from pinned host memory through pipeline.ViewPipeline (reference: the serial loop of
    mvs/mvs_cas/predict.py:126-133 over a block of IO/params_io.py:430-444).  Strong scaling: the block is fixed,
    every rank's pipeline starts cold (its first view uploads all V images) and the job ends with its slowest rank."""
    import torch

    from deep3d_aerial_b200 import _lib, shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the sweep engine has no CPU path")
    _lib.load()
    torch.set_grad_enabled(False)
    rank, world, local = shard.init()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    wl = "cfg2"
    v, c, d, h, w, mode, groups, desc = WORKLOADS[wl]
    vox, b1, b2, brel = algorithmic_bytes(wl)
    n_block = args.block_views
    mine = shard.partition(list(range(n_block)), world, rank, "contiguous")
    slots = [make_inputs(wl, 1000 * rank + s, dev) for s in range(2)]
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _lib.launch_count()
    rec = e2e_block(args, wl, dev, rank, slots, mine, n_block, warm_views=max(3, args.warmup))
    clocks = sampler.stop()
    launches = int(shard.join_sum(_lib.launch_count() - launches0))
    seconds = shard.join_max(rec.pop("seconds"))
    h2d = shard.join_sum(rec["h2d_bytes_per_step"] * len(mine)) / n_block
    if rank != 0:
        return None
    value = n_block * vox / seconds / 1e9
    peak, peak_src = measured_peak()
    line = {
        "metric": "cost-volume Gvoxels/s", "value": value, "unit": "Gvoxel/s", "n_gpus": world, "steps": n_block,
        "warmup": max(3, args.warmup), "ms_per_step": seconds / n_block * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "scene block of %d reference views, each %s; dealt contiguously to %d rank(s), end to end "
                               "from pinned host memory" % (n_block, desc, world),
                   "views_per_rank": len(mine), "sharding": "reference views, no collective",
                   "l2": "every view's volume (15.69 GB) exceeds the 126 MB L2"},
        "ref_views_per_s": n_block / seconds,
        "roofline": {"bound": "hbm", "achieved": n_block * (b1 + b2) / seconds / 1e9 / world, "peak": peak, "unit": "GB/s",
                     "frac": n_block * (b1 + b2) / seconds / 1e9 / world / peak, "traffic": None,
                     "kernel": "whole view end to end (sweep + regression), per GPU", "algorithmic_bytes": b1 + b2,
                     "peak_source": peak_src},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": dict({"value": value, "unit": "Gvoxel/s"}, **dict(rec, h2d_bytes_per_step=int(h2d))),
        "cpu_baseline": None,
        "note": "BASELINE.json config 5 (views/s of a fixed block); the headline line is the default --workload cfg2",
    }
    return line


# ------------------------------------------------------------------------------------------------ row f3
def run_fuse(args):
    """the depth-map fusion consistency check of one reference view against its 10 source
    views (fuse/fusion_3d_normal.py --fusion_num) at the full 1856 x 2752 depth-map size, one
    `d3d_consistency_fuse` launch per step.  `value`: pixel pairs (reference pixel x source view) per second with
    every map resident in HBM; `e2e`: the same from pinned host maps (reference maps + geometry up, count / fused
    points / final mask / filtered depth down; the source maps stay resident, as they do across the reference
    views of a scene block); cpu_baseline: the numpy oracle (= the reference's CuPy code on numpy) on a row band."""
    import numpy as np
    import torch

    from deep3d_aerial_b200 import _lib, fusion, shard, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the fusion kernel has no CPU path")
    _lib.load()
    torch.set_grad_enabled(False)
    rank, world, local = shard.init()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    S, H, W = 10, 2752, 1856
    sc = synth.fusion_scene(num_src=S, height=H, width=W, seed=11 + rank, focal=4000.0)
    d, n, k, e, prob = sc["ref"]
    th = dict(position_threshold=1.0, depth_threshold=0.01, normal_threshold_cos=math.cos(math.radians(10.0)),
              confidence_threshold=0.2, min_consistent=4)
    geom_host = fusion.pair_geometry(k, e, [v[2] for v in sc["src"]], [v[3] for v in sc["src"]])
    up = lambda x: torch.from_numpy(x).to(dev)                                   # noqa: E731
    ref_dev = (up(d), up(n), up(prob))
    geom = up(geom_host)
    src_d, src_n = [up(v[0]) for v in sc["src"]], [up(v[1]) for v in sc["src"]]
    out = fusion.fuse_view(*ref_dev, geom, src_d, src_n, **th)
    pairs = H * W * S
    # algorithmic bytes per step: reference maps read once (depth 4 + normal 12 + prob 4), per source view the
    # gathered depth + normal (16) and the mask (1), the copy of every source map (4 + 4), and the fused outputs
    # (count 4 + point 12 + final mask 1 + filtered depth 4)
    alg_bytes = H * W * (20 + S * (16 + 1 + 8) + 21)

    def step():
        fusion.fuse_view(*ref_dev, geom, src_d, src_n, out=out, **th)

    for _ in range(max(3, args.warmup)):
        step()
    sampler = ClockSampler(local)
    sampler.start()
    shard.barrier()
    torch.cuda.synchronize()
    launches0 = _lib.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step()
    t1.record()
    torch.cuda.synchronize()
    shard.barrier()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop()
    ms_total = shard.join_max(t0.elapsed_time(t1))
    ms_step = ms_total / args.steps
    value = shard.join_sum(pairs * args.steps) / (ms_total * 1e-3) / 1e9

    e2e = None
    if not args.no_e2e:
        pin = [torch.from_numpy(x).pin_memory() for x in (d, n, prob, geom_host)]
        pipe = fusion.FusionPipeline(H, W, S, dev, **th)

        def run(nviews):                                # every view: its maps up, its results down (overlapped)
            got = None
            for i in range(nviews):
                pipe.submit(pin[0], pin[1], pin[2], pin[3], src_d, src_n)
                if i >= 1:
                    got = pipe.collect()
            for got in pipe.drain():
                pass
            return got

        host_out = run(3)
        shard.barrier()
        torch.cuda.synchronize()
        t_host = time.perf_counter()
        host_out = run(args.steps)
        torch.cuda.synchronize()
        ms_e2e = shard.join_max((time.perf_counter() - t_host) * 1e3)
        shard.barrier()
        assert torch.equal(host_out["count"], out["count"].cpu())
        e2e = {"value": shard.join_sum(pairs * args.steps) / (ms_e2e * 1e-3) / 1e9, "unit": "Gpixel-pair/s",
               "ms_per_step": ms_e2e / args.steps,
               "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in pin),
               "d2h_bytes_per_step": sum(t.numel() * t.element_size() for t in host_out.values()),
               "timer": "host wall clock around submit/collect (fusion.FusionPipeline: copies overlap the kernel)"}
    total_launches = int(shard.join_sum(launches))
    if rank != 0:
        return None
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import fuse_np
        rows = 256
        kw = dict(position_threshold=1.0, depth_threshold=0.01, normal_cos=th["normal_threshold_cos"], confidence_threshold=0.2)
        t = time.perf_counter()
        fuse_np.fuse_view(d[:rows], n[:rows], k, e, prob[:rows], sc["src"], min_consistent=4, **kw)
        dt = time.perf_counter() - t
        cpu = {"value": rows * W * S / dt / 1e9, "unit": "Gpixel-pair/s", "cores": 1, "kind": "port",
               "sample": "the first %d of %d rows of the reference view against all %d source views, numpy %s"
                         % (rows, H, S, np.__version__)}
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (ms_step * 1e-3) / 1e9
    line = {
        "metric": "consistency-checked pixel pairs/s", "value": value, "unit": "Gpixel-pair/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "depth-map fusion consistency check, 1 reference view x %d source views at %dx%d "
                               "(fuse/consistency_check_n.py)" % (S, W, H),
                   "l2": "maps of one step (1.3 GB) exceed the 126 MB L2", "views_per_step_per_gpu": 1,
                   "final_mask_fraction": float(out["final_mask"].float().mean())},
        "ref_views_per_s": world * 1e3 / ms_step,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "consistency_fuse_kernel", "algorithmic_bytes": alg_bytes,
                     "peak_source": peak_src},
        "clocks": clocks, "gpu_launches": total_launches, "e2e": e2e, "cpu_baseline": cpu,
        "note": "row f3 of SURVEY.md 8f; the headline line is --workload cfg2",
    }
    return line


def _brief(line):
    """What the headline line keeps of another workload's line."""
    keep = ("value", "unit", "ms_per_step", "steps", "warmup", "scaling", "ref_views_per_s", "kernel_ms", "roofline", "e2e",
            "cpu_baseline", "gpu_launches", "note", "ms_per_step_dense_block")
    out = {k: line[k] for k in keep if k in line}
    out["workload"] = line["config"]["workload"]
    if "views" in line["config"]:
        out["views"] = line["config"]["views"]
    return out


def main():
    args = parse()
    if args.impl == "reference":
        if args.workload in ("cfg3", "cfg5", "fuse"):
            args.workload = "cfg2"     # the reference arm is quoted on the headline configuration
        return run_reference(args)
    if args.workload == "cfg3":
        line = run_cascade(args)
    elif args.workload == "cfg5":
        line = run_scene_block(args)
    elif args.workload == "fuse":
        line = run_fuse(args)
    else:
        line = run_ours(args)
        # the default run also records the other configurations of BASELINE.json, briefly, next to the headline (N = 1:
        # the scaling runs are about the headline alone)
        if line is not None and args.workload == "cfg2" and args.gpus == 1 and not args.no_extra and \
                int(os.environ.get("WORLD_SIZE", "1")) == 1:
            import copy
            import torch
            extra = {}
            for name, fn, over in (("cfg4", run_ours, dict(workload="cfg4", steps=8, warmup=3, no_e2e=True)),
                                   ("cfg3", run_cascade, dict(steps=10, warmup=3)),
                                   ("cfg5", run_scene_block, dict(block_views=16, warmup=3))):
                sub = copy.copy(args)
                for k, val in over.items():
                    setattr(sub, k, val)
                sub.no_cpu_baseline = True if name != "cfg3" else args.no_cpu_baseline
                torch.cuda.empty_cache()
                try:
                    extra[name] = _brief(fn(sub))
                except Exception as exc:  # noqa: BLE001 -- an extra never costs the headline its line
                    extra[name] = {"error": "%s: %s" % (type(exc).__name__, exc)}
            line["extra_workloads"] = extra
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    rc = main()
    try:
        from deep3d_aerial_b200 import shard as _shard
        _shard.finalize()
    except Exception:  # noqa: BLE001 -- teardown only
        pass
    sys.exit(rc)
