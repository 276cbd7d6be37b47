#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 path tracer.

Metric (BASELINE.json): Mpath-segments/s (and s/frame) of the offline path-traced render
  workload C2: procedural Sponza-class atrium (~300 k triangles, 24 textured PBR materials, HDR
  environment), 1920x1080, batch size 16, depth 9; a frame is 1024 spp = 64 batches.
One "step" = one batch (16 samples per pixel over the whole image) of that render.

  python bench.py --gpus N --steps K --warmup W            # the CUDA product
  python bench.py --impl reference ...                      # the reference estimator on the host cores (CPU oracle)

N > 1: one process per GPU under torchrun.  Every rank owns one engine / one context of the product; the contexts join ONE
NCCL communicator inside the product (ptc_comm_unique_id / ptc_comm_init_rank - torch.distributed only carries the 128-byte id,
the barrier and the max-over-ranks of the timings).  The scene is replicated, the batches are dealt round-robin to the ranks
(sample split, weak scaling: K batches per rank) and the product sums the three accumulation buffers onto rank 0 with one
ncclReduce inside the timed region.  `e2e` goes through RendererPathTracing::render() with host buffers on every rank, and
`frame` is the fixed 1024-spp frame (strong scaling: the frame's 64 batches split over the ranks), measured, not extrapolated.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from oracle import loader as oracle_loader  # noqa: E402  (cpu_baseline leg and --impl reference only)

FRAME_SPP = 1024
SHADING_RECORD_BYTES = 144  # vviewer_b200/csrc/lbvh.cuh::k_gather_shading


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--scene", default="Atrium")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--depth", type=int, default=0)
    ap.add_argument("--texsize", type=int, default=1024)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-frame", action="store_true", help="skip the measured 1024-spp frame (strong-scaling figure)")
    ap.add_argument("--split", default="sample", choices=["sample", "tile"], help="how N > 1 GPUs partition the image")
    ap.add_argument("--cpu-spp", type=int, default=1, help="samples per pixel of one step of the reference arm (--impl reference)")
    ap.add_argument("--cpu-baseline-spp", type=int, default=12, help="samples per pixel of the cpu_baseline leg of the CUDA arm (about 10-15 s of CPU work)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes_per_ray(n_tris):
    """k_extend: one ideal root-to-leaf descent of the 8-wide compressed BVH:
    D = ceil(log8(N / 3)) nodes of 80 B + one leaf of 3 triangles of 48 B + the ray's state records."""
    import math
    depth = max(1, math.ceil(math.log(max(n_tris / 3.0, 2.0), 8)))
    state = 4 + 16 + 16 + 16  # queue id, origin+rng, direction+flags read; hit record written
    return state + depth * 80 + 3 * 48, depth


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def workload_name(args, ri):
    return "C2 %s %dx%d batch %d depth %d (frame = %d spp)" % (args.scene, ri["width"], ri["height"], ri["batch_size"], ri["depth"], FRAME_SPP)


def build_engine(args):
    from vviewer_b200 import capi
    eng = capi.HostEngine()
    eng.build_scene(args.scene, texture_size=args.texsize, scale=args.scale)
    eng.set_render_info(width=args.width, height=args.height, batch_size=args.batch, depth=args.depth)
    return eng


def run_reference(args, rank, world):
    """Reference arm: the reference estimator (CPU oracle restating the GLSL shaders) on the host cores.
    The reference's own binary cannot run here (Vulkan RT pipeline, no lavapipe: BASELINE.md §2)."""
    if rank != 0:
        return
    from vviewer_b200 import capi
    eng = build_engine(args)  # the host feeder only: the scene description is handed to the oracle below
    ri = eng.render_info()
    oracle = oracle_loader.load_oracle()
    try:  # torchrun exports OMP_NUM_THREADS=1 to its workers: the reference arm uses ALL host cores whatever the environment says
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(os.cpu_count()))
    except OSError:
        pass
    ctx = capi.Context(oracle)
    ctx.upload_scene(eng.scene_desc())
    ctx.build_accel()
    rp = eng.render_params()
    spp = max(1, args.cpu_spp)
    rp.batch_size = spp
    seg = 0
    ms = 0.0
    for i in range(args.warmup + args.steps):
        rp.samples = spp
        ctx.render(rp, want_aovs=False)
        st = ctx.stats()
        if i >= args.warmup:
            seg += st["segments"]
            ms += st["render_ms"]
    cores = os.cpu_count()
    value = seg / ms / 1e3 if ms > 0 else 0.0
    sample = "%dx%d x %d spp per step (a bounded sample: a full step of this workload is %d spp)" % (ri["width"], ri["height"], spp, ri["batch_size"])
    line = {"impl": "reference", "metric": "Mpath-segments/s", "value": value, "unit": "Msegments/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args, ri), "triangles": ctx.stats()["n_triangles"]},
            "cpu_baseline": {"value": value, "unit": "Msegments/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Msegments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_cuda(args, rank, world, local_rank):
    import numpy as np
    import torch
    from vviewer_b200 import capi

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def share_id(make):
        """128 opaque bytes from rank 0 to everyone: the only thing torch.distributed carries for the product's communicator"""
        box = [make() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    cuda = capi.load_cuda()
    eng = build_engine(args)
    if not eng.backend_ok():
        raise RuntimeError("CUDA backend failed: " + eng.last_error())
    ri = eng.render_info()
    W, H, B, D = ri["width"], ri["height"], ri["batch_size"], ri["depth"]
    desc = eng.scene_desc()
    ctx = capi.Context(cuda, device=local_rank)
    if world > 1:
        ctx.comm_init_rank(share_id(ctx.unique_id), rank, world)   # device-timed leg
        eng.comm_init_rank(share_id(eng.comm_unique_id), rank, world)  # plugin leg
        eng.set_render_options(split=args.split)
    ctx.upload_scene(desc)
    ctx.build_accel()  # the first build of a process also loads the module and sizes its scratch buffers
    ctx.build_accel()
    build_stats = ctx.stats()
    split_mode = capi.PTC_SPLIT_TILE if args.split == "tile" else capi.PTC_SPLIT_SAMPLE

    def render(n_batches, flags=0):
        """n_batches per rank; with a communicator the product partitions the render and reduces onto rank 0 (device timed)"""
        rp = eng.render_params()
        rp.samples = n_batches * B * world
        rp.batch_size = B
        rp.flags = flags
        if world > 1:
            rp.split_mode = split_mode
        ctx.render_device(rp, None, None, None)
        return ctx.stats()

    def sync():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    render(max(args.warmup, 3))
    sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sync()
    t0 = time.perf_counter()
    st = render(args.steps)             # K batches on this rank + the product's ncclReduce, device-timed with CUDA events on the library's streams
    sync()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = st["render_ms"]            # max over this context's devices, reduce included
    clocks = sampler.stop() if rank == 0 else None

    def over_ranks(values):
        t = torch.tensor(values, dtype=torch.float64, device="cuda")
        if dist is None:
            return t.tolist(), t.tolist()
        tmax, tsum = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        return tmax.tolist(), tsum.tolist()

    mx, sm = over_ranks([dev_ms, float(st["segments"]), float(st["kernel_launches"]), wall_ms, st["reduce_ms"]])
    dev_ms, wall_ms, reduce_ms = mx[0], mx[3], mx[4]
    segments, launches = sm[1], int(sm[2])

    # ---- end to end through the public API (RendererPathTracing::render via the C++ plugin) on EVERY rank: host scene in, host
    # images out on rank 0; includes flatten, H2D upload of the scene, BVH build, render, the NCCL reduce and D2H of the 3 targets
    out = [np.empty(W * H * 4, np.float32) for _ in range(3)]  # the caller's host images, reused by every call

    def plugin_render(total_batches):
        eng.set_render_info(samples=total_batches * B)
        sync()
        t = time.perf_counter()
        eng.render_to_memory(out)
        dt = time.perf_counter() - t
        est = eng.stats()
        mxs, sms = over_ranks([dt, float(est["segments"])])
        return mxs[0], sms[1]

    plugin_render(world)  # warm-up of the plugin path (allocations, communicator)
    e2e_s, e2e_seg = plugin_render(args.steps * world)
    d = desc.contents
    # textures / environment are resident after the first call (ptc_texture.uid): not copied in the timed call
    h2d = d.n_vertices * 68 + d.n_indices * 4 + d.n_instances * 176 + d.n_materials * 128 + d.n_light_instances * 64 + d.n_light_data * 64
    d2h = 3 * W * H * 16
    e2e = {"value": e2e_seg / e2e_s / 1e6, "unit": "Msegments/s", "h2d_bytes_per_step": int(h2d / args.steps), "d2h_bytes_per_step": int(d2h / args.steps / world),
           "seconds": e2e_s,
           "note": "one RendererPathTracing::render() call per rank, %d batches per rank: flatten + scene upload (geometry, instances, materials; textures and "
                   "environment keep their device copies by identity, like the reference's import-time upload) + BVH build + render + NCCL reduce "
                   "inside the product + readback on rank 0; h2d per rank, d2h on rank 0 divided over the ranks" % args.steps}
    frame = None
    if not args.no_frame:
        f_s, f_seg = plugin_render(FRAME_SPP // B)  # the WHOLE frame, its batches split over the ranks: strong scaling, measured
        frame = {"spp": FRAME_SPP, "seconds": f_s, "Msegments_per_s": f_seg / f_s / 1e6, "split": args.split if world > 1 else "none",
                 "note": "fixed %dx%d x %d spp frame through RendererPathTracing::render() with host buffers (upload + build + render + reduce + readback)" % (W, H, FRAME_SPP)}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    value = segments / dev_ms / 1e3  # Msegments/s, whole job
    roofline, cpu = None, None
    if True:
        # ---- roofline of the dominant kernel (k_extend), timed live with CUDA events on the launching stream.  With several ranks this
        # is rank 0's GPU alone, on a context of its own without the communicator (the other ranks wait at the final barrier)
        if world == 1:
            stp = render(2, flags=capi.PTC_FLAG_TIME_KERNELS)
        else:
            rctx = capi.Context(cuda, device=local_rank)
            rctx.upload_scene(desc)
            rctx.build_accel()
            rp1 = eng.render_params()
            rp1.samples, rp1.batch_size, rp1.flags, rp1.split_mode = 2 * B, B, capi.PTC_FLAG_TIME_KERNELS, capi.PTC_SPLIT_NONE
            rctx.render_device(rp1, None, None, None)
            stp = rctx.stats()
            rctx.close()
        bytes_per_ray, bvh_depth = algorithmic_bytes_per_ray(build_stats["n_triangles"])
        peak, peak_kind = measured_peak()
        rays = stp["segments"]
        avg_launch_ms = stp["trace_ms"] / max(stp["trace_launches"], 1)
        achieved = rays * bytes_per_ray / (stp["trace_ms"] * 1e-3) / 1e9 if stp["trace_ms"] > 0 else 0.0
        share = {"extend": stp["trace_ms"], "shade": stp["shade_ms"], "shadow_probe": stp["shadow_ms"]}
        traffic, shade_traffic = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj.get("k_extend_dram_bytes_per_launch")
                shade_traffic = tj.get("k_shade_dram_bytes_per_launch")
            except Exception:
                traffic = None
        # second kernel (k_shade, DRAM bound): algorithmic bytes per segment with the actual record sizes - queue id 4,
        # hit + origin + direction + throughput read 64, new origin + direction + throughput written 48, shading record, two texture
        # taps of 16 B (albedo; normal + roughness of the bench scene's materials share one packed tap)
        shade_bytes = 4 + 64 + 48 + SHADING_RECORD_BYTES + 2 * 16
        shade_gbs = stp["segments"] * shade_bytes / (stp["shade_ms"] * 1e-3) / 1e9 if stp["shade_ms"] > 0 else 0.0
        roofline = {"bound": "hbm", "kernel": "k_extend", "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "bytes_per_ray": bytes_per_ray, "bvh_depth": bvh_depth,
                    "avg_launch_ms": avg_launch_ms, "kernel_ms_share": share,
                    "k_shade": {"achieved": shade_gbs, "frac": shade_gbs / peak, "bytes_per_segment": shade_bytes, "traffic": shade_traffic}}

        # ---- CPU baseline: the oracle on the host cores, bounded sample of the same workload
        if world == 1 and not args.no_cpu_baseline:
            octx = capi.Context(oracle_loader.load_oracle())
            octx.upload_scene(desc)
            octx.build_accel()
            rp = eng.render_params()
            rp.samples = rp.batch_size = max(1, args.cpu_baseline_spp)
            octx.render(rp, want_aovs=False)
            ost = octx.stats()
            cpu = {"value": ost["segments"] / ost["render_ms"] / 1e3, "unit": "Msegments/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": "%dx%d x %d spp (%d/%d of a step), %.1f s of all host cores (OpenMP over image rows)" % (W, H, rp.samples, rp.samples, B, ost["render_ms"] / 1e3)}
            octx.close()

    ms_per_step = dev_ms / args.steps
    line = {"metric": "Mpath-segments/s", "value": value, "unit": "Msegments/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args, ri),
                       "triangles": build_stats["n_triangles"], "l2": "inputs larger than L2 (%.1f GB of path state streamed per step)" % (W * H * B * 192 / 1e9),
                       "split": args.split if world > 1 else "none"},
            "s_per_frame": frame["seconds"] if frame else None, "frame": frame, "segments_per_path": segments / (args.steps * world * W * H * B),
            "build_ms": build_stats["build_ms"], "reduce_ms": reduce_ms, "wall_ms": wall_ms, "clocks": clocks, "gpu_launches": launches, "roofline": roofline,
            "cpu_baseline": cpu, "e2e": e2e}
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm is the CPU implementation on ALL host cores
        # (set before the OpenMP runtime of the oracle library starts)
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())
        run_reference(args, rank, world)
    else:
        run_cuda(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
