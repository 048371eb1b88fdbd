"""bench.py — surgery-render frames/s (512^2, 100k Gaussians) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path (FLAME blendshapes + skinning -> triangle-bound Gaussian
transform -> tile binning -> compositing) over one 300-frame, 512x512, 100k-Gaussian clip per rank
(BASELINE.json configs[2]; weak scaling: every rank renders its own clip, e.g. its own surgical
plan, and rank 0 gathers the finished uint8 frames over NCCL).  Prints ONE JSON line on rank 0.

  value     frames/s, whole job, inputs resident in HBM, CUDA-event time of K steps, max over ranks
  e2e       the same metric through the C-ABI session call with HOST buffers (pinned): parameters
            go host->device and uint8 frames come device->host inside the timed region
  roofline  the dominant kernel's achieved algorithmic GB/s (SURVEY.md §8d bytes x measured work,
            divided by its CUDA-event time from a separate profiling pass) against MEASURED_PEAKS.json
  cpu_baseline  the C oracle (oracle/, a port — the reference has no CPU renderer) on the host cores,
            bounded sample of the same workload

--impl reference times the oracle port alone (the reference arm of this tier).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WIDTH = HEIGHT = 512
N_GAUSS = 100_000
N_FRAMES = 300
METRIC = "surgery-render frames/s (512^2, 100k Gaussians)"
UNIT = "frames/s"


def workload_name(n_frames):
    return (f"configs[2]: {WIDTH}x{HEIGHT} x {n_frames}-frame surgery video, {N_GAUSS // 1000}k FLAME-bound "
            f"Gaussians, one clip per rank")


def make_inputs(n_frames, seed=0):
    import omfs_b200  # noqa: F401
    from omfs_b200 import avatar, synthetic
    model, params, av, cam = synthetic.make_scene(n_gauss=N_GAUSS, n_frames=n_frames, width=WIDTH, height=HEIGHT,
                                                  seed=seed)
    return model, params, avatar.bake(av), cam


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))), "measured"
    return 6650.0, 1590.0, "fallback"


def bind_to_gpu_numa_node(gpu_index):
    """Pin this rank's threads to the CPUs NVML reports as local to its GPU, BEFORE any pinned host buffer
    is allocated (first touch then places the buffers on that NUMA node): with one process per GPU the
    device->host frame traffic of 8 ranks otherwise crosses the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {w * 64 + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus ({min(cpus)}-{max(cpus)})"
    except Exception as e:  # affinity is an optimisation, never a failure
        return f"unbound ({type(e).__name__})"
    return "unbound"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_fps(model, params, baked, cam, n_sample):
    """Frames/s of the oracle port on the host cores for the first n_sample frames."""
    import oracle
    chunk = 20  # frames per oracle call: bounds the host memory of the pair lists
    t0 = time.perf_counter()
    for lo in range(0, n_sample, chunk):
        hi = min(n_sample, lo + chunk)
        oracle.render(model, params.slice(lo, hi), baked, [cam.pack()] * (hi - lo), WIDTH, HEIGHT)
    dt = time.perf_counter() - t0
    return n_sample / dt, dt, oracle.num_threads()


def run_reference(args, rank):
    """The reference arm: the oracle port (the reference has no CPU renderer of its own; its in-tree
    pieces of the path are the scalar parameter edits) on all host cores, bounded sample per step."""
    if rank != 0:
        return 0
    import oracle
    n_sample = args.ref_frames
    model, params, baked, cam = make_inputs(max(n_sample, 1))
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_fps(model, params, baked, cam, min(4, n_sample))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_fps(model, params, baked, cam, n_sample)
    dt = time.perf_counter() - t0
    fps = args.steps * n_sample / dt
    cores = oracle.num_threads()
    sample = f"{n_sample} of {N_FRAMES} frames per step ({WIDTH}x{HEIGHT}, {N_GAUSS} Gaussians), {cores} OpenMP threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(N_FRAMES), "sample": sample},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--frames", type=int, default=N_FRAMES)
    ap.add_argument("--batch", type=int, default=60, help="segments (frames) per launch group")
    ap.add_argument("--gemm", type=int, default=0, help="0 tensor-core blendshape GEMM, 1 CUDA-core")
    ap.add_argument("--cpu-frames", type=int, default=300, help="frames of the cpu_baseline sample")
    ap.add_argument("--ref-frames", type=int, default=60, help="frames per step of --impl reference")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="diagnosis only: skip the frame gather (flagged in config)")
    ap.add_argument("--gather", default="p2p", choices=("p2p", "nccl"),
                    help="frame exchange: copy-engine peer-to-peer pushes into rank 0 (default) or one NCCL gather")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch
    import torch.distributed as dist
    import omfs_b200  # noqa: F401
    from omfs_b200 import runtime

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    runtime.check(runtime.load_library().omfs_device_check(local_rank))

    T = args.frames
    # every rank renders its own clip (different expression/pose track, same subject and avatar)
    model, params, baked, cam = make_inputs(T, seed=0)
    if rank > 0:
        from omfs_b200 import synthetic
        params = synthetic.make_frame_params(T, seed=99 + rank)
        params.shape[:] = synthetic.make_frame_params(1, seed=99).shape
    sess = runtime.Session(model, baked, WIDTH, HEIGHT, max_batch=args.batch, device=local_rank, gemm_impl=args.gemm)
    sess.set_subject(params.shape, params.static_offset)
    hw = WIDTH * HEIGHT

    # ---- resident inputs (torch owns the device memory; the C-ABI sees raw pointers)
    def dev_t(a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)

    d_in = {k: dev_t(getattr(params, k)) for k in ("expr", "rotation", "neck_pose", "jaw_pose", "eyes_pose", "translation")}
    d_cam = dev_t(cam.pack()[None])
    d_ptrs = {k: v.data_ptr() for k, v in d_in.items()}
    d_ptrs["cams"] = d_cam.data_ptr()
    # finished frames: double-buffered so that the gather of step i overlaps the rendering of step i+1
    frames_bufs = [torch.empty((T, HEIGHT, WIDTH, 3), dtype=torch.uint8, device=dev) for _ in range(2 if world > 1 else 1)]
    frames_u8 = frames_bufs[0]
    gathered = None
    peer_gather = None
    slot_bytes = T * HEIGHT * WIDTH * 3
    if world > 1 and args.gather == "nccl":
        gathered = torch.empty((world, T, HEIGHT, WIDTH, 3), dtype=torch.uint8, device=dev) if rank == 0 else None
    elif world > 1:
        from omfs_b200 import sharding

        def exchange(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out

        peer_gather = sharding.PeerFrameGather(slot_bytes, rank, world, exchange)
    # a dedicated stream: the kernels and the timing events go through it; the NCCL gather runs on a second
    # stream, ordered by events (rendered -> gather may start; gathered -> the buffer may be overwritten)
    stream = torch.cuda.Stream(device=dev)
    comm_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ev_rendered = [torch.cuda.Event() for _ in frames_bufs]
    ev_gathered = [torch.cuda.Event() for _ in frames_bufs]
    step_no = [0]

    def step_device():
        b = step_no[0] % len(frames_bufs)
        step_no[0] += 1
        if world > 1:
            stream.wait_event(ev_gathered[b])
        sess.render_device(d_ptrs, T, 1, d_out_u8=frames_bufs[b].data_ptr(), stream=stream.cuda_stream)
        if world > 1 and not args.no_gather:
            # the only collective of the path: finished frames to rank 0 over NVLink
            ev_rendered[b].record(stream)
            comm_stream.wait_event(ev_rendered[b])
            if peer_gather is not None:
                # every rank pushes its block into its slot of rank 0's buffer: sender-side copy engine
                peer_gather.push(frames_bufs[b].data_ptr(), slot_bytes, comm_stream.cuda_stream)
                ev_gathered[b].record(comm_stream)
            else:
                with torch.cuda.stream(comm_stream):
                    dist.gather(frames_bufs[b], list(gathered.unbind(0)) if rank == 0 else None, dst=0)
                    ev_gathered[b].record(comm_stream)

    def drain():
        # the timed region ends only when every step's frames have arrived on rank 0
        if world > 1:
            stream.wait_stream(comm_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    drain()
    barrier()
    sess.sync()
    launches0 = runtime.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    drain()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = runtime.launch_count() - launches0
    sess.sync()
    pairs_per_frame = sess.stats()["pairs"] / T
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    ms_ranks = [ms]
    if world > 1:
        all_ms = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(all_ms, torch.tensor([ms], dtype=torch.float64, device=dev))
        ms_ranks = [float(x.item()) for x in all_ms]
    value = world * T * args.steps / (ms_max / 1e3)

    # ---- end to end through the host-buffer C-ABI call (pinned host memory both ways)
    host_in = {k: runtime.PinnedArray(getattr(params, k).shape, np.float32) for k in
               ("expr", "rotation", "neck_pose", "jaw_pose", "eyes_pose", "translation")}
    for k, v in host_in.items():
        v.array[...] = getattr(params, k)
    host_out = runtime.PinnedArray((T, HEIGHT, WIDTH, 3), np.uint8)

    class HostParams:
        pass

    hp = HostParams()
    for k, v in host_in.items():
        setattr(hp, k, v.array)
    hp.dynamic_offset = None

    def step_host():
        sess.render_host(hp, [cam], want_u8=True, out_u8=host_out.array)

    e2e_steps = max(2, min(args.steps, 5))
    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * T * e2e_steps / float(t_e.item())
    h2d = sum(v.array.nbytes for v in host_in.values()) + 160
    d2h = host_out.array.nbytes
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-stage profile for the roofline (separate pass: costs a host sync per batch)
    sess.set_profiling(True)
    for _ in range(2):
        sess.render_device(d_ptrs, T, 1, d_out_u8=frames_u8.data_ptr(), stream=stream.cuda_stream)
        sess.sync()
    sess.set_profiling(False)
    for _ in range(3):
        sess.render_device(d_ptrs, T, 1, d_out_u8=frames_u8.data_ptr(), stream=stream.cuda_stream)
        sess.sync()
    sess.set_profiling(True)
    prof_reps = 3
    for _ in range(prof_reps):
        sess.render_device(d_ptrs, T, 1, d_out_u8=frames_u8.data_ptr(), stream=stream.cuda_stream)
        sess.sync()
    st = sess.stage_ms()
    sess.set_profiling(False)

    if rank == 0:
        hbm, tflops, peak_kind = peaks()
        R = pairs_per_frame
        d = sess.dims()
        sort_bits = runtime.load_library().omfs_binning_sort_bits(args.batch, WIDTH, HEIGHT)
        tiles = ((WIDTH + 15) // 16) * ((HEIGHT + 15) // 16)
        # algorithmic bytes per FRAME for each stage.  flame .. bind_preprocess, composite: SURVEY.md §8d.
        # The binning is charged what THIS design has to move: depth sort = one 4-byte
        # histogram read + 4 passes x 16 B per Gaussian; tile counts/ranges = 20 B per Gaussian + 16 B per
        # tile; emit+scatter = 24 B per Gaussian read + 4 B per pair written once.
        alg = {
            "flame": 4.0 * 3 * d["V"] * 2 + 44.0 * d["V"],                  # GEMM output + LBS stream
            "face_frames": 80.0 * d["F"],
            "bind_preprocess": 288.0 * N_GAUSS,
            "depth_sort": (4.0 + 16.0 * 4) * N_GAUSS,
            "tile_ranges": 20.0 * N_GAUSS + 16.0 * tiles,
            "emit_scatter": 24.0 * N_GAUSS + 4.0 * R,
            "composite": 40.0 * R + 12.0 * hw,
        }
        published_sort_bytes = (12.0 + 8.0 + 24.0 * ((sort_bits + 7) // 8) + 8.0) * R   # SURVEY U7+U8+U9
        stages = {}
        total_ms = sum(v["ms"] for v in st.values())
        for name, v in st.items():
            if not v["calls"]:
                continue
            per_launch_ms = v["ms"] / v["calls"]
            frames_per_launch = T * prof_reps / v["calls"]
            gbs = alg[name] * frames_per_launch / (per_launch_ms * 1e-3) / 1e9
            stages[name] = {"ms_per_launch": per_launch_ms, "frames_per_launch": frames_per_launch,
                            "alg_MB_per_frame": alg[name] / 1e6, "achieved_GBs": gbs, "frac_of_hbm": gbs / hbm,
                            "share_of_step": v["ms"] / total_ms}
        dom = max(stages, key=lambda k: stages[k]["share_of_step"])
        # measured DRAM traffic of the dominant kernel from the committed ncu --set full capture, rescaled
        # to this run's frames per launch (null if that kernel was not captured)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            t = json.load(open(tpath)).get(dom)
            if t:
                traffic = t["dram_bytes_per_launch"] * stages[dom]["frames_per_launch"] / t["frames_per_launch"]
        roofline = {"kernel": dom, "bound": "hbm", "achieved": stages[dom]["achieved_GBs"], "peak": hbm,
                    "unit": "GB/s", "frac": stages[dom]["achieved_GBs"] / hbm, "traffic": traffic,
                    "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)",
                    "note": "algorithmic bytes per launch (SURVEY 8d: 40 B x tile pairs + 12 B x pixels, x frames per "
                            "launch) / in-situ CUDA-event time of the stage.  The compositing kernel is bound by the "
                            "issue rate and the L1/shared data pipe, not by HBM (ncu: issue slots 71 % busy, L1 data "
                            "pipe 73 %, DRAM 4 % of peak, profiles/r1_ncu_summary.md), so its HBM fraction is low by "
                            "construction; bind_preprocess, the other kernel BASELINE.json names, is in `stages`"}
        # tensor-pipe figure for the blendshape GEMM: 2*T*K3*npad flops per launch group
        flops_gemm = 2.0 * 3 * d["kpad"] * d["npad"]
        cpu = None
        if not args.no_cpu:
            n_cpu = min(args.cpu_frames, T)
            fps, dtc, cores = cpu_oracle_fps(model, params, baked, cam, n_cpu)
            cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"first {n_cpu} of {T} frames of the same clip, {dtc:.1f} s of host time"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(T), "frames_per_step_per_rank": T, "batch_segments": args.batch,
                       "tile_pairs_per_frame": R, "binning": "segmented depth sort (4 passes/Gaussian) + tile counts + fused emit/counting-sort "
                       "(published U7-U9: %d-pass 64-bit-key sort, %.1f MB/frame)" % ((sort_bits + 7) // 8,
                                                                                      published_sort_bytes / 1e6),
                       "gemm": "tcgen05 tf32x3" if args.gemm == 0 else "cuda-core fp32",
                       "l2": "per-batch working set (P0-P2 %.0f MB + keys/values %.0f MB) exceeds the 126 MB L2; "
                             "frame-invariant avatar streams (24 MB) stay L2-resident by design" % (
                                 48e-6 * args.batch * N_GAUSS, 24e-6 * R * args.batch),
                       "gather": ("none" if world == 1 else
                                  ("copy-engine peer-to-peer pushes (CUDA IPC, NVLink)" if args.gather == "p2p" else "NCCL gather") +
                                  " of every step's uint8 frames into rank 0 inside the timed region, on a second stream: step i's "
                                  "exchange overlaps step i+1's rendering; every rank waits for its last push before its end event "
                                  "and the time is the max over ranks"),
                       **({"host_numa": numa} if numa else {})},
            **({"diagnosis": "--no-gather: NOT a valid multi-GPU number"} if args.no_gather and world > 1 else {}),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "note": "omfs_session_render_host: pinned host params in, uint8 frames out"},
            "gpu_launches": int(launches),
            "ms_per_step_by_rank": [m / args.steps for m in ms_ranks],
            "clocks": clocks,
            "roofline": roofline,
            "stages": stages,
            "gemm_flops_per_frame": flops_gemm,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    barrier()
    if peer_gather is not None:
        peer_gather.close()
    sess.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
