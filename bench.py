"""bench.py — surgery-render frames/s (512^2, 100k Gaussians) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4] [--scaling strong|weak]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path (FLAME blendshapes + skinning -> triangle-bound Gaussian transform -> tile
binning -> compositing) over ONE clip of BASELINE.json configs[2] (300 frames, 512x512, 100k Gaussians).

  --scaling strong (default)  the configuration as BASELINE.json states it: the ONE clip is frame-sharded over the N
            ranks (sharding.frame_block) and every rank's finished uint8 frames are gathered into rank 0's clip
            buffer over NVLink inside the timed region; value = clip frames / time.
  --scaling weak              every rank renders its own clip (e.g. its own surgical plan); reported beside the strong
            figure as `weak` when N > 1.

  value     frames/s, whole job, inputs resident in HBM, CUDA-event time of K steps, max over ranks
  e2e       the same metric through the reference-facing C-ABI call with HOST buffers: parameters go host->device and
            the finished frames come back inside the timed region as PNG files encoded on the device
            (omfs_session_render_host_png: what render_with_gaussians writes to disk); `e2e_raw` is the same with raw
            uint8 frames (omfs_session_render_host)
  roofline  the dominant kernel's achieved algorithmic GB/s (SURVEY.md §8d bytes x measured work, divided by its
            CUDA-event time from a separate profiling pass) against MEASURED_PEAKS.json
  cpu_baseline  the C oracle (oracle/, a port — the reference has no CPU renderer) on the host cores, bounded sample
            of the same workload, N = 1 only

--config 3 / 4 run the other named workloads (1024^2 x 16 views x 500k Gaussians, sharded by frame; the 64-plan x
120-frame sweep, sharded by plan) through the same machinery.  --impl reference times the oracle port alone.
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

UNIT = "frames/s"
KEYS = ("expr", "rotation", "neck_pose", "jaw_pose", "eyes_pose", "translation")

# BASELINE.json configs[2..4].  `units` are what the ranks shard (frames, frames, plans); a unit renders
# `segs_per_unit` images.
WORKLOADS = {
    2: dict(metric="surgery-render frames/s (512^2, 100k Gaussians)", width=512, height=512, n_gauss=100_000,
            units=300, unit_name="frames", views=1, frames_per_unit=1, batch=75, pairs_per_seg=0,
            name="configs[2]: 512x512 x {units}-frame surgery video, 100k FLAME-bound Gaussians"),
    3: dict(metric="multi-view render images/s (1024^2, 16 views, 500k Gaussians)", width=1024, height=1024,
            n_gauss=500_000, units=300, unit_name="frames", views=16, frames_per_unit=1, batch=128,
            pairs_per_seg=4_500_000,
            name="configs[3]: 1024x1024 x 16 camera views x {units} frames, 500k Gaussians (images = frames x views)"),
    4: dict(metric="plan-sweep images/s (512^2, 100k Gaussians, 120 frames per plan)", width=512, height=512,
            n_gauss=100_000, units=64, unit_name="plans", views=1, frames_per_unit=120, batch=60, pairs_per_seg=0,
            name="configs[4]: surgical plan sweep, {units} BSSO plans x 120 frames, batched blendshape GEMM + render"),
}


def host_threads():
    """Hardware threads this process may use, read BEFORE any NUMA pinning."""
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def make_inputs(wl, n_units, seed=0):
    """model, FrameParams over all n_units * frames_per_unit time steps, baked avatar, cameras."""
    import omfs_b200  # noqa: F401
    from omfs_b200 import avatar, cameras, render_surgery as rs, synthetic
    fpu = wl["frames_per_unit"]
    model, params, av, cam = synthetic.make_scene(n_gauss=wl["n_gauss"], n_frames=fpu if fpu > 1 else n_units,
                                                  width=wl["width"], height=wl["height"], seed=seed)
    cams = [cam]
    if wl["views"] > 1:
        cams = cameras.ring_cameras(wl["views"], synthetic.camera_distance(wl["width"], wl["height"]), (0, 0, 0), 0.3,
                                    wl["width"], wl["height"])
    if fpu > 1:
        # plan sweep: every plan is the reference's scalar BSSO edit (render_surgery.py:40-42, 130-139) of the same
        # 120-frame clip, plans on the UI's regular grid in [-15, 15] mm (app.py:817-822)
        recs = []
        for mm in np.linspace(-15.0, 15.0, 64)[:n_units]:
            rec = rs._edit_record(params.as_dict(), 0.0, rs.compute_offset(float(mm), 1.0), None)
            recs.append(synthetic.FrameParams.from_dict(rec, n_verts=model.n_verts))
        cat = lambda k: np.concatenate([getattr(r, k) for r in recs], axis=0)
        params = synthetic.FrameParams(params.shape, cat("expr"), cat("rotation"), cat("neck_pose"), cat("jaw_pose"),
                                       cat("eyes_pose"), cat("translation"), params.static_offset, cat("dynamic_offset"))
    return model, params, avatar.bake(av), cams


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


def cpu_oracle_fps(wl, model, params, baked, cams, n_sample, threads):
    """Images/s of the oracle port on `threads` host threads for the first n_sample time steps (all views)."""
    import oracle
    used = oracle.set_num_threads(threads)
    nv = len(cams)
    chunk = max(1, 20 // nv)  # time steps per oracle call: bounds the host memory of the pair lists
    t0 = time.perf_counter()
    for lo in range(0, n_sample, chunk):
        hi = min(n_sample, lo + chunk)
        for c in cams:
            oracle.render(model, params.slice(lo, hi), baked, [c.pack()] * (hi - lo), wl["width"], wl["height"])
    dt = time.perf_counter() - t0
    return n_sample * nv / dt, dt, used


def run_reference(args, wl, rank):
    """The reference arm: the oracle port (the reference has no CPU renderer of its own; its in-tree pieces of the
    path are the scalar parameter edits) on ALL host threads — set explicitly, a torchrun rank would otherwise
    inherit OMP_NUM_THREADS=1 — over a bounded sample of the workload per step.  Rank 0 alone works and prints."""
    if rank != 0:
        return 0
    threads = host_threads()
    n_sample = args.ref_frames
    model, params, baked, cams = make_inputs(wl, max(1, n_sample) if wl["frames_per_unit"] == 1 else 1)
    n_sample = min(n_sample, params.n_frames)
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_fps(wl, model, params, baked, cams, min(4, n_sample), threads)
    t0 = time.perf_counter()
    cores = threads
    for _ in range(args.steps):
        _, _, cores = cpu_oracle_fps(wl, model, params, baked, cams, n_sample, threads)
    dt = time.perf_counter() - t0
    fps = args.steps * n_sample * len(cams) / dt
    sample = (f"{n_sample} of the workload's time steps per step x {len(cams)} view(s) ({wl['width']}x{wl['height']}, "
              f"{wl['n_gauss']} Gaussians), {cores} OpenMP threads of {threads} hardware threads")
    line = {
        "impl": "reference", "metric": wl["metric"], "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"].format(units=args.units)},
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
    ap.add_argument("--config", type=int, default=2, choices=(2, 3, 4), help="BASELINE.json configs[] index")
    ap.add_argument("--scaling", default="strong", choices=("strong", "weak"))
    ap.add_argument("--frames", "--units", dest="units", type=int, default=0,
                    help="frames (configs 2, 3) or plans (config 4) per step; 0 = the configuration's own")
    ap.add_argument("--batch", type=int, default=0, help="segments (images) per launch group; 0 = the workload's default")
    ap.add_argument("--gemm", type=int, default=0, help="0 tensor-core blendshape GEMM, 1 CUDA-core")
    ap.add_argument("--cpu-frames", type=int, default=300, help="time steps of the cpu_baseline sample")
    ap.add_argument("--ref-frames", type=int, default=60, help="time steps per step of --impl reference")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--stream-depth", type=int, default=3, choices=(2, 3), help="clips in flight of the streamed e2e path")
    ap.add_argument("--sync-steps", action="store_true",
                    help="A/B: every step joins its compositing into the rendering stream before the next one starts")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the weak-scaling companion measurement")
    ap.add_argument("--no-gather", action="store_true", help="diagnosis only: skip the frame gather (flagged in config)")
    ap.add_argument("--gather", default="p2p", choices=("p2p", "fused", "nccl"),
                    help="frame exchange: copy-engine peer-to-peer pushes into rank 0 (default), the compositing kernel's "
                         "own stores into rank 0's peer-mapped slot (fused: no local frame buffer, no copy), or one NCCL gather")
    args = ap.parse_args()
    wl = WORKLOADS[args.config]
    if args.units <= 0:
        args.units = wl["units"]
    if args.batch <= 0:
        args.batch = wl["batch"]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, wl, rank)
    n_host_threads = host_threads()

    import torch
    import torch.distributed as dist
    import omfs_b200  # noqa: F401
    from omfs_b200 import runtime, sharding, synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    host_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")   # host-side meetings that must not spin a GPU
    L = runtime.load_library()
    runtime.check(L.omfs_device_check(local_rank))

    W, H, NG, nv, fpu = wl["width"], wl["height"], wl["n_gauss"], wl["views"], wl["frames_per_unit"]
    U = args.units
    frame_bytes = H * W * 3
    segs_per_unit = nv * fpu
    model, params_all, baked, cams = make_inputs(wl, U, seed=0)
    d_cam = torch.from_numpy(np.ascontiguousarray(np.stack([c.pack() for c in cams]), dtype=np.float32)).to(dev)

    def dev_params(p):
        t = {k: torch.from_numpy(np.ascontiguousarray(getattr(p, k), dtype=np.float32)).to(dev) for k in KEYS}
        ptrs = {k: v.data_ptr() for k, v in t.items()}
        ptrs["cams"] = d_cam.data_ptr()
        return t, ptrs

    def exchange(obj):
        out = [None] * world
        dist.all_gather_object(out, obj, group=host_group)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.Stream(device=dev)
    comm_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    def frame_sums(t_u8, n_frames):
        """Per-frame checksums (sum of the frame's 32-bit words), computed where the data is."""
        if n_frames == 0:
            return []
        return t_u8.reshape(n_frames, -1).view(torch.int32).sum(dim=1, dtype=torch.int64).tolist()

    # ------------------------------------------------------------------------------------------------------
    def run_mode(mode):
        """Device-resident timing of `mode` ('strong': the one clip sharded over the ranks; 'weak': one clip per rank)."""
        if mode == "strong":
            lo, hi = sharding.frame_block(U, rank, world)
            total_units = U
            params = params_all
        else:
            lo, hi = 0, U
            total_units = U * world
            params = params_all
            if rank > 0 and fpu == 1:   # every rank its own expression/pose track, same subject and avatar
                params = synthetic.make_frame_params(U, seed=99 + rank)
                params.shape[:] = params_all.shape
        n_local_units = hi - lo
        T_local = n_local_units * fpu                 # time steps this rank renders per step
        S_local = T_local * nv                        # images this rank renders per step
        local = params.slice(lo * fpu, hi * fpu)
        # launch groups: the rank's block in equal groups no larger than the workload's default (150 images -> 3 x 50,
        # 38 -> 1 x 38).  Steps overlap each other through the session's deferred join, so a short block needs no
        # split to keep both streams busy (measured: 38 frames as one group 1.09 ms, as two 1.17 ms per step).
        batch = args.batch
        if mode == "strong" and world > 1 and S_local > 0:
            groups = -(-S_local // batch)
            batch = max(nv, -(-S_local // groups))
            batch = -(-batch // nv) * nv
        sess = runtime.Session(model, baked, W, H, max_batch=batch, device=local_rank, gemm_impl=args.gemm,
                               pair_capacity=int(wl["pairs_per_seg"]) * batch)
        sess.set_subject(params.shape, params.static_offset)
        # a step's last compositing launch overlaps the next step's front end; consumers are ordered by sess.join()
        sess.set_deferred_join(not args.sync_steps)
        keep, d_ptrs = dev_params(local)
        per_units = -(-total_units // world) if mode == "strong" else U     # slot size on rank 0, in units
        slot_bytes = per_units * segs_per_unit * frame_bytes
        n_bufs = 2 if world > 1 else 1
        frames_bufs = [torch.empty((max(S_local, 1), H, W, 3), dtype=torch.uint8, device=dev) for _ in range(n_bufs)]
        gathered, peer = None, None
        if world > 1 and args.gather == "nccl":
            gathered = torch.empty((world, slot_bytes), dtype=torch.uint8, device=dev) if rank == 0 else None
            pad = torch.zeros(slot_bytes, dtype=torch.uint8, device=dev)
        elif world > 1:
            peer = sharding.PeerFrameGather(slot_bytes, rank, world, exchange)
        ev_rendered = [torch.cuda.Event() for _ in frames_bufs]
        ev_gathered = [torch.cuda.Event() for _ in frames_bufs]
        step_no = [0]
        local_only = [False]   # content check of the fused exchange: one step into the local buffer, nothing sent
        my_bytes = S_local * frame_bytes

        def step():
            b = step_no[0] % n_bufs
            step_no[0] += 1
            if world > 1:
                stream.wait_event(ev_gathered[b])
            fused = world > 1 and args.gather == "fused" and not args.no_gather and not local_only[0]
            if S_local:
                # fused: the compositing kernel's uint8 stores ARE the exchange (rank 0's slot is mapped into this process)
                sess.render_device(d_ptrs, T_local, nv, d_out_u8=peer.slot_ptr() if fused else frames_bufs[b].data_ptr(),
                                   stream=stream.cuda_stream)
            if fused:
                return
            if local_only[0]:
                return
            if world > 1 and not args.no_gather:
                # the only exchange of the path: finished frames to rank 0 over NVLink, on a second stream, which waits
                # for the step's compositing (the rendering stream itself goes straight on to the next step)
                if args.sync_steps:
                    ev_rendered[b].record(stream)
                    comm_stream.wait_event(ev_rendered[b])
                else:
                    sess.join(comm_stream.cuda_stream)
                if peer is not None:
                    # every rank pushes its block into its slot of rank 0's buffer: sender-side copy engine
                    peer.push(frames_bufs[b].data_ptr(), my_bytes, comm_stream.cuda_stream)
                else:
                    with torch.cuda.stream(comm_stream):
                        pad[:my_bytes] = frames_bufs[b].reshape(-1)[:my_bytes]
                        dist.gather(pad, list(gathered.unbind(0)) if rank == 0 else None, dst=0)
                ev_gathered[b].record(comm_stream)

        def drain():   # the timed region ends only when every step is composited and its frames have arrived on rank 0
            sess.join(stream.cuda_stream)
            if world > 1:
                stream.wait_stream(comm_stream)

        for _ in range(args.warmup):
            step()
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
            step()
        drain()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = runtime.launch_count() - launches0
        clocks = sampler.stop() if rank == 0 else None
        sess.sync()
        pairs_per_seg = sess.stats()["pairs"] / max(S_local, 1)
        ms_ranks = [ms]
        if world > 1:
            all_ms = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
            dist.all_gather(all_ms, torch.tensor([ms], dtype=torch.float64, device=dev))
            ms_ranks = [float(x.item()) for x in all_ms]
        ms_max = max(ms_ranks)
        value = total_units * segs_per_unit * args.steps / (ms_max / 1e3)

        # ---- content check, outside the timed region: rank 0's buffer is cleared, one more step runs, and every
        # slot must then hold exactly the frames its rank rendered (per-frame checksums computed on each side)
        verified = None
        if world > 1 and not args.no_gather:
            if rank == 0:
                if peer is not None:
                    runtime.check(L.omfs_device_memset(peer.base, 0, world * slot_bytes))
                else:
                    gathered.zero_()
            barrier()
            step()
            drain()
            if args.gather == "fused":   # the same frames once more, into the local buffer, for the sums below
                local_only[0] = True
                step()
                drain()
                local_only[0] = False
            barrier()
            b = (step_no[0] - 1) % n_bufs
            mine = frame_sums(frames_bufs[b][:S_local], S_local)
            every = exchange(mine)
            if rank == 0:
                if peer is not None:
                    g = torch.empty(world * slot_bytes, dtype=torch.uint8, device=dev)
                    runtime.check(L.omfs_push_frames(g.data_ptr(), peer.base, world * slot_bytes, None))
                    torch.cuda.synchronize()
                    g = g.view(world, slot_bytes)
                else:
                    g = gathered
                bad = []
                for r in range(world):
                    n_r = len(every[r])
                    got = frame_sums(g[r][: n_r * frame_bytes], n_r)
                    if got != every[r]:
                        bad.append(r)
                if mode == "strong" and fpu == 1 and nv == 1 and not bad:
                    # the slots, read back to back, ARE the clip in frame order
                    clip = g.reshape(-1)[: U * frame_bytes]
                    if frame_sums(clip, U) != [x for r in range(world) for x in every[r]]:
                        bad.append(-1)
                if bad:
                    raise SystemExit(f"bench.py: gathered frames on rank 0 differ from what ranks {bad} rendered "
                                     f"({mode} scaling)")
                verified = {"slots": world, "frames": sum(len(e) for e in every), "method":
                            "rank 0's buffer cleared, one extra step, per-frame 32-bit-word sums of every slot compared "
                            "with the sums each rank computed over its own frames"}
        res = dict(mode=mode, value=value, ms_per_step=ms_max / args.steps, ms_by_rank=[m / args.steps for m in ms_ranks],
                   launches=int(launches), clocks=clocks, pairs_per_seg=pairs_per_seg, batch=batch, S_local=S_local,
                   T_local=T_local, verified=verified, sess=sess, d_ptrs=d_ptrs, keep=keep, local=local,
                   frames_buf=frames_bufs[0], peer=peer)
        return res

    def close_mode(res):
        if res["peer"] is not None:
            res["peer"].close()
        res["sess"].close()

    main_res = run_mode(args.scaling)
    sess = main_res["sess"]
    S_local, T_local, local = main_res["S_local"], main_res["T_local"], main_res["local"]

    # ---- end to end through the host-buffer C-ABI calls (pinned host memory both ways), same sharding
    host_in = {k: runtime.PinnedArray(getattr(local, k).shape, np.float32) for k in KEYS}
    for k, v in host_in.items():
        v.array[...] = getattr(local, k)

    class HostParams:
        pass

    hp = HostParams()
    for k, v in host_in.items():
        setattr(hp, k, v.array)
    hp.dynamic_offset = None
    # e2e calls cover at most `e2e_chunk` time steps each, which bounds the pinned output buffers (configs[3])
    e2e_chunk = max(1, min(T_local, (1 << 30) // max(1, nv * frame_bytes))) if T_local else 1
    n_chunk_seg = e2e_chunk * nv
    png_cap = int(L.omfs_png_max_bytes(W, H))
    host_png = runtime.PinnedArray((max(1, n_chunk_seg) * png_cap,), np.uint8)
    host_off = runtime.PinnedArray((n_chunk_seg + 1,), np.uint64)
    host_raw = runtime.PinnedArray((max(1, n_chunk_seg), H, W, 3), np.uint8)

    def hp_slice(lo, hi):
        q = HostParams()
        for k in KEYS:
            setattr(q, k, getattr(hp, k)[lo:hi])
        q.dynamic_offset = None
        return q

    png_bytes = [0]

    def step_host_png():
        png_bytes[0] = 0
        for lo in range(0, T_local, e2e_chunk):
            hi = min(T_local, lo + e2e_chunk)
            _, off = sess.render_host_png(hp_slice(lo, hi), cams, out_png=host_png.array, out_offsets=host_off.array)
            png_bytes[0] += int(off[-1])

    def step_host_raw():
        for lo in range(0, T_local, e2e_chunk):
            hi = min(T_local, lo + e2e_chunk)
            sess.render_host(hp_slice(lo, hi), cams, want_u8=True, out_u8=host_raw.array[: (hi - lo) * nv])

    e2e_steps = max(2, min(args.steps, 5))

    def time_host(fn):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        units = U if args.scaling == "strong" else U * world
        return units * segs_per_unit * e2e_steps / float(t_e.item())

    e2e_png = time_host(step_host_png)
    # the sink's streams really are the frames: decode this rank's first and last frame of the last chunk
    if S_local:
        from PIL import Image
        import io
        n_last = ((T_local - 1) % e2e_chunk + 1) * nv
        off = host_off.array
        dec = np.asarray(Image.open(io.BytesIO(host_png.array[int(off[n_last - 1]):int(off[n_last])].tobytes())))
        step_host_raw()
        if not np.array_equal(dec, host_raw.array[n_last - 1]):
            raise SystemExit("bench.py: the device-encoded PNG of the last frame does not decode to the raw frame")
    png_total = torch.tensor([png_bytes[0]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(png_total)
    e2e_raw = time_host(step_host_raw)

    # ---- the same path as a STREAM of clips: submit step i, then collect step i-1 (two clips in flight, two sets of
    # pinned output buffers).  Every step still uploads its own parameters from pinned host memory and delivers its own
    # PNG streams to host memory inside the timed region; what changes is that step i+1 renders while the tail of
    # step i (encode of the last group, offsets, stream copy) is in flight — the overlap a blocking call cannot have.
    n_lanes = 3 if args.stream_depth >= 3 else 2
    lanes = [(host_png, host_off)] + [(runtime.PinnedArray((max(1, n_chunk_seg) * png_cap,), np.uint8),
                                       runtime.PinnedArray((n_chunk_seg + 1,), np.uint64)) for _ in range(n_lanes - 1)]
    lane_no = [0]
    in_flight = [0]
    streamed_bytes = [0]

    def collect_one():
        _, off = sess.collect_host_png()
        streamed_bytes[0] += int(off[-1])
        in_flight[0] -= 1

    def step_host_streamed():
        for lo in range(0, T_local, e2e_chunk):
            hi = min(T_local, lo + e2e_chunk)
            if in_flight[0] == n_lanes:
                collect_one()
            png_, off_ = lanes[lane_no[0] % n_lanes]
            lane_no[0] += 1
            sess.submit_host_png(hp_slice(lo, hi), cams, png_.array, off_.array)
            in_flight[0] += 1

    def time_streamed():
        step_host_streamed()
        while in_flight[0]:
            collect_one()
        barrier()
        streamed_bytes[0] = 0
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_host_streamed()
        while in_flight[0]:
            collect_one()
        torch.cuda.synchronize()
        t_e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        units_ = U if args.scaling == "strong" else U * world
        return units_ * segs_per_unit * e2e_steps / float(t_e.item())

    e2e_streamed = time_streamed() if S_local or world > 1 else None
    if S_local and streamed_bytes[0] != png_bytes[0] * e2e_steps:
        raise SystemExit("bench.py: the streamed clips delivered %d bytes, the blocking calls %d per step" % (
            streamed_bytes[0], png_bytes[0]))
    h2d = sum(v.array.nbytes for v in host_in.values()) + 160 * nv
    h2d_total = h2d * (world if world > 1 else 1)

    # ---- per-stage profile for the roofline (separate pass: costs a host sync per batch), rank 0's share
    d_ptrs, frames_buf = main_res["d_ptrs"], main_res["frames_buf"]
    st, prof_reps = None, 3
    if S_local:
        def dev_pass():
            sess.render_device(d_ptrs, T_local, nv, d_out_u8=frames_buf.data_ptr(), stream=stream.cuda_stream)
            sess.sync()
        sess.set_profiling(True)
        for _ in range(2):
            dev_pass()
        sess.set_profiling(False)
        for _ in range(2):
            dev_pass()
        sess.set_profiling(True)
        for _ in range(prof_reps):
            dev_pass()
        st = sess.stage_ms()
        sess.set_profiling(False)
    dims = sess.dims()

    # ---- the companion weak-scaling figure (N > 1 only)
    weak = None
    if world > 1 and args.scaling == "strong" and not args.no_weak:
        close_mode(main_res)
        w = run_mode("weak")
        weak = {"value": w["value"], "unit": UNIT, "ms_per_step": w["ms_per_step"], "ms_per_step_by_rank": w["ms_by_rank"],
                "units_per_rank": U, "gathered_frames_verified": w["verified"],
                "note": "one clip per rank (round-1 definition), frames of every rank gathered on rank 0 in the timed region"}
        close_mode(w)
        main_res["peer"], main_res["sess"] = None, None

    if rank == 0:
        hbm, tflops, peak_kind = peaks()
        R = main_res["pairs_per_seg"]
        hw = W * H
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        sort_bits = L.omfs_binning_sort_bits(main_res["batch"], W, H)
        # algorithmic bytes per IMAGE for each stage.  flame .. bind_preprocess, composite: SURVEY.md §8d (FLAME and
        # face frames are per time step: divided by the views that share them).  The binning is charged what THIS
        # design has to move: depth sort = one 4-byte histogram read + 4 passes x 16 B per Gaussian;
        # tile counts/ranges = 20 B per Gaussian + 16 B per tile; emit+scatter = 24 B per Gaussian + 4 B per pair.
        alg = {
            "flame": (4.0 * 3 * dims["V"] * 2 + 44.0 * dims["V"]) / nv,
            "face_frames": 80.0 * dims["F"] / nv,
            "bind_preprocess": 288.0 * NG,
            "depth_sort": (4.0 + 16.0 * 4) * NG,
            "tile_ranges": 20.0 * NG + 16.0 * tiles,
            "emit_scatter": 24.0 * NG + 4.0 * R,
            "composite": 40.0 * R + 12.0 * hw,
        }
        published_sort_bytes = (12.0 + 8.0 + 24.0 * ((sort_bits + 7) // 8) + 8.0) * R   # SURVEY U7+U8+U9
        stages, roofline = {}, None
        if st:
            total_ms = sum(v["ms"] for v in st.values())
            for name, v in st.items():
                if not v["calls"]:
                    continue
                per_launch_ms = v["ms"] / v["calls"]
                imgs_per_launch = S_local * prof_reps / v["calls"]
                gbs = alg[name] * imgs_per_launch / (per_launch_ms * 1e-3) / 1e9
                stages[name] = {"ms_per_launch": per_launch_ms, "images_per_launch": imgs_per_launch,
                                "alg_MB_per_image": alg[name] / 1e6, "achieved_GBs": gbs, "frac_of_hbm": gbs / hbm,
                                "share_of_step": v["ms"] / total_ms}
            dom = max(stages, key=lambda k: stages[k]["share_of_step"])
            # measured DRAM traffic of the dominant kernel from the committed ncu --set full capture, rescaled
            # to this run's images per launch (null if that kernel was not captured for this workload)
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
            if os.path.exists(tpath) and args.config == 2:
                t = json.load(open(tpath)).get(dom)
                if t:
                    traffic = t["dram_bytes_per_launch"] * stages[dom]["images_per_launch"] / t["frames_per_launch"]
            roofline = {"kernel": dom, "bound": "hbm", "achieved": stages[dom]["achieved_GBs"], "peak": hbm,
                        "unit": "GB/s", "frac": stages[dom]["achieved_GBs"] / hbm, "traffic": traffic,
                        "traffic_source": "profiles/ncu_traffic.json (ncu --set full capture of this kernel, rescaled "
                                          "to this run's images per launch)" if traffic else None,
                        "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)",
                        "note": "algorithmic bytes per launch (SURVEY 8d: 40 B x tile pairs + 12 B x pixels, x images per "
                                "launch) / in-situ CUDA-event time of the stage.  The compositing kernel is bound by the "
                                "issue rate and the L1/shared data pipe, not by HBM (ncu: profiles/), so its HBM fraction "
                                "is low by construction; bind_preprocess, the other kernel BASELINE.json names, is in "
                                "`stages`"}
        cpu = None
        if not args.no_cpu and world == 1:
            n_cpu = min(args.cpu_frames, params_all.n_frames)
            if nv > 1:
                n_cpu = min(n_cpu, 2)
            fps, dtc, cores = cpu_oracle_fps(wl, model, params_all, baked, cams, n_cpu, n_host_threads)
            cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"first {n_cpu} of {params_all.n_frames} time steps x {nv} view(s) of the same workload, "
                             f"{dtc:.1f} s of host time, {cores} OpenMP threads of {n_host_threads} hardware threads"}
        per_rank = "%d %s per rank per step" % (-(-U // world) if args.scaling == "strong" else U, wl["unit_name"])
        line = {
            "metric": wl["metric"], "value": main_res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"].format(units=U) + (
                           "" if world == 1 else (", the ONE workload sharded by %s over %d ranks" % (wl["unit_name"][:-1], world)
                                                  if args.scaling == "strong" else ", one workload PER RANK")),
                       "sharding": per_rank, "batch_segments": main_res["batch"], "tile_pairs_per_image": R,
                       "binning": "segmented depth sort (4 passes/Gaussian) + tile counts + fused emit/counting-sort "
                                  "(published U7-U9: %d-pass 64-bit-key sort, %.1f MB/image)" % (
                                      (sort_bits + 7) // 8, published_sort_bytes / 1e6),
                       "gemm": "tcgen05 tf32x3" if args.gemm == 0 else "cuda-core fp32",
                       "l2": "per-batch working set (P0-P2 %.0f MB + values %.0f MB) exceeds the 126 MB L2; "
                             "frame-invariant avatar streams (%.0f MB) stay L2-resident by design" % (
                                 48e-6 * main_res["batch"] * NG, 4e-6 * R * main_res["batch"], 240e-6 * NG),
                       "gather": ("none" if world == 1 else
                                  ({"p2p": "copy-engine peer-to-peer pushes (CUDA IPC, NVLink)",
                                    "fused": "compositing kernel stores straight into rank 0's peer-mapped slot (CUDA IPC, NVLink)",
                                    "nccl": "NCCL gather"}[args.gather]) +
                                  " of every step's uint8 frames into rank 0 inside the timed region, on a second stream: "
                                  "step i's exchange overlaps step i+1's rendering; every rank waits for its last push "
                                  "before its end event and the time is the max over ranks"),
                       "steps_overlap": ("synchronous steps (--sync-steps)" if args.sync_steps else
                                         "consecutive steps overlap on the device like the launch groups inside one step "
                                         "(session deferred join): the timed region ends after a join of every step's "
                                         "compositing and, for N > 1, of every step's frame push"),
                       "host_threads": n_host_threads,
                       **({"host_numa": numa} if numa else {})},
            **({"diagnosis": "--no-gather: NOT a valid multi-GPU number"} if args.no_gather and world > 1 else {}),
            "e2e": {"value": e2e_streamed, "unit": UNIT, "h2d_bytes_per_step": int(h2d_total),
                    "d2h_bytes_per_step": int(png_total.item()), "steps": e2e_steps,
                    "note": "omfs_session_submit_host_png / omfs_session_collect_host_png on every rank, up to three clips in "
                            "flight: every step uploads its own parameters from pinned host memory and delivers its own "
                            "frames to host memory as PNG files encoded on the device (filter + deflate + CRC in "
                            "csrc/png.cu) inside the timed region; step i+1 renders while the tail of step i is encoded "
                            "and copied.  The streams are byte-identical to the blocking call's (tests/test_gpu_png.py; "
                            "byte count checked here), whose last frame is decoded and compared with the raw frame after "
                            "the timed region; d2h bytes = the streams actually copied"},
            "e2e_blocking": {"value": e2e_png, "unit": UNIT, "h2d_bytes_per_step": int(h2d_total),
                             "d2h_bytes_per_step": int(png_total.item()), "steps": e2e_steps,
                             "note": "omfs_session_render_host_png: one blocking call per step (the round-1 form of e2e, "
                                     "with PNG streams instead of raw frames); the difference to `e2e` is latency a "
                                     "blocking call cannot hide (first group's front end, last group's encode and copy)"},
            "e2e_raw": {"value": e2e_raw, "unit": UNIT, "h2d_bytes_per_step": int(h2d_total),
                        "d2h_bytes_per_step": int(U * (1 if args.scaling == "strong" else world) * segs_per_unit * frame_bytes),
                        "steps": e2e_steps, "note": "omfs_session_render_host: raw uint8 frames out, blocking"},
            "gpu_launches": main_res["launches"],
            "ms_per_step_by_rank": main_res["ms_by_rank"],
            "gathered_frames_verified": main_res["verified"],
            "clocks": main_res["clocks"],
            "roofline": roofline,
            "stages": stages,
            "gemm_flops_per_frame": 2.0 * 3 * dims["kpad"] * dims["npad"],
            "cpu_baseline": cpu,
            **({"weak": weak} if weak else {}),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(group=host_group)   # the other ranks wait on the host, not with a spinning NCCL kernel
    if main_res["sess"] is not None:
        close_mode(main_res)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
