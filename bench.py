#!/usr/bin/env python
"""bench.py — throughput of the LINE-MOD match path on B200 (BASELINE.json metric: RGB-D frames/s,
640x480, N templates; similarity GB/s vs roofline).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (CPU arm: the oracle port on the host cores)

A step = one pass of the hot path over a batch of B synthetic frames (distinct frames, B*1.5 MB of
inputs > L2) against the template set of BASELINE.json configs[1] (3 000 templates, 10 % planted).
value  = frames/s with the frames resident in HBM (device pipeline only, CUDA events on the library's stream).
e2e    = frames/s through lmb200_match_batch with pinned HOST frames: H2D of every frame, kernels,
         D2H of the match lists and the host sort/unique inside the timed region.
Multi-GPU (default --shard frames, BASELINE configs[2]): each rank streams its own frames against the
full template set (weak scaling, no data-path collective).  --shard templates (configs[3]) splits
the template set across ranks and merges the match lists with one ncclAllGather per step.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS, COLS = 480, 640
FRAME_BYTES = ROWS * COLS * 5


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=96, help="frames per step per GPU")
    ap.add_argument("--templates", type=int, default=3000)
    ap.add_argument("--threshold", type=float, default=80.0)
    ap.add_argument("--shard", default="frames", choices=["frames", "templates"])
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--template-cache", default="", help="YAML(.gz) written/read through the product's persistence; skips addTemplate when present")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons WHILE the timed region runs (B200_PROFILING.md clocks line).
    Polls NVML in-process every ~2 ms (the timed region is tens of ms; nvidia-smi -lms 200 would see one sample);
    falls back to nvidia-smi when pynvml is unavailable."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.mask, self.stop_flag, self.max_mhz, self.src = index, [], 0, False, None, "nvml"
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv, self.src = None, "nvidia-smi"

    def run(self):
        while not self.stop_flag:
            try:
                if self.nv is not None:
                    self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                    try:
                        self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    except Exception:
                        self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    time.sleep(0.002)
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    self.sm.append(float(out[0])); self.max_mhz = float(out[1]); self.mask |= int(out[2].strip(), 16)
                    time.sleep(0.1)
            except Exception:
                time.sleep(0.01)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": [n for bit, n in self.REASONS.items() if self.mask & bit], "samples": len(sm), "source": self.src}


def pin_to_gpu_numa_node(index):
    """One process per GPU: run (and first-touch the pinned frame buffers) on the CPUs of the GPU's NUMA node, so eight
    ranks do not pull their H2D traffic across the socket interconnect.  Best effort; returns a description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bdf = bus.lower()[-12:]                      # 0000:1b:00.0
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return "numa_node unknown"
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return "numa node %d, %d cpus" % (node, len(allowed))
        return "numa node %d has no allowed cpus" % node
    except Exception as e:
        return "not pinned (%s)" % type(e).__name__


def build_templates_product(det, n_templates, bgr, depth):
    """configs[1] template set through the PRODUCT's own addTemplate: ~10 % planted on frame 0, rest random (seed 99)."""
    from line_mod_pipeline_b200 import synth
    planted = 0
    for m in planted_mask_list(n_templates):
        tid, _ = det.addTemplate([bgr, depth], "planted", m)
        planted += tid >= 0
    for tp in synth.random_templates(n_templates - planted):
        det.addSyntheticTemplate(tp, "rand")
    return planted


def planted_mask_list(n_templates):
    """Masks for the planted ~10 %: frame 0's object silhouettes + random rectangles/ellipses on frame 0."""
    from line_mod_pipeline_b200 import synth
    return synth.object_masks(0) + synth.planted_masks(n_templates // 10, seed=17)


def copy_templates_to_oracle(det, ora):
    from oracle import oracle as O
    for cid in det.classIds():
        for t in range(det.numTemplates(cid)):
            ora.add_synthetic(det.getTemplates(cid, t), cid)


def cpu_sample(ora_templates_from, n_frames, threshold, threads, frames_fn):
    """Times the oracle (CPU port of the reference path) on n_frames frames; returns (fps, seconds, matches).
    threads > 1: that many frames in flight, one oracle call (= upstream's serial matchClass) per host thread — frames
    are independent, and this uses the cores far better than splitting one frame's templates over threads
    (measured: 106 vs 18 frames/s on 8 cores).  ctypes releases the GIL during the calls."""
    from concurrent.futures import ThreadPoolExecutor
    ora = ora_templates_from
    frames = [frames_fn(i) for i in range(n_frames)]
    one = lambda f: len(ora.match([f[0], f[1]], threshold, threads=1).matches(0))
    t0 = time.perf_counter()
    if threads > 1:
        with ThreadPoolExecutor(threads) as ex:
            nm = sum(ex.map(one, frames))
    else:
        nm = sum(one(f) for f in frames)
    dt = time.perf_counter() - t0
    return n_frames / dt, dt, nm


def run_reference(args):
    """CPU arm: the oracle restatement of cv::linemod::Detector::match (real OpenCV-contrib linemod is not
    buildable offline — DESIGN.md), all host threads, same config/metric/unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from line_mod_pipeline_b200 import synth
    from oracle import oracle as O
    threads = O.max_threads()
    lut = synth.default_normal_lut()
    ora = O.Detector([dict(type=O.CG), dict(type=O.DN)], [5, 8], normal_lut=lut)
    bgr0, depth0 = synth.make_frame(0)
    planted = 0
    for m in planted_mask_list(args.templates):
        tid, _ = ora.add_template([bgr0, depth0], "planted", m)
        planted += tid >= 0
    for tp in synth.random_templates(args.templates - planted):
        ora.add_synthetic(tp, "rand")
    per_step = 2 * threads                     # bounded sample: two frames per host thread and step
    frames = [synth.make_frame(i % 96) for i in range(per_step)]
    for _ in range(args.warmup):
        cpu_sample(ora, min(per_step, threads), args.threshold, threads, lambda i: frames[i])
    dt = 0.0
    for s in range(args.steps):
        dt += cpu_sample(ora, per_step, args.threshold, threads, lambda i: frames[i])[1]
    fps = args.steps * per_step / dt
    # upstream's matchClass is serial: the same path on ONE thread, for the record (2 frames)
    fps1 = cpu_sample(ora, 2, args.threshold, 1, lambda i: frames[i])[0]
    line = {"impl": "reference", "metric": "rgbd_frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "configs[1]: 640x480 RGB-D frame vs %d templates (CG+DN, T={5,8}), threshold %g" % (args.templates, args.threshold),
                       "frames_per_step": per_step},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                             "sample": "%d steps x %d frames x %d templates, oracle C++ port, %d frames in flight (one per host thread)" % (args.steps, per_step, args.templates, threads),
                             "single_thread_value": fps1},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import line_mod_pipeline_b200 as lm
    from line_mod_pipeline_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    affinity = pin_to_gpu_numa_node(local) if world > 1 else "single process: not pinned"
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    B, K, W = args.frames, args.steps, max(args.warmup, 3)
    n_tpl = args.templates * (world if args.shard == "templates" else 1)
    bgr0, depth0 = synth.make_frame(0)
    if args.template_cache and os.path.exists(args.template_cache):
        det0 = lm.Detector.read(args.template_cache)
        det = lm.getDefaultLINEMOD(device=local, max_batch=B * (2 if args.shard == "templates" else 1))
        for cid in det0.classIds():
            for t in range(det0.numTemplates(cid)):
                det.addSyntheticTemplate(det0.getTemplates(cid, t), cid)
        planted = det.numTemplates("planted")
        det0.close()
    else:
        det = lm.getDefaultLINEMOD(device=local, max_batch=B * (2 if args.shard == "templates" else 1))
        planted = build_templates_product(det, n_tpl, bgr0, depth0)
        if args.template_cache and rank == 0:
            det.write(args.template_cache)
    if args.shard == "templates" and world > 1:
        det.setTemplateShard(rank, world)
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.from_numpy(lm.comm_unique_id().copy())
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        det.commInit(uid.cpu().numpy(), rank, world)

    # ---- frames: pinned host memory (e2e) + resident copies (value)
    L = lm.capi.lib()
    nb, nd = ROWS * COLS * 3, ROWS * COLS * 2
    ptr = C.c_void_p()
    assert L.lmb200_host_alloc(B * (nb + nd), C.byref(ptr)) == 0
    host = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(B * (nb + nd),))
    frames = []
    first = 0 if args.shard == "templates" else rank * B
    for i in range(B):
        bgr, depth = synth.make_frame(first + i)
        hb = host[i * (nb + nd): i * (nb + nd) + nb].reshape(ROWS, COLS, 3)
        hd = host[i * (nb + nd) + nb: (i + 1) * (nb + nd)].view(np.uint16).reshape(ROWS, COLS)
        hb[:] = bgr; hd[:] = depth
        frames.append([hb, hd])
    det.uploadFrames(frames, 0)

    allg = args.shard == "templates" and world > 1
    if args.shard == "templates":
        det.uploadFrames(frames, B)      # second slot half: step k+1 is computed while step k's matches are gathered

    state = {"half": 0, "pending": None, "last": None}

    def step():
        """frames mode: enqueue the device pipeline for the B resident frames.
        templates mode: enqueue step k on one slot half, then gather + merge step k-1 from the other half
        (ncclAllGather + host merge overlap the kernels of step k)."""
        if not allg:
            det.matchResident(0, B, args.threshold)
            return
        h = state["half"]
        det.matchResident(h * B, B, args.threshold)
        if state["pending"] is not None:
            state["last"] = det.fetchResident(state["pending"] * B, B, allgather=True, cap=2048 * B)
        state["pending"], state["half"] = h, h ^ 1

    def drain():
        if allg and state["pending"] is not None:
            state["last"] = det.fetchResident(state["pending"] * B, B, allgather=True, cap=2048 * B)
            state["pending"] = None

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        det.synchronize()

    for _ in range(W):
        step()
    drain()
    det.synchronize()
    det.setProfiling(True)
    det.getProfile(reset=True)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t_wall = time.perf_counter()
    det.timerRecord(0)          # CUDA events on the library's compute stream
    for _ in range(K):
        step()
    det.timerRecord(1)
    drain()                      # templates mode: the last step's gather + merge belongs to the timed region
    barrier()
    t_wall = time.perf_counter() - t_wall
    ms = det.timerElapsedMs()
    if allg:
        ms = 1e3 * t_wall        # the gather/merge of the last step ends after the compute-stream event: use the wall clock between the barriers
    clocks = sampler.summary()
    if allg:                     # outside the timed region (also collects device-side counters)
        step(); drain()
        res = state["last"]
    else:
        res = det.fetchResident(0, B)
    prof = det.getProfile(reset=True)
    det.setProfiling(False)
    n_matches = int(sum(len(r) for r in res))
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    frames_total = B * K * (world if args.shard == "frames" else 1)
    value = frames_total / (ms * 1e-3)

    # ---- e2e: host frames through lmb200_match_batch (H2D + kernels + D2H + host sort/unique)
    e2e = None
    if not args.no_e2e and not allg:
        # marshal once: every timed step is exactly one C-ABI batch (submit + collect); consecutive steps are pipelined
        # (step k+1 is submitted before step k is collected), the way a frame stream is processed
        preps = [det.prepareBatch(frames, cap=2048 * B) for _ in range(2)]
        for _ in range(3):
            det.matchPrepared(preps[0], args.threshold)
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            n_sync = det.matchPrepared(preps[0], args.threshold)
        torch.cuda.synchronize()
        dt_sync = time.perf_counter() - t0
        for k in range(4):   # warm-up of the pipelined path (allocates the second ticket's pinned staging)
            tk = det.submitPrepared(preps[k & 1], args.threshold)
            if k:
                det.collectPrepared(*pending)
            pending = (preps[k & 1], tk)
        det.collectPrepared(*pending)
        barrier()
        t0 = time.perf_counter()
        pending = None
        for k in range(K):
            tk = det.submitPrepared(preps[k & 1], args.threshold)
            if pending is not None:
                n_e2e = det.collectPrepared(*pending)
            pending = (preps[k & 1], tk)
        n_e2e = det.collectPrepared(*pending)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert n_e2e == n_matches and n_sync == n_matches, "e2e path returned %d/%d matches, resident path %d" % (n_e2e, n_sync, n_matches)
        if dist is not None:
            t = torch.tensor([dt, dt_sync], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt, dt_sync = float(t[0].item()), float(t[1].item())
        e2e = {"value": B * K * world / dt, "unit": "frames/s", "h2d_bytes_per_step": B * FRAME_BYTES,
               "d2h_bytes_per_step": B * (24 + 1024 * 16), "ms_per_step": 1e3 * dt / K,
               "mode": "lmb200_match_batch_submit/_collect, step k+1 submitted before step k is collected",
               "blocking_call": {"value": B * K * world / dt_sync, "ms_per_step": 1e3 * dt_sync / K}}

    # ---- single-frame latency of the reference-facing call (configs[1] literally: one frame, host buffers in, matches out)
    single = None
    if not allg and not args.no_e2e:
        prep1 = det.prepareBatch(frames[:1], cap=8192)
        lat = []
        for i in range(60):
            t0 = time.perf_counter()
            det.matchPrepared(prep1, args.threshold)
            lat.append(time.perf_counter() - t0)
        lat = sorted(lat[10:])
        single = {"median_ms": 1e3 * lat[len(lat) // 2], "p90_ms": 1e3 * lat[int(len(lat) * 0.9)], "frames_per_s": 1.0 / lat[len(lat) // 2]}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant similarity kernel (algorithmic bytes / CUDA-event time)
    peak, peak_src = measured_peaks()
    launches = {k: v for k, v in prof["launches"].items() if k != "upload"}
    kernels = {}
    for k in launches:
        if launches[k]:
            kernels[k] = {"ms_total": round(prof["ms"][k], 4), "launches": launches[k], "ms_per_launch": round(prof["ms"][k] / launches[k], 5)}
    bytes_frame_side = K * B * (FRAME_BYTES + 2 * 8 * (ROWS * COLS + ROWS * COLS // 4))
    # similarityLocal bytes are counted on the device per step; the fetch above read one step's counters
    prof["bytes_local"] = prof["bytes_local"] * (K if not allg else 1)
    alg = {"sim_coarse": prof["bytes_coarse"], "sim_local": prof["bytes_local"], "linearize": K * B * 2 * 9 * (ROWS * COLS + ROWS * COLS // 4)}
    for k, b in alg.items():
        if k in kernels and prof["ms"][k] > 0:
            kernels[k]["alg_bytes_per_launch"] = b / launches[k]
            kernels[k]["alg_GBps"] = round(b / (prof["ms"][k] * 1e-3) / 1e9, 1)
    dom = "sim_coarse"
    ach = kernels.get(dom, {}).get("alg_GBps", 0.0)
    traffic = None   # DRAM bytes per launch of the same kernel from the committed ncu --set full capture (same launch shape only)
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["similarity_coarse_kernel"]
        if B == 96 and n_tpl == 3000 and args.shard == "frames":
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"kernel": "similarity_coarse_kernel", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": round(ach / peak, 4) if peak else None, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": kernels.get(dom, {}).get("alg_bytes_per_launch"),
                "note": "achieved = algorithmic bytes (sum nf*P, SURVEY 8d) / kernel time. The linear memories are L2-resident (the "
                        "gather is bounded by L2; HBM copy bandwidth is the reported denominator), and the kernel's exact early exit "
                        "(a pass ends once no position can still exceed the threshold) skips part of the algorithmic reads"}
    sim_ms = prof["ms"]["sim_coarse"] + prof["ms"]["sim_local"]
    sim_gbps = (prof["bytes_coarse"] + prof["bytes_local"]) / (sim_ms * 1e-3) / 1e9 if sim_ms > 0 else None

    # ---- cpu baseline: the oracle port on a bounded sample (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import oracle as O
        ora = O.Detector([dict(type=O.CG), dict(type=O.DN)], [5, 8], sim_lut=det.getSimilarityLut(), normal_lut=det.getNormalLut())
        copy_templates_to_oracle(det, ora)
        threads = O.max_threads()
        nfr = args.cpu_frames or 4 * B            # ~20 thread-seconds of CPU work (the step's frames, cycled)
        fps, dt, nm = cpu_sample(ora, nfr, args.threshold, threads, lambda i: frames[i % B])
        fps1, _, _ = cpu_sample(ora, 2, args.threshold, 1, lambda i: frames[i % B])   # upstream's matchClass is serial
        cpu = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": "%d frames (the step's frames, cycled) x %d templates, oracle C++ port, %d frames in flight (one per host thread), %.1f s" % (nfr, n_tpl, threads, dt),
               "single_thread_value": fps1}

    line = {"metric": "rgbd_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": "configs[1]: 640x480 RGB-D frame vs %d templates (CG+DN, T={5,8}), threshold %g; step = batch of %d distinct frames per GPU"
                                   % (n_tpl, args.threshold, B),
                       "frames_per_step_per_gpu": B, "cpu_affinity": affinity, "templates": n_tpl, "planted_templates": planted, "shard": args.shard,
                       "l2": "inputs larger than L2: %.0f MB of frames + %.0f MB of linear memories per step" % (B * FRAME_BYTES / 1e6, B * 6.144)},
            "e2e": e2e, "gpu_launches": int(sum(launches.values())), "roofline": roofline, "cpu_baseline": cpu,
            "clocks": clocks, "kernels": kernels, "similarity_GBps": sim_gbps, "single_frame": single, "matches_per_step": n_matches,
            "candidates_per_step": None}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
