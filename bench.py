#!/usr/bin/env python
"""bench.py — throughput of the LINE-MOD match path on B200 (BASELINE.json metric: RGB-D frames/s,
640x480, N templates; similarity GB/s vs roofline).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (CPU arm: the oracle port on the host cores)

Headline leg (BASELINE configs[2]): a step = one pass of the hot path over a batch of 64 distinct synthetic frames per GPU
against the 3 000-template set of configs[1] (10 % planted).  Ranks stream disjoint frames (weak scaling, no data-path
collective).
  value = frames/s with the frames resident in HBM (device pipeline only, CUDA events on the library's stream)
  e2e   = frames/s through lmb200_match_batch_submit/_collect with pinned HOST frames: H2D of every frame, kernels, D2H
          of the match lists and the host sort/unique inside the timed region; reported next to the measured H2D roof.
Further legs in the same JSON line:
  roofs            measured on this device: L2 read, L1 read, HBM read (lmb200_microbench), H2D with all ranks copying at once
  roofline         the dominant kernel against its measured roof; per-kernel fractions in `kernels`
  no_early_exit    the coarse kernel with its exact early exit switched off (workload dependence of the exit)
  config1          the reference's own frame vs the 1 950 lagergehaeuse templates (N=1 only): the exit rarely fires there
  template_sharded BASELINE configs[3] — 20 000 templates sharded over the N GPUs (north_star's split): quantisers
                   sharded by frame block + NCCL all-gather of the quantized maps, ncclAllGather of the match lists,
                   distributed host epilogue; with an in-run parity bit against the oracle and the efficiency against the
                   full set on one GPU measured in the same run
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import gc
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("LMB200_QUIET", "1")   # the stand-in NORMAL_LUT warning is recorded in the JSON line instead of stderr

ROWS, COLS = 480, 640
FRAME_BYTES = ROWS * COLS * 5
LM_BYTES = 2 * 8 * (ROWS * COLS + ROWS * COLS // 4)     # linear memories per frame (two modalities, two levels), bytes as upstream counts them


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=64, help="frames per step per GPU (configs[2]: 64)")
    ap.add_argument("--slots", type=int, default=0, help="frame slots of the detector (max_batch); 0 = frames per step.  More slots = larger chunks / deeper buffering in the streaming batch path")
    ap.add_argument("--templates", type=int, default=3000)
    ap.add_argument("--threshold", type=float, default=80.0)
    ap.add_argument("--ts-templates", type=int, default=20000, help="template_sharded leg: templates (configs[3]: 20 000)")
    ap.add_argument("--ts-frames", type=int, default=128, help="template_sharded leg: frames per step (all ranks together)")
    ap.add_argument("--no-ts-grid", action="store_true", help="template_sharded leg: skip the 2-D (template shards x frame groups) layout")
    ap.add_argument("--ts-groups", type=int, default=4, help="template_sharded leg: slot groups = steps in flight + 1 (2..4)")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--template-cache", default="", help="YAML(.gz) written/read through the product's persistence; skips addTemplate when present")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ts", action="store_true", help="skip the template_sharded leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the no_early_exit and config1 legs")
    ap.add_argument("--only-ts", action="store_true", help="run only the template_sharded leg (diagnosis)")
    return ap.parse_args()


def workload_config(args):
    """Identical for both arms (the driver compares them)."""
    return {"workload": "configs[2]: batches of %d synthetic 640x480 RGB-D frames per GPU, streamed, vs the %d templates of configs[1] "
                        "(CG+DN, T={5,8}, ~10%% planted), threshold %g" % (args.frames, args.templates, args.threshold),
            "templates": args.templates, "rows": ROWS, "cols": COLS, "threshold": args.threshold}


def measured_hbm():
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


def pin_to_gpu_cpus(index):
    """One process per GPU: run (and first-touch the pinned frame buffers) on the CPUs next to the GPU, so eight ranks do
    not pull their H2D traffic across the socket interconnect.  sysfs numa_node first; when the platform reports -1
    (round 1's boxes did), NVML's own CPU affinity mask of the device.  Best effort; returns a description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        cpus, how = [], ""
        try:
            bus = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus.lower()[-12:]).read())
            if node >= 0:
                for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
                    lo, _, hi = part.partition("-")
                    cpus += list(range(int(lo), int(hi or lo) + 1))
                how = "sysfs numa node %d" % node
        except Exception:
            pass
        if not cpus:
            ncpu = os.cpu_count() or 64
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
            cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
            how = "nvmlDeviceGetCpuAffinity"
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed and len(allowed) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, allowed)
            return "%s: pinned to %d cpus" % (how, len(allowed))
        return "%s: mask covers every allowed cpu (%d) - nothing to pin" % (how, len(allowed))
    except Exception as e:
        return "not pinned (%s)" % type(e).__name__


def planted_mask_list(n_templates):
    """Masks for the planted ~10 %: frame 0's object silhouettes + random rectangles/ellipses on frame 0."""
    from line_mod_pipeline_b200 import synth
    return synth.object_masks(0) + synth.planted_masks(n_templates // 10, seed=17)


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


def build_templates_sharded_set(det, n_templates, bgr, depth):
    """configs[3]: 10 classes x n/10 templates ('multi-object bin-picking set'); class obj00 starts with views planted on
    frame 0 (bulk addTemplates through the product), everything else random (seed 123)."""
    from line_mod_pipeline_b200 import synth
    masks = synth.object_masks(0) + synth.planted_masks(n_templates // 100, seed=17)
    res = det.addTemplates([[bgr, depth]] * len(masks), "obj00", masks)
    planted = sum(1 for tid, _ in res if tid >= 0)
    per_class = n_templates // 10
    tps = synth.random_templates(10 * per_class - planted, seed=123)
    k = 0
    for c in range(10):
        want = per_class - (planted if c == 0 else 0)
        for tp in tps[k:k + want]:
            det.addSyntheticTemplate(tp, "obj%02d" % c)
        k += want
    return planted


def copy_templates_to_oracle(det, ora):
    for cid in det.classIds():
        for t in range(det.numTemplates(cid)):
            ora.add_synthetic(det.getTemplates(cid, t), cid)


def cpu_sample(ora, n_frames, threshold, threads, frames_fn):
    """Times the oracle (CPU port of the reference path) on n_frames frames; returns (fps, seconds, matches).
    threads > 1: that many frames in flight, one oracle call (= upstream's serial matchClass) per host thread — frames
    are independent, and this uses the cores far better than splitting one frame's templates over threads
    (measured: 106 vs 18 frames/s on 8 cores).  ctypes releases the GIL during the calls."""
    from concurrent.futures import ThreadPoolExecutor
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
    frames = [synth.make_frame(i % args.frames) for i in range(per_step)]
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
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                             "sample": "%d steps x %d frames (a bounded sample of the batch: two frames per host thread) x %d templates, oracle C++ port, %d frames in flight (one per host thread)"
                                       % (args.steps, per_step, args.templates, threads),
                             "single_thread_value": fps1},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def microbench(L, kind, nbytes=0, iters=8):
    g = C.c_double(0)
    rc = L.lmb200_microbench(kind, nbytes, iters, C.byref(g))
    return g.value if rc == 0 else None


def tup(m):
    return [(int(a.x), int(a.y), float(a.similarity), int(a.class_index), int(a.template_id)) for a in m]


def template_sharded_leg(args, lm, synth, torch, dist, rank, world, local, K, W):
    """BASELINE configs[3] / north_star's multi-GPU split.  Returns the dict for the JSON line (rank 0) or None."""
    Bt = args.ts_frames - args.ts_frames % max(1, world)
    thr = args.threshold
    G = max(2, min(4, args.ts_groups))                    # slot groups: G-1 steps in flight on the GPU while the host fetches the oldest
    det = lm.getDefaultLINEMOD(device=local, max_batch=G * Bt)
    if world > 1:
        det.setOption("host_threads", max(2, len(os.sched_getaffinity(0)) // world))
    bgr0, depth0 = synth.make_frame(0)
    planted = build_templates_sharded_set(det, args.ts_templates, bgr0, depth0)
    L = lm.capi.lib()
    nb, nd = ROWS * COLS * 3, ROWS * COLS * 2
    ptr = C.c_void_p()
    assert L.lmb200_host_alloc(Bt * (nb + nd), C.byref(ptr)) == 0
    host = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(Bt * (nb + nd),))
    frames = []
    for i in range(Bt):                                   # every rank holds the same frames (template-sharded mode)
        bgr, depth = synth.make_frame(i)
        hb = host[i * (nb + nd): i * (nb + nd) + nb].reshape(ROWS, COLS, 3)
        hd = host[i * (nb + nd) + nb: (i + 1) * (nb + nd)].view(np.uint16).reshape(ROWS, COLS)
        hb[:] = bgr; hd[:] = depth
        frames.append([hb, hd])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        det.synchronize()

    # ---- full set on ONE GPU (every rank measures it on its own device; rank 0's figure is reported).  Pipelined like the
    #      sharded step: step k is enqueued on one slot half, then step k-1 is fetched (D2H + host sort/unique) from the other.
    for g in range(G):
        det.uploadFrames(frames, g * Bt)
    k1 = max(3, K // 2)

    fprep = det.prepareFetch(Bt, cap=4096 * Bt)        # result buffers marshalled once: a fetch is one C-ABI call

    def one_gpu_steps(n):
        for k in range(n):
            det.matchResident((k % G) * Bt, Bt, thr)
            if k >= G - 1:
                det.fetchResidentPrepared(fprep, ((k - (G - 1)) % G) * Bt)
        for k in range(max(0, n - (G - 1)), n):
            det.fetchResidentPrepared(fprep, (k % G) * Bt)
        return det.lists(fprep)

    one_gpu_steps(G)
    barrier()
    t0 = time.perf_counter()
    res1 = one_gpu_steps(k1)
    det.synchronize()
    dt1 = time.perf_counter() - t0
    single = {"value": Bt * k1 / dt1, "ms_per_step": 1e3 * dt1 / k1}
    out = {"workload": "configs[3]: %d templates in 10 classes, batches of %d frames (the same frames on every rank), threshold %g" % (det.numTemplates(), Bt, thr),
           "templates": det.numTemplates(), "planted_templates": planted, "frames_per_step": Bt,
           "full_set_on_1_gpu": single}
    n_ref = int(sum(len(r) for r in res1))
    if world == 1:
        value, ms_step, e2e, last = single["value"], single["ms_per_step"], None, res1
        # e2e on one GPU: host frames through the streaming batch call
        prep = [det.prepareBatch(frames, cap=4096 * Bt) for _ in range(2)]
        for k in range(3):
            det.matchPrepared(prep[0], thr)
        t0 = time.perf_counter()
        pending = None
        for k in range(k1):
            tk = det.submitPrepared(prep[k & 1], thr)
            if pending is not None:
                det.collectPrepared(*pending)
            pending = (prep[k & 1], tk)
        det.collectPrepared(*pending)
        dte = time.perf_counter() - t0
        e2e = {"value": Bt * k1 / dte, "unit": "frames/s", "ms_per_step": 1e3 * dte / k1, "h2d_bytes_per_step": Bt * FRAME_BYTES,
               "mode": "lmb200_match_batch_submit/_collect"}
        out.update({"n_gpus": 1, "value": value, "ms_per_step": ms_step, "e2e": e2e, "efficiency_vs_full_set_on_1_gpu": 1.0,
                    "frame_side": "one GPU: nothing to shard"})
    else:
        det.setTemplateShard(rank, world)
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.from_numpy(lm.comm_unique_id().copy())
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        det.commInit(uid.cpu().numpy(), rank, world)
        nown = Bt // world

        def make_runner(Bs, fp, up, own_off):
            """Pipelined sharded steps of Bs frames on G slot groups: step(upload), drain(), timed(upload) -> seconds (max over ranks)."""
            state = {"k": 0, "pending": []}

            def step(upload):
                g = state["k"] % G
                state["k"] += 1
                if upload:                                    # e2e: only the rank's own frame block crosses PCIe
                    det.uploadPrepared(up, g * Bs + own_off)
                det.matchResidentSharded(g * Bs, Bs, thr)
                state["pending"].append(g)
                if len(state["pending"]) >= G:                # keep G-1 steps in flight behind the one being fetched
                    det.fetchResidentPrepared(fp, state["pending"].pop(0) * Bs, allgather=True)

            def drain():
                while state["pending"]:
                    det.fetchResidentPrepared(fp, state["pending"].pop(0) * Bs, allgather=True)

            def timed(upload):
                for _ in range(max(2 * G + 2, W)):            # every slot group twice: first uses allocate pinned and device buffers
                    step(upload)
                drain()
                barrier()
                gc.disable()                                  # a collection on ANY rank stalls every rank at the next collective
                t0 = time.perf_counter()
                for _ in range(K):
                    step(upload)
                drain()                                       # the last step's gather + merge belongs to the timed region
                barrier()
                dt = time.perf_counter() - t0
                gc.enable()
                t = torch.tensor([dt], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                return float(t.item())
            return step, drain, timed

        uprep = det.prepareUpload(frames[rank * nown:(rank + 1) * nown])
        step, drain, timed = make_runner(Bt, fprep, uprep, rank * nown)

        dt = timed(False)
        last = det.lists(fprep)
        # where the step goes on the device (CUDA events of rank 0, a few extra steps outside the timed region)
        det.setProfiling(True); det.getProfile(reset=True)
        for _ in range(4):
            step(False)
        drain()
        pr = det.getProfile(reset=True)
        det.setProfiling(False)
        out["device_ms_per_step"] = {k: round(v / 4, 4) for k, v in pr["ms"].items() if v > 0}
        out["device_ms_per_step"]["sum"] = round(sum(pr["ms"].values()) / 4, 4)
        det.setOption("upload_async", 1)
        dte = timed(True)
        det.setOption("upload_async", 0)
        # A/B of the two design choices of the sharded step (same timed loop, no uploads)
        variants = {}
        for name, opt, val, dflt in (("host_epilogue_thread", "shard_device_epilogue", 0, 1), ("spread_on_frame_lane", "shard_overlap", 2, 1),
                                     ("no_lane_overlap", "shard_overlap", 0, 1)):
            det.setOption(opt, val)
            variants[name + "_ms_per_step"] = round(1e3 * timed(False) / K, 4)
            det.setOption(opt, dflt)
        out["variants"] = variants
        value = Bt * K / dt
        e2e = {"value": Bt * K / dte, "unit": "frames/s", "ms_per_step": 1e3 * dte / K, "h2d_bytes_per_step_per_gpu": Bt // world * FRAME_BYTES,
               "mode": "every rank uploads only its own frame block from pinned host memory (lmb200_upload_frames, async), lmb200_match_resident_sharded, lmb200_fetch_resident_allgather (complete match lists in host memory of every rank); %d steps in flight" % G}
        # all ranks must hold the identical merged lists
        sig = torch.tensor([sum(len(g) for g in last), int(sum(float(g.similarity.sum()) for g in last))], device="cuda", dtype=torch.int64)
        lo, hi = sig.clone(), sig.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        out.update({"n_gpus": world, "value": value, "ms_per_step": 1e3 * dt / K, "e2e": e2e,
                    "efficiency_vs_full_set_on_1_gpu": value / (world * single["value"]),
                    "ranks_agree": bool(torch.equal(lo, hi)),
                    "frame_side": "sharded: each rank quantises %d of the %d frames, NCCL all-gather of the quantized maps, every rank spreads all frames" % (Bt // world, Bt),
                    "collectives_per_step": "1 NCCL group (quantized maps, frame lane, main communicator) + 1 ncclAllGather (match buffers, compute lane, second communicator); std::sort + std::unique of every frame on the device (epilogue lane), finished lists land in pinned host memory"})
        # ---- 2-D layout on the same box: T template shards x 2 frame groups (world = 2 T).  Every group is a template-sharded
        #      step of its own (own NCCL communicator) on half of the step's frames, so the replicated part of the step — spread +
        #      linearize of every frame on every rank — halves.  Same call sequence, nothing new in the library.
        if world >= 4 and world % 2 == 0 and Bt % (2 * world) == 0 and not args.no_ts_grid:
            Tn = world // 2
            tr, fg = rank % Tn, rank // Tn
            det.synchronize()
            det.setTemplateShard(tr, Tn)
            ids = []
            for leader in range(0, world, Tn):
                u = torch.zeros(128, dtype=torch.uint8)
                if rank == leader:
                    u = torch.from_numpy(lm.comm_unique_id().copy())
                u = u.cuda()
                dist.broadcast(u, leader)
                ids.append(u.cpu().numpy())
            det.commInit(ids[fg], tr, Tn)
            Bg = Bt // 2
            gframes = frames[fg * Bg:(fg + 1) * Bg]
            for g in range(G):
                det.uploadFrames(gframes, g * Bg)
            fprep2 = det.prepareFetch(Bg, cap=4096 * Bg)
            nown2 = Bg // Tn
            uprep2 = det.prepareUpload(gframes[tr * nown2:(tr + 1) * nown2])
            _, _, timed2 = make_runner(Bg, fprep2, uprep2, tr * nown2)
            dt2 = timed2(False)
            last2 = det.lists(fprep2)
            det.setOption("upload_async", 1)
            dte2 = timed2(True)
            det.setOption("upload_async", 0)
            same = all(tup(last2[i]) == tup(res1[fg * Bg + i]) for i in range(Bg))      # every frame, against this rank's own 1-GPU run
            okt = torch.tensor([1 if same else 0], device="cuda")
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            v2 = Bt * K / dt2
            out["grid_2d"] = {"layout": "%d template shards x 2 frame groups" % Tn, "frames_per_step": Bt, "value": v2, "ms_per_step": 1e3 * dt2 / K,
                              "e2e": {"value": Bt * K / dte2, "unit": "frames/s", "ms_per_step": 1e3 * dte2 / K},
                              "efficiency_vs_full_set_on_1_gpu": v2 / (world * single["value"]),
                              "equals_1_gpu_run_on_every_frame_and_rank": bool(int(okt.item()) == 1)}
    # ---- in-run parity bit against the oracle (rank 0; three frames of the last timed step)
    if rank == 0:
        from oracle import oracle as O
        ora = O.Detector([dict(type=O.CG), dict(type=O.DN)], [5, 8], sim_lut=det.getSimilarityLut(), normal_lut=det.getNormalLut())
        copy_templates_to_oracle(det, ora)
        check = sorted(set([0, 1, Bt - 1]))
        ok = True
        for i in check:
            want = tup(ora.match(frames[i], thr, threads=O.max_threads()).matches(0))
            ok = ok and tup(last[i]) == want
        ok = ok and int(sum(len(r) for r in last)) == n_ref     # and the whole step equals the one-GPU run in size
        out.update({"parity": bool(ok and out.get("ranks_agree", True)), "parity_frames": check, "matches_per_step": int(sum(len(r) for r in last))})
    det.close()
    L.lmb200_host_free(ptr)
    return out if rank == 0 else None


def config1_leg(args, lm, torch, local, K):
    """The reference's own frame (benchmark/img0.png + depth0.png) vs the 1 950 lagergehaeuse templates: a realistic set
    (> 1 000 matches per frame at threshold 80) on which the coarse kernel's early exit rarely fires."""
    tpl = os.path.join(ROOT, "tests", "golden", "lagergehaeuse_templates.yml.gz")
    fix = os.path.join(ROOT, "tests", "golden", "fixture_frame.npz")
    if not (os.path.exists(tpl) and os.path.exists(fix)):
        return None
    z = np.load(fix)
    bgr, depth = np.ascontiguousarray(z["bgr"]), np.ascontiguousarray(z["depth"])
    B = args.frames
    det0 = lm.Detector.read(tpl)
    det = lm.getDefaultLINEMOD(device=local, max_batch=B, candidate_capacity=65536)
    for cid in det0.classIds():
        for t in range(det0.numTemplates(cid)):
            det.addSyntheticTemplate(det0.getTemplates(cid, t), cid)
    det0.close()
    det.uploadFrames([[bgr, depth]] * B, 0)
    out = {"workload": "configs[0]: the reference's frame (x%d copies per step) vs %d lagergehaeuse templates, threshold %g" % (B, det.numTemplates(), args.threshold)}
    for label, ee in (("early_exit", 1), ("no_early_exit", 0)):
        det.setOption("early_exit", ee)
        for _ in range(3):
            det.matchResident(0, B, args.threshold)
        det.fetchResident(0, B, cap=8192 * B)
        det.setProfiling(True); det.getProfile(reset=True)
        det.timerRecord(0)
        for _ in range(K):
            det.matchResident(0, B, args.threshold)
        det.timerRecord(1)
        ms = det.timerElapsedMs()
        res = det.fetchResident(0, B, cap=8192 * B)
        prof = det.getProfile(reset=True)
        det.setProfiling(False)
        out[label] = {"value": B * K / (ms * 1e-3), "ms_per_step": ms / K, "sim_coarse_ms_per_launch": prof["ms"]["sim_coarse"] / max(1, prof["launches"]["sim_coarse"]),
                      "sim_local_ms_per_launch": prof["ms"]["sim_local"] / max(1, prof["launches"]["sim_local"]), "matches_per_frame": len(res[0])}
    det.close()
    return out


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import line_mod_pipeline_b200 as lm
    from line_mod_pipeline_b200 import synth
    K_ = lm.capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    affinity = pin_to_gpu_cpus(local) if world > 1 else "single process: not pinned"
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    B, K, W = args.frames, args.steps, max(args.warmup, 3)
    n_tpl = args.templates
    L = lm.capi.lib()
    if args.only_ts:
        ts = template_sharded_leg(args, lm, synth, torch, dist, rank, world, local, K, W)
        if rank == 0:
            print(json.dumps({"template_sharded": ts}))
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- measured roofs of this device (and of the shared host path: every rank copies at once)
    roofs = {"l2_read_GBps": microbench(L, K_.MB_L2_READ), "l1_read_GBps": microbench(L, K_.MB_L1_READ),
             "hbm_read_GBps": microbench(L, K_.MB_HBM_READ), "how": "lmb200_microbench (csrc/microbench.cu): 128-bit loads, two kernel forms, best of 8 launches each"}
    if dist is not None:
        dist.barrier()
    h2d = microbench(L, K_.MB_H2D, B * FRAME_BYTES, 8)
    if dist is not None:
        t = torch.tensor([h2d, -h2d], device="cuda", dtype=torch.float64)
        s = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.MIN); dist.all_reduce(s, op=dist.ReduceOp.SUM)
        roofs.update({"h2d_GBps_per_gpu_min": float(t[0].item()), "h2d_GBps_per_gpu_max": -float(t[1].item()), "h2d_GBps_all_gpus": float(s[0].item()),
                      "h2d_how": "%d ranks copying %d MB from pinned host memory at the same time" % (world, B * FRAME_BYTES // 1000000)})
    else:
        roofs.update({"h2d_GBps_per_gpu_min": h2d, "h2d_GBps_per_gpu_max": h2d, "h2d_GBps_all_gpus": h2d, "h2d_how": "one rank, %d MB from pinned host memory" % (B * FRAME_BYTES // 1000000)})

    bgr0, depth0 = synth.make_frame(0)
    if args.template_cache and os.path.exists(args.template_cache):
        det0 = lm.Detector.read(args.template_cache)
        det = lm.getDefaultLINEMOD(device=local, max_batch=max(2 * B, args.slots))
        for cid in det0.classIds():
            for t in range(det0.numTemplates(cid)):
                det.addSyntheticTemplate(det0.getTemplates(cid, t), cid)
        planted = det.numTemplates("planted")
        det0.close()
    else:
        det = lm.getDefaultLINEMOD(device=local, max_batch=max(2 * B, args.slots))
        planted = build_templates_product(det, n_tpl, bgr0, depth0)
        if args.template_cache and rank == 0:
            det.write(args.template_cache)

    if world > 1:   # one process per GPU: the host epilogue threads of all ranks share the box's cores
        det.setOption("host_threads", max(2, len(os.sched_getaffinity(0)) // world))
    # ---- frames: pinned host memory (e2e) + resident copies (value); ranks stream disjoint frames
    nb, nd = ROWS * COLS * 3, ROWS * COLS * 2
    ptr = C.c_void_p()
    assert L.lmb200_host_alloc(B * (nb + nd), C.byref(ptr)) == 0
    host = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(B * (nb + nd),))
    frames = []
    for i in range(B):
        bgr, depth = synth.make_frame(rank * B + i)
        hb = host[i * (nb + nd): i * (nb + nd) + nb].reshape(ROWS, COLS, 3)
        hd = host[i * (nb + nd) + nb: (i + 1) * (nb + nd)].view(np.uint16).reshape(ROWS, COLS)
        hb[:] = bgr; hd[:] = depth
        frames.append([hb, hd])
    det.uploadFrames(frames, 0)
    det.uploadFrames(frames, B)      # second slot group (double buffering): consecutive steps alternate between the two

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        det.synchronize()

    def timed_resident(k, alternate=False):
        det.timerRecord(0)          # CUDA events on the library's compute stream
        for i in range(k):
            det.matchResident((i & 1) * B if alternate else 0, B, args.threshold)
        det.timerRecord(1)
        return det.timerElapsedMs()

    for _ in range(W):
        det.matchResident(0, B, args.threshold)
    det.synchronize()
    det.fetchResident(0, B)          # (grows the candidate stores now if this workload needs it)
    det.setProfiling(True)
    det.getProfile(reset=True)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ms = timed_resident(K)
    barrier()
    res = det.fetchResident(0, B)
    prof = det.getProfile(reset=True)
    det.setProfiling(False)
    n_matches = int(sum(len(r) for r in res))
    # ---- the headline pass: the same K steps without the per-kernel profiling events, alternating between the two slot
    #      groups, so the frame side of step k+1 (frame lane) overlaps the template side of step k (compute lane)
    for _ in range(W):
        timed_resident(2, alternate=True)
    barrier()
    ms_ov = timed_resident(K, alternate=True)
    barrier()
    clocks = sampler.summary()
    n_ov = [int(sum(len(r) for r in det.fetchResident(g * B, B))) for g in (0, 1)]
    assert n_ov == [n_matches, n_matches], "the overlapped steps returned %r matches, the serial pass %d" % (n_ov, n_matches)
    if dist is not None:
        t = torch.tensor([ms, ms_ov], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_ov = float(t[0].item()), float(t[1].item())
    serial = {"value": B * K * world / (ms * 1e-3), "ms_per_step": ms / K,
              "what": "the same K steps on ONE slot group with the library's per-kernel CUDA events on (one lane): the pass `kernels`, `roofline` and the shares come from"}
    value = B * K * world / (ms_ov * 1e-3)

    # ---- the same step with the coarse kernel's exact early exit switched off
    no_exit = None
    if not args.no_extra:
        det.setOption("early_exit", 0)
        for _ in range(2):
            det.matchResident(0, B, args.threshold)
        det.setProfiling(True); det.getProfile(reset=True)
        ms_ne = timed_resident(max(3, K // 2))
        res_ne = det.fetchResident(0, B)
        p_ne = det.getProfile(reset=True)
        det.setProfiling(False)
        det.setOption("early_exit", 1)
        assert int(sum(len(r) for r in res_ne)) == n_matches, "the early exit changed the result"
        no_exit = {"value_per_gpu": B * max(3, K // 2) / (ms_ne * 1e-3), "sim_coarse_ms_per_launch": p_ne["ms"]["sim_coarse"] / max(1, p_ne["launches"]["sim_coarse"]),
                   "requested_bytes_per_launch": 16 * p_ne["chunks_coarse"],
                   "note": "identical match lists; on this synthetic set ~94 % of the templates are random and leave after one modality when the exit is on"}

    # ---- e2e: host frames through lmb200_match_batch (H2D + kernels + D2H + host sort/unique)
    e2e = None
    single = None
    if not args.no_e2e:
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
        h2d_GBps = B * K * world * FRAME_BYTES / dt / 1e9
        e2e = {"value": B * K * world / dt, "unit": "frames/s", "h2d_bytes_per_step": B * FRAME_BYTES,
               "d2h_bytes_per_step": B * (32 + 1024 * 16), "ms_per_step": 1e3 * dt / K,
               "mode": "lmb200_match_batch_submit/_collect, step k+1 submitted before step k is collected",
               "h2d_GBps_all_gpus": h2d_GBps, "frac_of_h2d_roof": h2d_GBps / roofs["h2d_GBps_all_gpus"] if roofs["h2d_GBps_all_gpus"] else None,
               # the timed region ends with the slowest rank, and the ranks' host links are not equal (NUMA / PCIe topology):
               # the same figure against N x the slowest link measured when all ranks copy at once
               "frac_of_slowest_link_roof": h2d_GBps / (world * roofs["h2d_GBps_per_gpu_min"]) if roofs.get("h2d_GBps_per_gpu_min") else None,
               "blocking_call": {"value": B * K * world / dt_sync, "ms_per_step": 1e3 * dt_sync / K}}
        # single-frame latency of the reference-facing call (configs[1] literally: one frame, host buffers in, sorted matches
        # out): lmb200_match, whose kernel sequence is replayed as a CUDA graph; the blocking batch call with one frame beside it
        def latency(fn):
            lat = []
            for i in range(80):
                t0 = time.perf_counter()
                fn()
                lat.append(time.perf_counter() - t0)
            lat = sorted(lat[20:])
            return {"median_ms": 1e3 * lat[len(lat) // 2], "p90_ms": 1e3 * lat[int(len(lat) * 0.9)], "frames_per_s": 1.0 / lat[len(lat) // 2]}
        prep_s = det.prepareSingle(frames[0], cap=8192)
        prep1 = det.prepareBatch(frames[:1], cap=8192)
        single = latency(lambda: det.matchPreparedSingle(prep_s, args.threshold))
        single["call"] = "lmb200_match (CUDA graph replay of the frame's kernel sequence), pinned host frame in, sorted match list out"
        det.setOption("cuda_graph", 0)
        single["without_cuda_graph"] = latency(lambda: det.matchPreparedSingle(prep_s, args.threshold))
        det.setOption("cuda_graph", 1)
        single["lmb200_match_batch_of_1"] = latency(lambda: det.matchPrepared(prep1, args.threshold))
        assert det.matchPreparedSingle(prep_s, args.threshold) == len(res[0]), "single-frame call and resident path disagree"

    # ---- other legs
    cfg1 = config1_leg(args, lm, torch, local, max(3, K // 2)) if (world == 1 and not args.no_extra) else None
    ts = None
    if not args.no_ts:
        if dist is not None:
            dist.barrier()
        ts = template_sharded_leg(args, lm, synth, torch, dist, rank, world, local, K, W)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- per-kernel accounting (CUDA-event time inside the timed region) and the roofline of the dominant kernel
    hbm, hbm_src = measured_hbm()
    launches = {k: v for k, v in prof["launches"].items() if k != "upload"}
    kernels = {}
    for k in launches:
        if launches[k]:
            kernels[k] = {"ms_total": round(prof["ms"][k], 4), "launches": launches[k], "ms_per_launch": round(prof["ms"][k] / launches[k], 5),
                          "share_of_step": round(prof["ms"][k] / ms * (1 if world == 1 else 1), 4)}
    px0, px1 = ROWS * COLS, ROWS * COLS // 4
    # algorithmic HBM bytes per step of the frame-side kernels (B frames): what must cross HBM at least once
    alg_step = {"pyrdown": B * (3 * px0 + 3 * px1), "cg_quantize": B * (3 * px0 + px0 + 3 * px1 + px1), "dn_quantize": B * (2 * px0 + px0),
                "median": B * 2 * px0, "decimate": B * (px0 // 4 + px1),
                "linearize": B * 2 * (px0 + 8 * px0 + px1 + 4 * px1)}     # per modality: read q, write 8 B/px (L0 strips) and 4 B/px (L1, nibble-packed)
    for k, b in alg_step.items():
        if k in kernels and prof["ms"][k] > 0:
            per_step_ms = prof["ms"][k] / K
            kernels[k]["alg_bytes_per_step"] = b
            kernels[k]["alg_GBps"] = round(b / (per_step_ms * 1e-3) / 1e9, 1)
            kernels[k]["frac_of_hbm"] = round(b / (per_step_ms * 1e-3) / 1e9 / hbm, 4)
    # similarity kernels: the linear memories they gather from are cache-resident, so their roofs are L1/L2, not HBM
    req_coarse = 16.0 * prof["chunks_coarse"] * K      # chunk loads are counted per step by the fetch above (one step's counters)
    req_local = 2.0 * prof["bytes_local"] * K          # two 8-byte loads per 8 gathered bytes (realignment window)
    prof["bytes_local"] = prof["bytes_local"] * K
    for k, alg, req in (("sim_coarse", prof["bytes_coarse"], req_coarse), ("sim_local", prof["bytes_local"], req_local)):
        if k in kernels and prof["ms"][k] > 0:
            t_s = prof["ms"][k] * 1e-3
            kernels[k].update({"algorithmic_bytes_per_launch": alg / launches[k], "algorithmic_GBps": round(alg / t_s / 1e9, 1),
                               "requested_bytes_per_launch": req / launches[k], "requested_GBps": round(req / t_s / 1e9, 1),
                               "requested_frac_of_l1_roof": round(req / t_s / 1e9 / roofs["l1_read_GBps"], 4) if roofs["l1_read_GBps"] else None})
    dom = max(kernels, key=lambda k: kernels[k]["ms_total"]) if kernels else None
    ncu = {}
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass
    roofline = None
    if dom in ("sim_coarse", "sim_local"):
        kname = "similarity_coarse_kernel" if dom == "sim_coarse" else "similarity_local_kernel"
        nj = ncu.get(kname, {})
        same_shape = nj.get("frames") == B and nj.get("templates") == n_tpl
        t_launch = prof["ms"][dom] / launches[dom] * 1e-3
        roofline = {"kernel": kname, "bound": "l2", "unit": "GB/s", "peak": roofs["l2_read_GBps"],
                    "peak_source": "measured live: lmb200_microbench L2_READ (128-bit loads over an L2-resident 48 MB buffer, L1 bypassed); HBM is not this kernel's roof: DRAM traffic per launch is %s bytes" % nj.get("dram_bytes_per_launch"),
                    "traffic": nj.get("dram_bytes_per_launch") if same_shape else None,
                    "requested_GBps": kernels[dom]["requested_GBps"], "requested_frac_of_l1_roof": kernels[dom]["requested_frac_of_l1_roof"],
                    "algorithmic_GBps": kernels[dom]["algorithmic_GBps"], "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes_per_launch"]}
        if same_shape and nj.get("l2_bytes_per_launch"):
            roofline["achieved"] = round(nj["l2_bytes_per_launch"] / t_launch / 1e9, 1)
            roofline["achieved_how"] = "L2->L1 read bytes + L1->L2 write bytes of one launch (ncu l1tex__m_xbar2l1tex_read_bytes + lts write sectors, profiles/ncu_traffic.json, same launch shape) / live CUDA-event time"
        else:
            roofline["achieved"] = kernels[dom]["requested_GBps"]
            roofline["achieved_how"] = "bytes the kernel requested (counted on the device, after the early exit) / live CUDA-event time: an UPPER bound of its L2 traffic (part of it hits L1)"
        roofline["frac"] = round(roofline["achieved"] / roofline["peak"], 4) if roofline["peak"] else None
    elif dom is not None:
        roofline = {"kernel": dom, "bound": "hbm", "unit": "GB/s", "peak": hbm, "peak_source": hbm_src,
                    "achieved": kernels[dom].get("alg_GBps"), "frac": kernels[dom].get("frac_of_hbm"), "traffic": None}
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

    cfg = workload_config(args)
    cfg.update({"frames_per_step_per_gpu": B, "cpu_affinity": affinity, "planted_templates": planted, "shard": "frames",
                "l2": "step inputs %.0f MB of frames + %.0f MB of linear memories written per step: larger than the 126 MB L2, nothing survives from step to step"
                      % (B * FRAME_BYTES / 1e6, B * (2 * (8 * px0 + 4 * px1)) / 1e6),
                "value_pass": "K steps alternating between two resident slot groups of %d frames (lmb200_match_resident): the frame side of step k+1 runs on the frame lane while the template side of step k runs on the compute lane" % B,
                "tables": {"similarity_lut": "circular (default)", "normal_lut": "stand-in" if det.normalLutIsStandin() else "user-supplied"}})
    line = {"metric": "rgbd_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_ov / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": cfg, "serial_profiled_pass": serial,
            "e2e": e2e, "gpu_launches": int(sum(launches.values())), "roofline": roofline, "cpu_baseline": cpu,
            "clocks": clocks, "roofs": roofs, "kernels": kernels, "similarity_GBps_algorithmic": sim_gbps, "single_frame": single,
            "matches_per_step": n_matches, "no_early_exit": no_exit, "config1": cfg1, "template_sharded": ts}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
