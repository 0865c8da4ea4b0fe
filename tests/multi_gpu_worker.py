"""Worker for tests/test_gpu_multi.py, bench.py-independent manual runs:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/multi_gpu_worker.py [n_templates]
BASELINE configs[3]: 20 000 templates in 10 classes, template-sharded across N GPUs, match lists combined with
ncclAllGather.  Both step variants must equal the oracle on every frame, on every rank:
  * lmb200_match_resident          (frame side replicated on every rank)
  * lmb200_match_resident_sharded  (quantisers sharded by frame block + NCCL all-gather of the quantized maps)
and lmb200_fetch_resident_allgather must do so on its synchronous path (after lmb200_match_resident; a sharded step whose
gathered headers carry the "record area too small" flag) and on the asynchronous one (gather on the compute lane + the
handle's epilogue thread), with one and with two sharded steps in flight."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import line_mod_pipeline_b200 as lm
from line_mod_pipeline_b200 import synth
from oracle import oracle as O


def tup(m):
    return [(int(a.x), int(a.y), float(a.similarity), int(a.class_index), int(a.template_id)) for a in m]


def main():
    os.environ.setdefault("LMB200_QUIET", "1")
    n_templates = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_frames = 2 * world * 2                      # multiple of world, >= 2*world: the distributed epilogue runs
    frames = [list(synth.make_frame(i)) for i in range(n_frames)]
    det = lm.getDefaultLINEMOD(device=local, max_batch=2 * n_frames)
    ora = O.Detector([dict(type=O.CG), dict(type=O.DN)], [5, 8], sim_lut=det.getSimilarityLut(), normal_lut=det.getNormalLut())
    planted = 0
    for m in synth.object_masks(0) + synth.object_masks(1):
        tid, _ = det.addTemplate(frames[0], "obj00", m)
        otid, _ = ora.add_template(frames[0], "obj00", m)
        assert tid == otid
        planted += tid >= 0
    per_class = n_templates // 10
    tps = synth.random_templates(n_templates - planted, seed=123)
    k = 0
    for c in range(10):
        cid = "obj%02d" % c
        want = per_class - (planted if c == 0 else 0)
        for tp in tps[k:k + want]:
            det.addSyntheticTemplate(tp, cid); ora.add_synthetic(tp, cid)
        k += want
    assert det.numTemplates() == 10 * per_class and det.numClasses() == 10
    det.setTemplateShard(rank, world)
    uid = torch.from_numpy(lm.comm_unique_id().copy()).cuda() if rank == 0 else torch.zeros(128, dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    det.commInit(uid.cpu().numpy(), rank, world)
    ok = True
    checked = 0
    wants = {}
    n = n_frames // world
    blank = [np.zeros_like(frames[0][0]), np.zeros_like(frames[0][1])]
    for variant in ("replicated", "sharded", "sharded_host_epilogue", "few_frames", "sharded_many_matches", "two_in_flight_a", "two_in_flight_b"):
        thr = 62.0 if variant in ("few_frames", "sharded_many_matches") else 80.0
        nf = n_frames if variant != "few_frames" else 3
        order = list(range(n_frames))
        if variant == "two_in_flight_b":
            continue                                # fetched inside two_in_flight_a's iteration
        # std::sort + std::unique on the device (default) or on the handle's host epilogue thread
        det.setOption("shard_device_epilogue", 0 if variant == "sharded_host_epilogue" else 1)
        if variant.startswith("sharded") or variant == "two_in_flight_a":
            # only the rank's own frame block is uploaded: the other blocks arrive as quantized maps over NCCL
            det.uploadFrames([frames[i] if rank * n <= i < (rank + 1) * n else blank for i in range(n_frames)], 0)
            det.matchResidentSharded(0, nf, thr)
            if variant == "two_in_flight_a":        # a second step on the other slot group (frames in reverse order) before the first fetch
                rev = order[::-1]
                det.uploadFrames([frames[rev[i]] if rank * n <= i < (rank + 1) * n else blank for i in range(n_frames)], n_frames)
                det.matchResidentSharded(n_frames, nf, thr)
        else:
            det.uploadFrames(frames[:nf], 0)
            det.matchResident(0, nf, thr)
        got = det.fetchResident(0, nf, allgather=True, cap=400000)
        if variant == "two_in_flight_a":
            got2 = det.fetchResident(n_frames, nf, allgather=True, cap=400000)
            if [tup(g) for g in got2] != [tup(g) for g in got[::-1]]:
                ok = False
                print("rank %d: the second step in flight differs from the first (reversed frame order)" % rank, flush=True)
        ncheck = min(nf, world) if variant == "sharded_many_matches" else nf   # the low threshold costs the oracle seconds per frame
        for i in range(rank, ncheck, world):       # every rank holds every list: each checks its share against the oracle
            key = (i, thr)
            if key not in wants:
                wants[key] = tup(ora.match(frames[i], thr, threads=max(1, O.max_threads() // world)).matches(0))
            g = tup(got[i])
            checked += 1
            if g != wants[key]:
                ok = False
                print("rank %d %s thr %g frame %d: %d vs %d matches DIFFER" % (rank, variant, thr, i, len(g), len(wants[key])), flush=True)
        # every rank must hold the identical merged result
        sig = torch.tensor([sum(len(g) for g in got), int(sum(float(g.similarity.sum()) for g in got))], device="cuda", dtype=torch.int64)
        lo, hi = sig.clone(), sig.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if not torch.equal(lo, hi):
            ok = False
            print("rank %d %s: ranks disagree on the merged lists" % (rank, variant), flush=True)
    prof = det.getProfile()
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_PARITY", "OK" if int(t.item()) == 1 else "FAIL", "world", world, "templates", det.numTemplates(), "frames", n_frames,
              "frames checked on rank 0", checked, "coarse bytes/rank", prof["bytes_coarse"], flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
