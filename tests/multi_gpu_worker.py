"""Worker for tests/test_gpu_multi.py and manual runs:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/multi_gpu_worker.py
Template-sharded match across N GPUs + ncclAllGather of the match buffers must equal the oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import line_mod_pipeline_b200 as lm
from line_mod_pipeline_b200 import synth
from oracle import oracle as O


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    frames = [list(synth.make_frame(i)) for i in range(3)]
    det = lm.getDefaultLINEMOD(device=local, max_batch=4)
    ora = O.Detector([dict(type=O.CG), dict(type=O.DN)], [5, 8], sim_lut=det.getSimilarityLut(), normal_lut=det.getNormalLut())
    for m in synth.object_masks(0) + synth.object_masks(1):
        src = frames[0]
        tid, _ = det.addTemplate(src, "planted", m)
        otid, _ = ora.add_template(src, "planted", m)
        assert tid == otid
    for tp in synth.random_templates(500):
        det.addSyntheticTemplate(tp, "rand"); ora.add_synthetic(tp, "rand")
    det.setTemplateShard(rank, world)
    uid = torch.from_numpy(lm.comm_unique_id().copy()).cuda() if rank == 0 else torch.zeros(128, dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    det.commInit(uid.cpu().numpy(), rank, world)
    det.uploadFrames(frames, 0)
    ok = True
    for thr in (80.0, 50.0):
        det.matchResident(0, 3, thr)
        got = det.fetchResident(0, 3, allgather=True, cap=200000)
        for i in range(3):
            want = ora.match(frames[i], thr, threads=4).matches(0)
            g = [(int(a.x), int(a.y), float(a.similarity), int(a.class_index), int(a.template_id)) for a in got[i]]
            w = [(int(a.x), int(a.y), float(a.similarity), int(a.class_index), int(a.template_id)) for a in want]
            if g != w:
                ok = False
                print("rank %d thr %g frame %d: %d vs %d matches DIFFER" % (rank, thr, i, len(g), len(w)), flush=True)
    # the shard really is a shard
    prof = det.getProfile()
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_PARITY", "OK" if int(t.item()) == 1 else "FAIL", "world", world, "coarse bytes/rank", prof["bytes_coarse"], flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
