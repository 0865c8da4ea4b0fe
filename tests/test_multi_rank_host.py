"""N>1 host logic on CPU (gloo, world_size 2 and 3): cost-balanced template shard plan + rank-ordered
gather + the product's deterministic merge must reproduce the single-rank result exactly
(SURVEY.md §8e).  Per-rank scoring is stood in by the oracle restricted to the rank's shard; the
shard plan and the merge are the product's own C ABI functions (lmb200_shard_plan / lmb200_merge_matches)."""
import os
import socket
import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, threshold, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import line_mod_pipeline_b200 as lm
    from line_mod_pipeline_b200 import synth
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    bgr, depth = synth.make_frame(0)
    lut = synth.default_normal_lut()
    ora = O.Detector([dict(type=O.CG), dict(type=O.DN)], [5, 8], normal_lut=lut)
    for m in synth.object_masks(0):
        ora.add_template([bgr, depth], "planted", m)
    tps = synth.random_templates(150)
    for tp in tps:
        ora.add_synthetic(tp, "rand")
    # generation order = class order (planted < rand), template id ascending
    n_planted = ora.num_templates("planted")
    n_total = n_planted + len(tps)
    costs = []
    for cid, n in (("planted", n_planted), ("rand", len(tps))):
        for t in range(n):
            tp = O.decode_pyramid(ora.get_template_flat(cid, t))
            c = 0
            for tm in tp[2:]:   # coarsest level, T=8, 40x30 linear memory
                wf, hf = (tm["width"] - 1) // 8 + 1, (tm["height"] - 1) // 8 + 1
                c += len(tm["features"]) * max(0, (30 - hf) * 40 + (40 - wf) + 1)
            costs.append(c + 1.0)
    begin = lm.shard_plan(costs, world)
    res = ora.match([bgr, depth], threshold, threads=2, debug=True)
    gen = res.matches(1)                                   # generation order, pre sort/unique
    gidx = np.where(gen.class_index == 0, gen.template_id, n_planted + gen.template_id)
    mine = gen[(gidx >= begin[rank]) & (gidx < begin[rank + 1])]
    part = np.zeros(len(mine), lm.MATCH_DTYPE)
    for k in ("x", "y", "similarity", "class_index", "template_id"):
        part[k] = mine[k]
    gathered = [None] * world
    dist.all_gather_object(gathered, part.tobytes())
    parts = [np.frombuffer(b, lm.MATCH_DTYPE) for b in gathered]
    merged = lm.merge_matches(parts)
    want = res.matches(0)
    ok = len(merged) == len(want) and all(
        (int(a.x), int(a.y), float(a.similarity), int(a.class_index), int(a.template_id)) ==
        (int(b.x), int(b.y), float(b.similarity), int(b.class_index), int(b.template_id)) for a, b in zip(merged, want))
    q.put((rank, ok, len(merged), begin, n_total))
    dist.barrier()
    dist.destroy_process_group()


def _worker_interleaved(rank, world, port, threshold, q):
    """The production wire format over gloo: interleaved shards (rank r scores selection positions r, r+world, ...), every
    rank packs its generation-order records like gather_pack_kernel, the buffers are all-gathered, and the product's host
    merge (lmb200_debug_merge_gathered = the code behind lmb200_fetch_resident_allgather) must return the single-rank list."""
    import sys
    import ctypes as C
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import line_mod_pipeline_b200 as lm
    from line_mod_pipeline_b200 import synth
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    frames = [synth.make_frame(i) for i in range(2)]
    ora = O.Detector([dict(type=O.CG), dict(type=O.DN)], [5, 8], normal_lut=synth.default_normal_lut())
    for m in synth.object_masks(0):
        ora.add_template(list(frames[0]), "planted", m)
    for tp in synth.random_templates(120):
        ora.add_synthetic(tp, "rand")
    n_planted, n_rand = ora.num_templates("planted"), ora.num_templates("rand")
    ntpl = n_planted + n_rand
    g_class = np.array([0] * n_planted + [1] * n_rand, np.int32)
    g_tid = np.array(list(range(n_planted)) + list(range(n_rand)), np.int32)
    pos_of_g = np.arange(ntpl, dtype=np.int32)            # selection = every class in order: position == global index
    wants, mine = [], []
    for fr in frames:
        res = ora.match(list(fr), threshold, threads=2, debug=True)
        gen = res.matches(1)                               # generation order, pre sort/unique
        gidx = np.where(gen.class_index == 0, gen.template_id, n_planted + gen.template_id).astype(np.int32)
        keep = (gidx % world) == rank
        rec = np.zeros((int(keep.sum()), 4), np.int32)
        rec[:, 0], rec[:, 1], rec[:, 2] = gidx[keep], gen.x[keep], gen.y[keep]
        rec[:, 3] = np.asarray(gen.similarity[keep], np.float32).view(np.int32)
        mine.append(rec)
        wants.append(res.matches(0))
    counts = [None] * world
    dist.all_gather_object(counts, [len(r) for r in mine])
    gcap = max(sum(c) for c in counts)
    nf = len(frames)
    buf = np.zeros((2 * nf + gcap, 4), np.int32)
    off = 0
    for f, rec in enumerate(mine):
        buf[2 * f] = (len(rec), 0, off, 0)
        buf[2 * nf + off: 2 * nf + off + len(rec)] = rec
        off += len(rec)
    gathered = [None] * world
    dist.all_gather_object(gathered, buf.tobytes())
    G = np.frombuffer(b"".join(gathered), np.int32).reshape(world, 2 * nf + gcap, 4).copy()
    out = np.zeros(max(1, sum(sum(c) for c in counts)), lm.MATCH_DTYPE)
    offs = (C.c_size_t * (nf + 1))()
    rc = lm.capi.lib().lmb200_debug_merge_gathered(G.ctypes.data, world, nf, gcap, pos_of_g.ctypes.data, g_class.ctypes.data, g_tid.ctypes.data,
                                                   ntpl, out.ctypes.data_as(C.POINTER(lm.capi.MatchRec)), len(out), offs)
    ok = rc == 0
    n = 0
    for f in range(nf):
        got, want = out[offs[f]:offs[f + 1]], wants[f]
        n += len(got)
        ok = ok and len(got) == len(want) and all(
            (int(a["x"]), int(a["y"]), float(a["similarity"]), int(a["class_index"]), int(a["template_id"])) ==
            (int(b.x), int(b.y), float(b.similarity), int(b.class_index), int(b.template_id)) for a, b in zip(got, want))
    q.put((rank, ok, n))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,threshold", [(2, 70.0), (3, 55.0)])
def test_interleaved_shards_gathered_wire_format_and_host_merge(world, threshold):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_interleaved, args=(r, world, port, threshold, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    for rank, ok, n in out:
        assert ok and n > 0, "rank %d: merged lists differ from the single-rank result" % rank
    assert len({o[2] for o in out}) == 1


@pytest.mark.parametrize("world,threshold", [(2, 80.0), (3, 55.0)])
def test_template_sharded_merge_equals_single_rank(world, threshold):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, threshold, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    for rank, ok, n, begin, n_total in out:
        assert ok, "rank %d: merged list differs from the single-rank result" % rank
        assert n > 0 and begin[0] == 0 and begin[-1] == n_total
    assert len({tuple(o[3]) for o in out}) == 1     # every rank computed the same plan
