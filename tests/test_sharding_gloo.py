"""The N>1 path on CPU: world_size-2 (and 3) gloo groups run the shard planner, each rank
processes only its own shards, and the gathered sound units / PCM must equal the unsharded
result.  The GPU kernels cannot run here, so each rank's "device" is the oracle's range
functions, cold-started exactly `enc_halo` frames (decode: `dec_halo` units) before the shard --
i.e. this checks the halo arithmetic the GPU shards rely on plus the
planner, the disjoint-output layout and the max-over-ranks reduction bench.py uses."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, stream_seconds, result_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import signals as S
    from carta1_b200 import sharding
    from oracle import oracle as O

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        streams = [S.cfg3_transients(sec, seed=100 + i, n_ch=2) for i, sec in enumerate(stream_seconds)]
        frames = [O.frame_count(len(ch[0])) for ch in streams]
        plan = sharding.plan(frames, world)
        sharding.check_plan(plan, frames)
        opts = O.make_options()
        tables = O.default_tables()
        # every rank writes its disjoint slices into zero-initialised full-size outputs
        su_out = [np.zeros((f * 2, 212), np.uint8) for f in frames]
        pcm_out = [np.zeros((2, f * 512), np.float32) for f in frames]
        work = 0
        for sh in plan[rank]:
            ch = streams[sh.stream]
            first, last = sh.pcm_span()
            # encode: the rank only sees PCM from `first` on (enc_halo frames of history)
            local = [np.ascontiguousarray(c[first:last]) for c in ch]
            su = O.encode_pcm(local, opts, tables)[sh.enc_halo * 2:]  # cold start at `first`, halo outputs dropped
            su_out[sh.stream][sh.begin * 2:sh.end * 2] = su
            work += sh.frames
        # decode needs the complete units: gather (SUM of disjoint slices == concatenation)
        for i in range(len(frames)):
            t = torch.from_numpy(su_out[i].astype(np.int32))
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            su_out[i] = t.numpy().astype(np.uint8)
        for sh in plan[rank]:
            ua, ub = sh.unit_span(2)
            pcm = O.decode_su(np.ascontiguousarray(su_out[sh.stream][ua:ub]), 2, tables)
            for c in range(2):
                pcm_out[sh.stream][c, sh.begin * 512:sh.end * 512] = pcm[c][sh.dec_halo * 512:]
        for i in range(len(frames)):
            t = torch.from_numpy(pcm_out[i].view(np.int32).astype(np.int64))
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            pcm_out[i] = t.numpy().astype(np.int32).view(np.float32)
        # bench.py's timing reduction: max over ranks
        tmax = torch.tensor([float(work)], dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        if rank == 0:
            ok = True
            for i, ch in enumerate(streams):
                want = O.encode_pcm(ch, opts, tables)
                ok &= bool(np.array_equal(su_out[i], want))
                ref = O.decode_su(want, 2, tables)
                ok &= all(np.array_equal(pcm_out[i][c].view(np.uint32), ref[c].view(np.uint32)) for c in range(2))
            result_q.put((ok, [[(s.stream, s.begin, s.end) for s in p] for p in plan], float(tmax.item()), sum(frames)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,stream_seconds", [(2, [0.5]), (2, [0.25, 0.4, 0.1]), (3, [0.3, 0.3])])
def test_sharded_ranks_reproduce_the_unsharded_result(oracle, world, stream_seconds):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, stream_seconds, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, plan, tmax, total = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok, plan
    assert len(plan) == world and all(len(p) > 0 for p in plan)
    assert total / world <= tmax <= total / world + 4  # balanced to within the cut-snapping slack


def test_plan_properties():
    from carta1_b200 import sharding

    for frames, world in (([310079], 8), ([5168] * 7, 4), ([1, 1, 1, 3], 2), ([0, 5], 3), ([2, 2], 8), ([], 2)):
        plan = sharding.plan(frames, world)
        assert len(plan) == world
        sharding.check_plan(plan, frames)
        for p in plan:
            for sh in p:
                assert sh.begin != 1 and (sh.enc_halo, sh.dec_halo) == ((2, 1) if sh.begin else (0, 0))
                assert sh.pcm_span()[0] >= 0 and sh.unit_span(2)[0] >= 0
    big = sharding.plan([310079] * 100, 8)  # config 5: 100 h stereo over 8 GPUs
    loads = [sum(s.frames for s in p) for p in big]
    assert max(loads) - min(loads) <= 2
    with pytest.raises(ValueError):
        sharding.plan([4], 0)
