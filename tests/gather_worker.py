"""Worker of the multi-process gather tests (launched once per rank by tests/test_abi_and_host.py and
tests/test_gpu_parity.py).  usage: gather_worker.py RANK WORLD PORT cpu|gpu

cpu: gloo group, synthetic tracks, TrackGather's torch.distributed transport on host tensors.
gpu: nccl group, one GPU per rank: a ragged mix of windows is sharded with shard_graphs, every rank builds + solves
     its shard (TrackPipeline), TrackGather brings the tracks to rank 0 (copy-engine puts through the CUDA IPC peer
     buffer), rank 0 compares them with its own solve of ALL graphs. Three steps, so that both slots of the double
     buffer and the flow control are exercised."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mot3d_b200 as mot3d  # noqa: E402

rank, world, port, mode = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]


def fake_tracks(sizes, seed):
    """Random consistent (graph_ptr, count, ptr, det) arrays in the layout of mot3d_extract_tracks."""
    rng = np.random.RandomState(seed)
    graph_ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    n, G = int(graph_ptr[-1]), len(sizes)
    count = np.zeros(G, dtype=np.int32)
    ptr = np.zeros(n + G + 1, dtype=np.int32)
    det = np.zeros(max(n, 1), dtype=np.int32)
    expect = []
    for g, sz in enumerate(sizes):
        g0 = int(graph_ptr[g])
        perm = rng.permutation(sz)
        used = perm[: rng.randint(0, sz + 1)]
        cuts = np.sort(rng.choice(np.arange(1, max(len(used), 1)), size=min(3, max(len(used) - 1, 0)), replace=False)) \
            if len(used) > 1 else np.array([], dtype=np.int64)
        pieces = [p for p in np.split(used, cuts) if len(p)]
        count[g] = len(pieces)
        off = 0
        for k, p in enumerate(pieces):
            ptr[g0 + g + k] = off
            det[g0 + off: g0 + off + len(p)] = p
            off += len(p)
        ptr[g0 + g + len(pieces)] = off
        expect.append([p.tolist() for p in pieces])
    return graph_ptr, count, ptr, det, expect


if mode == "cpu":
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    sizes = [[7, 3, 9], [4, 5, 1, 6]][rank % 2]  # ragged shards
    gp, count, ptr, det, expect = fake_tracks(sizes, 100 + rank)
    assert mot3d.unpack_tracks_host(gp, mot3d.pack_tracks_host(gp, count, ptr, det)) == expect
    tg = mot3d.TrackGather(int(gp[-1]), len(sizes))
    assert tg.transport == "collective"
    for step in range(2):
        tg.gather(gp, count, ptr, det)
    if rank == 0:
        for r in range(world):
            gp_r, _, _, _, expect_r = fake_tracks([[7, 3, 9], [4, 5, 1, 6]][r % 2], 100 + r)
            assert tg.tracks(r, gp_r) == expect_r, r
        print("GATHER_OK")
    dist.destroy_process_group()
    sys.exit(0)

# ------------------------------------------------------------------------------------------------------------ gpu
from mot3d_b200.pipeline import TrackPipeline  # noqa: E402

torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world, device_id=dev)
sys.path.insert(0, ROOT)
from bench import gen_sequence, DIST_KW  # noqa: E402

# a ragged mix: windows of 50 / 20 / 8 frames with 100 / 30 / 8 detections per frame
graphs = []
for k, (frames, per_frame) in enumerate([(50, 100), (20, 30), (8, 8), (50, 100), (20, 30), (8, 8), (30, 60), (50, 100)]):
    f, p = gen_sequence(7 + k, frames=frames, per_frame=per_frame, extent=1000.0 * (per_frame / 100.0) ** 0.5)
    graphs.append((f, p))
sizes = [len(f) for f, _ in graphs]
spec = mot3d.CostSpec(kind="detections_2d", distance=DIST_KW, confidence=dict(mul=1, bias=0), weight_source_sink=1,
                      max_jump=8)


def solve(ids):
    gp = np.concatenate([[0], np.cumsum([sizes[g] for g in ids])]).astype(np.int32)
    frame = np.concatenate([graphs[g][0] for g in ids])
    pos = np.concatenate([graphs[g][1] for g in ids])
    hb, _ = mot3d.DetectionBatch.from_arrays(gp, frame, pos, np.full(len(frame), 0.9))
    db = hb.to(dev)
    gb = mot3d.build_graphs(db, spec)
    pipe = TrackPipeline(len(frame), len(ids), gb.n_edges + 64, device=dev)
    return gp, db, pipe


parts = mot3d.shard_graphs(sizes, world)
mine = parts[rank].tolist()
gp, db, pipe = solve(mine)
tg = mot3d.TrackGather(int(gp[-1]), len(mine), device=dev, transport="p2p" if len(sys.argv) < 6 else sys.argv[5])
for step in range(3):
    pipe.run(db, spec)
    tg.gather(db.graph_ptr, pipe.track_count, pipe.track_ptr, pipe.track_det)
    if rank == 0 and step < 2:
        tg.release()
torch.cuda.synchronize(dev)
assert not tg.timed_out()
if rank == 0:
    gp_all, db_all, pipe_all = solve(list(range(len(graphs))))
    pipe_all.run(db_all, spec)
    want = pipe_all.tracks(gp_all)
    for r in range(world):
        ids = parts[r].tolist()
        gp_r = np.concatenate([[0], np.cumsum([sizes[g] for g in ids])]).astype(np.int32)
        got = tg.tracks(r, gp_r)
        for k, g in enumerate(ids):
            assert got[k] == want[g], (r, g)
    tg.release()
    print("GATHER_OK transport=%s" % tg.transport)
tg.close()
dist.destroy_process_group()
