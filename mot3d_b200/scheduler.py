"""Batch scheduler: whole graphs are the only unit of parallelism across GPUs.

Independent windows / camera views / sequences are assigned to ranks with a longest-processing-time rule on an
estimate of their solve work; every rank builds and solves its own graphs with no data-path collective. The only
communication is one gather of the (ragged) track outputs to rank 0 -- NCCL on GPUs, gloo in the CPU tests.
"""
import numpy as np


def shard_graphs(sizes, world_size, cost=None):
    """Longest-processing-time assignment. sizes[g] = detections of graph g; cost[g] defaults to sizes[g]**1.5
    (arcs grow ~linearly and augmentations ~sqrt with the graph). Returns a list of index arrays, one per rank,
    each ascending, deterministic."""
    sizes = np.asarray(sizes, dtype=np.float64)
    if cost is None:
        cost = sizes ** 1.5
    cost = np.asarray(cost, dtype=np.float64)
    order = np.lexsort((np.arange(len(cost)), -cost))  # heaviest first, ties by index
    loads = np.zeros(world_size)
    owner = np.empty(len(cost), dtype=np.int64)
    for g in order.tolist():
        r = int(np.argmin(loads))  # first minimum -> deterministic
        owner[g] = r
        loads[r] += cost[g]
    return [np.flatnonzero(owner == r) for r in range(world_size)]


def gather_tracks(local_graph_ids, local_tracks, local_costs, n_graphs_total, group=None):
    """Gather per-graph results to rank 0 with torch.distributed (all ranks must call).

    local_graph_ids : global ids of the graphs this rank solved
    local_tracks    : per local graph, a list of tracks (lists of detection indices local to the graph)
    local_costs     : per local graph, total cost (int)
    Returns on rank 0: (tracks per global graph id, costs per global graph id); None elsewhere.
    The payload is flattened to one int64 tensor per rank: [G_r, (gid, cost, K, len_0..len_K-1, dets...)*]."""
    import torch
    import torch.distributed as dist
    flat = [len(local_graph_ids)]
    for gid, tracks, cost in zip(local_graph_ids, local_tracks, local_costs):
        flat += [int(gid), int(cost), len(tracks)]
        flat += [len(t) for t in tracks]
        for t in tracks:
            flat += [int(x) for x in t]
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    payload = torch.tensor(flat, dtype=torch.int64, device=dev)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lens = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(lens, torch.tensor([payload.numel()], dtype=torch.int64, device=dev), group=group)
    max_len = int(max(int(x.item()) for x in lens))
    padded = torch.zeros(max_len, dtype=torch.int64, device=dev)
    padded[:payload.numel()] = payload
    bufs = [torch.zeros(max_len, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    if rank != 0:
        return None
    tracks_out = [None] * n_graphs_total
    costs_out = [None] * n_graphs_total
    for r in range(world):
        a = bufs[r][:int(lens[r].item())].cpu().numpy()
        p = 1
        for _ in range(int(a[0])):
            gid, cost, K = int(a[p]), int(a[p + 1]), int(a[p + 2])
            p += 3
            ls = a[p:p + K].tolist()
            p += K
            tr = []
            for l in ls:
                tr.append(a[p:p + l].tolist())
                p += l
            tracks_out[gid] = tr
            costs_out[gid] = cost
    return tracks_out, costs_out
