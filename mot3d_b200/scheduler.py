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


# ----------------------------------------------------------------------------------------------------------------
# Device-resident gather: one fixed-size int32 payload per rank (include/mot3d_b200.h, "track gather"), put into rank
# 0's receive slots by the copy engines (CUDA IPC peer buffer). torch.distributed only carries the 64-byte handle and
# the shapes, once. NCCL / gloo gather of the same payload is the fallback (no peer access; CPU tests).

def pack_tracks_host(graph_ptr, track_count, track_ptr, track_det):
    """numpy mirror of mot3d_pack_tracks: [K_g | L_g | track_det with MOT3D_TRACK_START_BIT on every track's first
    detection] as int32 (the bit pattern of the device payload)."""
    graph_ptr = np.asarray(graph_ptr, dtype=np.int64)
    G, n = len(graph_ptr) - 1, int(graph_ptr[-1])
    out = np.zeros(2 * G + n, dtype=np.uint32)
    for g in range(G):
        g0, K = int(graph_ptr[g]), int(track_count[g])
        tp = np.asarray(track_ptr[g0 + g: g0 + g + K + 1], dtype=np.int64)
        L = int(tp[K]) if K > 0 else 0
        out[g], out[G + g] = K, L
        det = out[2 * G + g0: 2 * G + g0 + L]
        det[:] = np.asarray(track_det[g0: g0 + L], dtype=np.uint32)
        det[tp[:K]] |= np.uint32(0x80000000)
    return out.view(np.int32)


def unpack_tracks_host(graph_ptr, payload):
    """numpy mirror of mot3d_unpack_tracks: payload -> per graph list of tracks (lists of local detection indices)."""
    graph_ptr = np.asarray(graph_ptr, dtype=np.int64)
    G = len(graph_ptr) - 1
    p = np.asarray(payload).view(np.uint32)
    out = []
    for g in range(G):
        g0, K, L = int(graph_ptr[g]), int(p[g]), int(p[G + g])
        det = p[2 * G + g0: 2 * G + g0 + L]
        starts = np.flatnonzero(det & np.uint32(0x80000000))
        assert len(starts) == K, "corrupt track payload"
        ids = (det & np.uint32(0x7fffffff)).astype(np.int64)
        bounds = list(starts) + [L]
        out.append([ids[bounds[k]:bounds[k + 1]].tolist() for k in range(K)])
    return out


class _DevMem(object):
    """Zero-copy torch view of device memory owned by the library (peer buffers)."""

    def __init__(self, ptr, words):
        self.__cuda_array_interface__ = {"shape": (int(words),), "typestr": "<i4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class TrackGather(object):
    """Track outputs of every rank -> rank 0, every step (all ranks construct it and call gather()).

        tg = TrackGather(n, G)                      # collective: shapes, peer buffer, handle exchange
        tg.gather(graph_ptr, count, ptr, det)       # every step, on the current stream; asynchronous
        tg.payload(r) / tg.tracks(r, graph_ptr_r)   # rank 0: what rank r sent in the last completed step

    transport "p2p": rank 0 owns two receive slots per rank (step parity) in a CUDA IPC peer buffer; a sender packs
    its payload (mot3d_pack_tracks) and puts it there with one device-to-device copy executed by the copy engines,
    then the step number into the slot's flag; rank 0 waits for the flags on its stream (mot3d_peer_wait) and, when
    its consumers of the step are enqueued, publishes "consumed" (release()), which a sender checks before it
    overwrites the slot two steps later. No SM takes part in the transfer.
    transport "collective": torch.distributed gather of the same payload (NCCL on GPUs, gloo on CPU tensors)."""

    FLAG_STRIDE = 32  # words between the flags of two ranks (one 128-byte line each)

    def __init__(self, n, n_graphs, group=None, device=None, transport="auto"):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n, self.G = int(n), int(n_graphs)
        backend = dist.get_backend(group) if dist.is_initialized() else None
        self.on_gpu = backend != "gloo" and torch.cuda.is_available()
        self.device = (torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)) \
            if self.on_gpu else torch.device("cpu")
        # shapes of every rank (ragged shards): rank 0 lays the slots out for the largest
        shp = torch.tensor([self.n, self.G], dtype=torch.int64, device=self.device)
        if self.world > 1:
            all_shp = [torch.zeros(2, dtype=torch.int64, device=self.device) for _ in range(self.world)]
            dist.all_gather(all_shp, shp, group=group)
            self.shapes = [(int(t[0]), int(t[1])) for t in all_shp]
        else:
            self.shapes = [(self.n, self.G)]
        self.words = [2 * g + n_ for n_, g in self.shapes]
        self.slot_words = (max(self.words) + 63) // 64 * 64
        self.step = 0
        self.transport = "collective"
        self._peer = None
        if self.on_gpu:
            from . import _lib
            self._lib = _lib
            self.local = torch.zeros(max(self.words[self.rank], 1), dtype=torch.int32, device=self.device)
            self.scratch = torch.zeros(8, dtype=torch.int32, device=self.device)  # [0] flag source, [1] time-out flag
            if transport in ("auto", "p2p") and self.world > 1:
                self._setup_p2p(strict=(transport == "p2p"))
        else:
            self.local = None
        if self.transport == "collective" and self.rank == 0:
            self.recv = [torch.zeros(max(w, 1), dtype=torch.int32, device=self.device) for w in self.words]

    # -------------------------------------------------------------------------------------------------- p2p set-up
    def _setup_p2p(self, strict):
        import ctypes
        torch, dist, L = self.torch, self.dist, self._lib.lib()
        ctrl_words = self.FLAG_STRIDE * (self.world + 1)  # flags of the ranks, then the "consumed" counter
        total_bytes = 4 * (ctrl_words + 2 * self.world * self.slot_words)
        handle = ctypes.create_string_buffer(self._lib.PEER_HANDLE_BYTES)
        ptr = ctypes.c_void_p()
        ok = 1
        box = [None]
        if self.rank == 0:
            rc = L.mot3d_peer_buffer_create(total_bytes, ctypes.byref(ptr), handle)
            ok = 1 if rc == 0 else 0
            box = [handle.raw if ok else None]
        if self.world > 1:
            dist.broadcast_object_list(box, src=0, group=self.group)
        if box[0] is None:
            ok = 0
        elif self.rank != 0:
            rc = L.mot3d_peer_buffer_open(box[0], ctypes.byref(ptr))
            ok = 1 if rc == 0 else 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:  # no peer access between some pair: fall back together
            if ptr.value:
                (L.mot3d_peer_buffer_destroy if self.rank == 0 else L.mot3d_peer_buffer_close)(ptr)
            if strict:
                raise RuntimeError("TrackGather: CUDA IPC peer buffer unavailable: " + L.mot3d_last_error().decode())
            return
        self._peer = ptr.value
        self._ctrl_words = ctrl_words
        self.transport = "p2p"
        if self.rank == 0:
            self.peer_view = torch.as_tensor(_DevMem(self._peer, total_bytes // 4), device=self.device)
        dist.barrier(group=self.group)

    def _slot_ptr(self, rank, step):
        return self._peer + 4 * (self._ctrl_words + ((step & 1) * self.world + rank) * self.slot_words)

    # -------------------------------------------------------------------------------------------------- every step
    def gather(self, graph_ptr, track_count, track_ptr, track_det):
        """Pack this rank's tracks and send them to rank 0. Device tensors (host arrays in a gloo group).
        Asynchronous on GPUs: everything is enqueued on the current stream."""
        torch, dist = self.torch, self.dist
        self.step += 1
        if not self.on_gpu:
            payload = torch.from_numpy(pack_tracks_host(graph_ptr, track_count, track_ptr, track_det).copy())
            if self.world == 1:
                self.recv = [payload]
            else:
                dist.gather(payload, self.recv if self.rank == 0 else None, dst=0, group=self.group) \
                    if len(set(self.words)) == 1 else self._gather_ragged(payload)
            return
        import ctypes
        from .batch import _ptr, _stream
        L, lib = self._lib.lib(), self._lib
        st = _stream(self.device)
        if self.transport == "p2p":
            dst = ctypes.c_void_p(self._slot_ptr(self.rank, self.step)) if self.rank == 0 else _ptr(self.local)
            lib.check(L.mot3d_pack_tracks(self.G, _ptr(graph_ptr), _ptr(track_count), _ptr(track_ptr), _ptr(track_det),
                                          dst, st))
            if self.rank == 0:
                lib.check(L.mot3d_peer_wait(ctypes.c_void_p(self._peer), self.FLAG_STRIDE, self.world, self.step,
                                            ctypes.c_void_p(self.scratch.data_ptr() + 4), st))
            else:
                if self.step > 2:  # the slot of step - 2 must have been consumed
                    lib.check(L.mot3d_peer_wait_word(ctypes.c_void_p(self._peer + 4 * self.FLAG_STRIDE * self.world),
                                                     self.step - 2, ctypes.c_void_p(self.scratch.data_ptr() + 4), st))
                lib.check(L.mot3d_peer_put(ctypes.c_void_p(self._slot_ptr(self.rank, self.step)), _ptr(self.local),
                                           4 * self.words[self.rank],
                                           ctypes.c_void_p(self._peer + 4 * self.FLAG_STRIDE * self.rank),
                                           _ptr(self.scratch), self.step, st))
            return
        lib.check(L.mot3d_pack_tracks(self.G, _ptr(graph_ptr), _ptr(track_count), _ptr(track_ptr), _ptr(track_det),
                                      _ptr(self.local), st))
        if self.world == 1:
            self.recv = [self.local]
        elif len(set(self.words)) == 1:
            dist.gather(self.local, self.recv if self.rank == 0 else None, dst=0, group=self.group)
        else:
            self._gather_ragged(self.local)

    def _gather_ragged(self, payload):
        """Shards of different sizes: point-to-point sends to rank 0 (NCCL / gloo)."""
        dist = self.dist
        if self.rank == 0:
            self.recv[0].copy_(payload)
            reqs = [dist.irecv(self.recv[r], src=r, group=self.group) for r in range(1, self.world)]
            for q in reqs:
                q.wait()
        else:
            dist.send(payload, dst=0, group=self.group)

    def release(self):
        """Rank 0, p2p transport: the consumers of the last gathered step have been enqueued on the current stream;
        senders may overwrite its slots (two steps from now)."""
        if self.transport == "p2p" and self.rank == 0:
            import ctypes
            from .batch import _stream
            self._lib.check(self._lib.lib().mot3d_peer_set_word(
                ctypes.c_void_p(self._peer + 4 * self.FLAG_STRIDE * self.world), self.step, _stream(self.device)))

    def timed_out(self):
        """A peer did not show up within ~2 s of a wait (synchronises)."""
        return self.on_gpu and int(self.scratch[1].item()) != 0

    # -------------------------------------------------------------------------------------------------- rank 0
    def payload(self, rank):
        """Rank 0: int32 tensor with the payload rank `rank` sent in the last step (a view; valid until release())."""
        assert self.rank == 0
        if self.transport == "p2p":
            off = self._ctrl_words + ((self.step & 1) * self.world + rank) * self.slot_words
            return self.peer_view[off: off + self.words[rank]]
        return self.recv[rank][: self.words[rank]]

    def tracks(self, rank, graph_ptr):
        """Rank 0: per graph of rank `rank` the list of tracks (synchronises and copies to the host)."""
        return unpack_tracks_host(graph_ptr, self.payload(rank).cpu().numpy())

    def close(self):
        if self._peer:
            import ctypes
            L = self._lib.lib()
            self.torch.cuda.synchronize(self.device)
            if self.world > 1:
                self.dist.barrier(group=self.group)
            if self.rank != 0:
                L.mot3d_peer_buffer_close(ctypes.c_void_p(self._peer))
            if self.world > 1:
                self.dist.barrier(group=self.group)  # the owner frees the buffer after every importer has closed it
            if self.rank == 0:
                self.peer_view = None
                L.mot3d_peer_buffer_destroy(ctypes.c_void_p(self._peer))
            self._peer = None
