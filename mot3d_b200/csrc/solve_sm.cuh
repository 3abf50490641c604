// Block-per-component solver on a staging area in SHARED memory: successive shortest paths on one connected
// component. Included by solve.cu.
//
//   staging     the component is copied ONCE into the block's shared memory, laid out for its size: 64-byte records
//               (flow, tree, distances, costs; see SRec), the rank of every record's frame layer, three 16-bit slot
//               lists, the change bits of the relabelling rounds, and its in-arcs as two arrays (32-bit cost, 16-bit
//               source slot: 6 bytes per arc). Every label - parent, branch, next, prev - is a slot of this area.
//               The work queue only hands a block components that fit (staging_bytes); costs fit 32 bits
//               (k_first_tree checked), distances are exact int64.
//   next path   block arg-min over dpost + exit cost (muSSP's sink heap, Graph.cpp:508-541); stop when the path does
//               not pay (wrapper.cpp:263-267). SEVERAL paths are taken from one set of labels: after a path through
//               the branch of S -> pre_r is flipped, every residual arc still has a non-negative reduced cost under
//               the old labels (the reversed arcs are tight), so no label falls, the labels outside that branch stay
//               exact and the ones inside are lower bounds. The next sink in (cost, slot) order is therefore the next
//               shortest path as long as it lies in a branch no path of the batch went through (the vendored
//               wrapper.cpp:237-300 re-labels after every single path; the result is the same, see below); the
//               first sink in a used branch ends the batch, and a next sink that does not pay ends the
//               component without a last re-labelling. The path costs and their order are those of one-by-one
//               successive shortest paths.
//   path update one thread walks the path back from T and rewrites next / prev (Graph::flip_path, Graph.cpp:266-335).
//   certificate once every detection is on a track: if every matched arc is the cheapest in-arc of its head, costs
//               mc <= 0, cn + mc <= 0, and no entry / exit cost is negative, no further path can pay (every piece of
//               a residual S -> T path is >= 0): finished without a last re-labelling.
//   re-label    muSSP's observation (Graph.cpp:341-343, 394-503): only the branch of the shortest-path tree that hung
//               from the path's first arc loses its labels (here: the branches of all paths of the batch, at
//               once). Their nodes start unlabelled and are relaxed in SYNCHRONOUS
//               ROUNDS in pull mode: a round computes, from the labels of the round before, the best label of a
//               record's pre node (in-arcs, S arc, reversed node arc) and post node (own pre node when unused,
//               reversed matched arc of the successor when on a track); all changes are published at once. Labels
//               only fall, the labels outside the branch are exact, so the fix-point is the exact distance of every
//               node, reached after as many rounds as the new shortest paths have records - forward and reversed arcs
//               travel in the same round. A parent is only replaced by a strictly better one read from the round
//               before: the parent graph stays a tree (a cycle of tight arcs would need every arc set strictly later
//               than the one before it) and nothing depends on timing.
//               A record is only looked at in a round when one of its inputs may have moved in the round before: a
//               post label in one of the `span` frame layers below it, or - on a track - a pre label in one of the
//               `span` layers above it (one bit per layer and kind, two buffers each). The records to look at are
//               compacted into a work list, so that a round with a handful of them costs one warp a single pass
//               through the relaxation code (four lanes per record); the new labels are staged next to the list and
//               published after a barrier.
#pragma once

struct __align__(16) SRec {
    int4 A;       // x = in-degree, y = first arc, z = prev, w = next (slots / NONE / TERM)
    int4 F;       // x = parent of the pre node (slot / PAR_*), y = branch of pre, z = branch of post (slot of the
                  // detection whose S -> pre arc starts the node's tree path; muSSP's ancestor_node_id, Graph.cpp:160)
    longlong2 E;  // exact distances from S of pre and post
    int4 W;       // costs: S -> pre, pre -> post, post -> T (NO_ARC32 = no such arc), matched arc prev -> this
};
static_assert(sizeof(SRec) == 64, "record layout");
constexpr int SM_REC_BYTES = sizeof(SRec) + 4 + 3 * 2 + 24;  // + layer rank, two slot lists, branch flags, staged result
constexpr int SM_ARC_BYTES = 4 + 2;
constexpr int SM_MAX_RECORDS = 16000;  // slots travel in 16 bits
__host__ __device__ __forceinline__ long long sm_chg_bytes(long long n8) { return 16 * ((n8 >> 5) + 3); }
// shared-memory footprint of a component of n detections with `arcs` in-arcs
__host__ __device__ __forceinline__ long long staging_bytes(long long n, long long arcs) {
    const long long n8 = (n + 7) & ~7ll, a8 = (arcs + 7) & ~7ll;
    return n8 * SM_REC_BYTES + sm_chg_bytes(n8) + a8 * SM_ARC_BYTES;
}

// s_work: [0] rounds, [2] node pulls, [3] arc pulls, [4] cycles in the rounds, [8] records of the branches,
//         [10 + k] phase cycles (arg-min, branch flags, staging, walk, rounds, -, certificate), [17] arcs staged,
//         [18] branches, [20] certificates
constexpr int S_WORK_LEN = 21;
struct PhaseClock {
    long long *w;
    long long t_mark;
    __device__ __forceinline__ void tick(int k) {
        if (threadIdx.x == 0) {
            const long long now = clock64();
            w[10 + k] += now - t_mark;
            t_mark = now;
        }
    }
    __device__ __forceinline__ void add(int k, long long v) {
        if (threadIdx.x == 0) w[17 + k] += v;
    }
};

// Solve positions [c0, c0 + n) of graph g on the block's shared memory (p.sbytes). Returns false, without side effects,
// when the in-arcs do not fit.
__device__ __noinline__ bool solve_component_sm(const SolveParams &p, const int cg, const int c0, const int n, const int g,
                                                int *s_scan, KeyIdx *s_red, int *s_ok, long long *s_work) {
    const int T = blockDim.x, tid = threadIdx.x;
    const int sub = tid % LPN, grp = tid / LPN, ngrp = T / LPN;
    const unsigned gmask = ((1u << LPN) - 1u) << ((tid & 31) / LPN * LPN);
    const int n8 = (n + 7) & ~7;
    SRec *const rec = (SRec *)smem_raw;
    int32_t *const tk = (int32_t *)(smem_raw + (size_t)n8 * sizeof(SRec));
    longlong2 *const res_e = (longlong2 *)(tk + n8);          // staged results of a round: new labels ...
    short4 *const res_f = (short4 *)(res_e + n8);             // ... parent, branches, what moved (1 = pre, 2 = post)
    unsigned short *const hop = (unsigned short *)(res_f + n8);  // detections whose matched in-arc changed (path walk)
    unsigned short *const pl = hop + n8;                      // work list of a relabelling round
    unsigned char *const flg = (unsigned char *)(pl + n8);    // DL_* of the branch being re-labelled (2 bytes / record)
    unsigned *const chg = (unsigned *)(flg + 2 * n8);         // change bits [post | pre][2][nwords]
    int32_t *const arc_cost = (int32_t *)(smem_raw + (size_t)n8 * SM_REC_BYTES + sm_chg_bytes(n8));
    const int racap = (int)(((long long)p.sbytes - n8 * SM_REC_BYTES - sm_chg_bytes(n8)) / SM_ARC_BYTES) & ~7;
    unsigned short *const arc_src = (unsigned short *)(arc_cost + (racap > 0 ? racap : 0));
    const int32_t *const perm = p.comp_perm + c0;

    // ---- stage the component. Arc offsets and layer ranks by block scans over contiguous chunks of records.
    const int chunk = (n + T - 1) / T;
    const int i0 = min(tid * chunk, n), i1 = min(i0 + chunk, n);
    int my_arcs = 0, my_layers = 0, prev_key = 0;
    bool first_flag = false;
    if (i0 < i1 && i0 > 0) prev_key = p.tkey[perm[i0 - 1]];
    for (int i = i0; i < i1; i++) {
        const int d = perm[i];
        const int key = p.tkey[d];
        my_arcs += p.in_ptr[d + 1] - p.in_ptr[d];
        const bool f = i > 0 && key != prev_key;  // a new frame layer starts here
        if (i == i0) first_flag = f;
        my_layers += f ? 1 : 0;
        prev_key = key;
    }
    int tot, nl;
    int a0 = block_excl_scan1(my_arcs, s_scan, 0, &tot);
    int l0 = block_excl_scan1(my_layers, s_scan, 1, &nl);
    if (tot > racap) {  // (the queue's band was chosen from the same numbers: cannot happen; the caller falls back)
        __syncthreads();
        return false;
    }
    PhaseClock cnt;
    cnt.w = s_work;
    cnt.t_mark = clock64();
    if (tid == 0) s_ok[2] = 0;  // largest layer span of an arc
    // records: every load of a record is issued before any is used (one memory level after the permutation)
    for (int i = i0; i < i1; i++) {
        const int q = c0 + i, d = perm[i];
        const int e0 = p.in_ptr[d], e1 = p.in_ptr[d + 1];
        const FtTmp ft = p.ft[q];
        const int par = ft.par, ap_ = ft.anc, ao_ = ft.anc;
        const int64_t dpre = ft.dpre, dpost = ft.dpost;
        const int64_t en = p.en[d], cn = p.cn[d], ex = p.ex[d];
        const int key = p.tkey[d];
        SRec r;
        r.A = make_int4(e1 - e0, a0, NONE, NONE);
        r.F = make_int4(par >= 0 ? par - c0 : par, ap_ >= 0 ? ap_ - c0 : -1, ao_ >= 0 ? ao_ - c0 : -1, e0);
        r.E = make_longlong2(dpre, dpost);
        r.W = make_int4(en >= INF_CUT ? NO_ARC32 : (int)en, (int)cn, ex >= INF_CUT ? NO_ARC32 : (int)ex, 0);
        rec[i] = r;
        a0 += e1 - e0;
        if (i == i0 ? first_flag : key != prev_key) l0++;
        prev_key = key;
        tk[i] = l0;
    }
    cnt.add(0, tot);
    __syncthreads();
    int span = 0;
    for (int ib = 0; ib < n; ib += ngrp) {  // arc lists: one lane group per record, two independent arcs per lane
        const int i = ib + grp;
        if (i < n) {
            const int e0 = rec[i].F.w, deg = rec[i].A.x, off = rec[i].A.y;
            const int my_layer = tk[i];
            for (int kb = sub; kb < deg; kb += 2 * LPN) {
                const int k1 = kb + LPN;
                const bool two = k1 < deg;
                const int src0 = p.in_src[e0 + kb], src1 = p.in_src[e0 + (two ? k1 : kb)];
                const int64_t w0 = p.in_cost[e0 + kb], w1 = p.in_cost[e0 + (two ? k1 : kb)];
                const int q0 = p.comp_inv[src0] - c0, q1 = p.comp_inv[src1] - c0;
                arc_cost[off + kb] = (int)w0;
                arc_src[off + kb] = (unsigned short)q0;
                span = max(span, my_layer - tk[q0]);
                if (two) {
                    arc_cost[off + k1] = (int)w1;
                    arc_src[off + k1] = (unsigned short)q1;
                    span = max(span, my_layer - tk[q1]);
                }
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) span = max(span, __shfl_xor_sync(0xffffffffu, span, d));
    if (lane_id() == 0 && span > 0) atomicMax(&s_ok[2], span);
    __syncthreads();
    span = s_ok[2];
    cnt.tick(2);

    const int nwords = (n8 >> 5) + 3;
    const unsigned lt = (1u << lane_id()) - 1u;
    int K = 0, n_used = 0;
    int64_t total = 0;
    for (int iter = 0; iter <= n; iter++) {  // at most one path per detection
        // ---- a batch of paths from ONE set of labels (see "next path" in the header): sinks in ascending order, each
        //      in a branch no earlier path of the batch went through. pl[] marks the roots of those branches.
        for (int s = tid; s < n; s += T) pl[s] = 0;
        bool finished = false;
#pragma unroll 1
        for (;;) {
        // ---- cheapest way into T
        KeyIdx best;
        best.key = INF;
        best.idx = -1;
        for (int i = tid; i < n; i += T) {
            const int64_t d = rec[i].E.y;
            const int x = rec[i].W.z;
            if (rec[i].A.w != TERM && d < INF_CUT && x != NO_ARC32) {
                KeyIdx v;
                v.key = d + x;
                v.idx = i;
                best = min_ki(best, v);
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const KeyIdx o = shfl_xor_ki(best, d);
            const bool take = o.key < best.key || (o.key == best.key && (unsigned)o.idx < (unsigned)best.idx);
            best.key = take ? o.key : best.key;
            best.idx = take ? o.idx : best.idx;
        }
        if (lane_id() == 0) s_red[warp_id()] = best;
        __syncthreads();
        KeyIdx sink;
        sink.key = INF;
        sink.idx = -1;
        for (int k = 0; k < (T + 31) >> 5; k++) {
            const KeyIdx o = s_red[k];
            if (o.key < sink.key || (o.key == sink.key && (unsigned)o.idx < (unsigned)sink.idx)) sink = o;
        }
        // (read before the barrier: thread 0 marks the root right after it)
        const int root = (sink.idx >= 0 && sink.key < 0) ? rec[sink.idx].F.z : -1;
        const bool root_used = root >= 0 && pl[root] != 0;
        __syncthreads();  // s_red is rewritten next iteration
        cnt.tick(0);
        finished = true;
        if (sink.idx < 0) break;  // T unreachable
        if (sink.key >= 0) {      // no sink pays, now or (the stale labels are lower bounds) after a re-labelling.
            if (K == 0 && tid == 0) {  // the first-path rule belongs to the whole graph (k_first_path_rule)
                p.first_cost[cg] = sink.key;
                p.first_sink[cg] = c0 + sink.idx;
            }
            break;
        }
        if (root < 0) break;  // cannot happen: a labelled node always has a branch
        finished = false;
        if (root_used) break;  // its branch already carries a path of this batch: its labels are stale, re-label first
        // ---- augment: walk the path back from T and rewrite next / prev
        if (tid == 0) {
            pl[root] = 1;
            int ok = 0, nhop = 0, nxt_new = TERM, cur = sink.idx, joined = 0;
            for (int step = 0; step <= n; step++) {
                const int4 A = rec[cur].A;         // z = prev, w = next
                const int par_cur = rec[cur].F.x;  // in flight together: the parent of pre_cur (used when cur joins)
                rec[cur].A.w = nxt_new;
                int m, q;
                if (A.z == NONE) {  // post_cur came from pre_cur: detection cur joins a track
                    m = cur;
                    q = par_cur;
                    joined++;
                } else {  // post_cur came from pre_m, m = its old successor: the arc cur -> m is cancelled
                    m = A.w;
                    q = PAR_NONE;
                    while (m >= 0) {
                        q = rec[m].F.x;
                        if (q != PAR_POST) break;
                        const int m2 = rec[m].A.w;  // pre_m came from post_m: m leaves its track
                        rec[m].A.w = NONE;
                        rec[m].A.z = NONE;
                        m = m2;
                        joined--;
                    }
                }
                if (m < 0) break;  // cannot happen on a consistent residual state
                if (q < 0) {       // PAR_S: the path starts here
                    rec[m].A.z = TERM;
                    ok = 1;
                    break;
                }
                rec[m].A.z = q;
                hop[nhop++] = (unsigned short)m;  // the cost of the newly matched arc q -> m is looked up below
                nxt_new = m;
                cur = q;
            }
            s_ok[0] = ok;
            s_ok[1] = nhop;
            s_ok[7] = joined;
        }
        __syncthreads();
        cnt.tick(3);
        if (!s_ok[0]) {  // defensive: inconsistent state, stop with what we have
            finished = true;
            break;
        }
        n_used += s_ok[7];
        for (int h = tid; h < s_ok[1]; h += T) {  // cost of every newly matched arc prev -> m
            const int m = hop[h], q = rec[m].A.z;
            const int off = rec[m].A.y, deg = rec[m].A.x;
            for (int k = 0; k < deg; k++)
                if ((int)arc_src[off + k] == q) {
                    rec[m].W.w = arc_cost[off + k];
                    break;
                }
        }
        if (tid == 0 && p.path_cost) p.path_cost[c0 + K] = sink.key;
        K++;
        total += sink.key;
        __syncthreads();
        }  // (batch)
        if (finished) break;
        // ---- finished? (every detection on a track + the certificate of the header)
        if (n_used == n) {
            bool good = true;
            for (int ib = 0; ib < n; ib += ngrp) {
                const int i = ib + grp;
                if (i < n) {
                    const int4 A = rec[i].A;
                    const int4 W = rec[i].W;
                    if ((W.x != NO_ARC32 && W.x < 0) || (W.z != NO_ARC32 && W.z < 0)) good = false;
                    if (A.z >= 0) {
                        const int mc = W.w;
                        if (mc > 0 || (int64_t)W.y + mc > 0) good = false;
                        for (int k = sub; k < A.x; k += LPN)
                            if ((int)arc_src[A.y + k] != A.z && arc_cost[A.y + k] < mc) good = false;
                    }
                }
            }
            const bool done = __syncthreads_and(good);
            cnt.tick(6);
            if (done) {
                cnt.add(3, 1);
                break;
            }
        }
        // ---- the branches hanging from the S -> pre_root arcs of the batch lose their labels
        int n_mine = 0;
        for (int s = tid; s < n; s += T) {
            int4 F = rec[s].F;
            // 1 = pre node, 2 = post node in one of the branches
            const int f = ((F.y >= 0 && pl[F.y]) ? 1 : 0) | ((F.z >= 0 && pl[F.z]) ? 2 : 0);
            flg[s] = (unsigned char)f;
            if (f) {
                longlong2 E = rec[s].E;
                if (f & 1) {
                    E.x = INF;
                    F.x = PAR_NONE;
                    F.y = -1;
                }
                if (f & 2) {
                    E.y = INF;
                    F.z = -1;
                }
                rec[s].E = E;
                rec[s].F = F;
                n_mine++;
            }
        }
        for (int w = tid; w < 4 * nwords; w += T) chg[w] = 0;
        if (tid == 0) s_ok[3] = 0;  // length of the round's work list
        cnt.add(1, 1);
        __syncthreads();
        cnt.tick(1);
        int n_arc_pull = 0, n_relax = 0, rounds = 0;
#pragma unroll 1
        for (; rounds < 4 * n + 8; rounds++) {
            const int rd = rounds & 1;
            const unsigned *const post_rd = chg + rd * nwords, *const pre_rd = chg + (2 + rd) * nwords;
            unsigned *const post_wr = chg + (rd ^ 1) * nwords, *const pre_wr = chg + (2 + (rd ^ 1)) * nwords;
#pragma unroll 1
            for (int base = 0; base < n; base += T) {  // (A) which records may have an input that moved?
                const int s = base + tid;
                const int f = s < n ? flg[s] : 0;
                bool dirty = false;
                if (f) {
                    dirty = rounds == 0 || span > 32;
                    if (!dirty) {
                        const int l = tk[s];
                        if (f & 1) {  // a post label in one of the `span` layers below
                            const int lo = max(l - span, 0), width = l - lo;
                            const unsigned long long two =
                                (unsigned long long)post_rd[lo >> 5] | ((unsigned long long)post_rd[(lo >> 5) + 1] << 32);
                            dirty = width > 0 && ((two >> (lo & 31)) & ((1ull << width) - 1ull)) != 0ull;
                        }
                        if (!dirty && (f & 2) && rec[s].A.w >= 0) {  // on a track: a pre label in the `span` layers above
                            const int lo = l + 1;
                            const unsigned long long two =
                                (unsigned long long)pre_rd[lo >> 5] | ((unsigned long long)pre_rd[(lo >> 5) + 1] << 32);
                            dirty = ((two >> (lo & 31)) & ((1ull << span) - 1ull)) != 0ull;
                        }
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, dirty);
                if (m) {
                    int at = 0;
                    if (lane_id() == 0) at = atomicAdd(&s_ok[3], __popc(m));
                    at = __shfl_sync(0xffffffffu, at, 0);
                    if (dirty) pl[at + __popc(m & lt)] = (unsigned short)s;
                }
            }
            __syncthreads();
            const int n_work = s_ok[3];
            if (n_work == 0) break;  // nothing moved in the round before: fix-point
#pragma unroll 1
            for (int wi = grp; wi < n_work; wi += ngrp) {  // (B) relax them from the labels of the round before, four
                const int s = pl[wi];                         //     lanes per record; the results are staged
                const int f = flg[s];
                const int4 A = rec[s].A;  // x = in-degree, y = first arc, z = prev, w = next
                const int4 W = rec[s].W;
                const longlong2 E = rec[s].E;
                const int4 F = rec[s].F;
                int64_t dpre = E.x, dpost = E.y;
                int par = F.x, apre = F.y, apost = F.z;
                if ((f & 2) && A.w >= 0) {  // on a track: reversed matched arc pre_next -> post_s
                    const int64_t dm = rec[A.w].E.x;
                    if (dm < INF_CUT && dm - (int64_t)rec[A.w].W.w < dpost) {
                        dpost = dm - (int64_t)rec[A.w].W.w;
                        apost = rec[A.w].F.y;
                    }
                }
                if (f & 1) {
                    const int32_t *const ac = arc_cost + A.y;
                    const unsigned short *const as = arc_src + A.y;
                    int64_t key = INF;
                    int kpar = 0x7fffffff;
#pragma unroll 1
                    for (int k = sub; k < A.x; k += 2 * LPN) {  // two independent arc / label loads in flight
                        const int k1 = k + LPN < A.x ? k + LPN : k;
                        const int s0 = as[k], s1 = as[k1];
                        const int64_t c0_ = rec[s0].E.y + (int64_t)ac[k], c1_ = rec[s1].E.y + (int64_t)ac[k1];
                        // (an unlabelled source gives INF + cost: it loses against every labelled one, see below)
                        if (s0 != A.z && (c0_ < key || (c0_ == key && s0 < kpar))) {  // lowest slot wins ties
                            key = c0_;
                            kpar = s0;
                        }
                        if (s1 != A.z && (c1_ < key || (c1_ == key && s1 < kpar))) {
                            key = c1_;
                            kpar = s1;
                        }
                    }
#pragma unroll
                    for (int d = 1; d < LPN; d <<= 1) {
                        const int64_t ok_ = __shfl_xor_sync(gmask, key, d);
                        const int op = __shfl_xor_sync(gmask, kpar, d);
                        if (ok_ < key || (ok_ == key && op < kpar)) {
                            key = ok_;
                            kpar = op;
                        }
                    }
                    if (sub == 0) n_arc_pull += A.x;
                    int kanc = -1;
                    if (key < INF_CUT) {
                        kanc = rec[kpar].F.z;
                    } else {
                        key = INF;
                        kpar = PAR_NONE;
                    }
                    if (A.z != TERM && W.x != NO_ARC32 && (int64_t)W.x < key) {
                        key = W.x;
                        kpar = PAR_S;
                        kanc = s;
                    }
                    if (A.z != NONE && dpost < INF_CUT && dpost - (int64_t)W.y < key) {  // reversed node arc
                        key = dpost - (int64_t)W.y;
                        kpar = PAR_POST;
                        kanc = apost;
                    }
                    if (key < dpre) {
                        dpre = key;
                        par = kpar;
                        apre = kanc;
                    }
                }
                if ((f & 2) && A.z == NONE && dpre < INF_CUT && dpre + (int64_t)W.y < dpost) {
                    dpost = dpre + (int64_t)W.y;  // unused record: post hangs from the own pre node
                    apost = apre;
                }
                if (sub == 0) {
                    const int upd = (dpre != E.x ? 1 : 0) | (dpost != E.y ? 2 : 0);
                    if (upd) res_e[wi] = make_longlong2(dpre, dpost);
                    res_f[wi] = make_short4((short)par, (short)apre, (short)apost, (short)upd);
                    n_relax++;
                }
            }
            __syncthreads();  // every label of the round before has been read
            for (int w = tid; w < nwords; w += T) {  // the next round's write buffers
                chg[rd * nwords + w] = 0;
                chg[(2 + rd) * nwords + w] = 0;
            }
            if (tid == 0) s_ok[3] = 0;
#pragma unroll 1
            for (int wi = tid; wi < n_work; wi += T) {  // (C) publish
                const short4 rf = res_f[wi];
                if (rf.w) {
                    const int s = pl[wi];
                    rec[s].E = res_e[wi];
                    int4 F = rec[s].F;
                    F.x = rf.x;
                    F.y = rf.y;
                    F.z = rf.z;
                    rec[s].F = F;
                    const int l = tk[s];
                    if (rf.w & 1) atomicOr(&pre_wr[l >> 5], 1u << (l & 31));
                    if (rf.w & 2) atomicOr(&post_wr[l >> 5], 1u << (l & 31));
                }
            }
            __syncthreads();
        }
        if (p.stats) {  // work counters of this branch
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                n_arc_pull += __shfl_xor_sync(0xffffffffu, n_arc_pull, d);
                n_relax += __shfl_xor_sync(0xffffffffu, n_relax, d);
                n_mine += __shfl_xor_sync(0xffffffffu, n_mine, d);
            }
            if (lane_id() == 0) {
                atomicAdd((unsigned long long *)&s_work[2], (unsigned long long)(2 * n_relax));
                atomicAdd((unsigned long long *)&s_work[3], (unsigned long long)n_arc_pull);
                atomicAdd((unsigned long long *)&s_work[8], (unsigned long long)n_mine);
            }
            if (tid == 0) s_work[0] += rounds + 1;
        }
        __syncthreads();
        cnt.tick(4);
    }
    // ---- results of this component: flow in detection indices, totals added to its graph
    __syncthreads();
    if (tid == 0) {
        if (K) {
            atomicAdd(&p.n_paths[g], K);
            atomicAdd((unsigned long long *)&p.total[g], (unsigned long long)total);
        }
        if (p.stats) {
            unsigned long long *o = (unsigned long long *)(p.stats + (size_t)MOT3D_SOLVE_STATS * g);
            atomicAdd(&o[0], (unsigned long long)s_work[0]);
            atomicAdd(&o[2], (unsigned long long)s_work[2]);
            atomicAdd(&o[3], (unsigned long long)(s_work[17] + s_work[3]));
            for (int k = 0; k < 6; k++) atomicAdd(&o[4 + k], (unsigned long long)s_work[10 + k]);
            atomicAdd(&o[9], (unsigned long long)s_work[16]);
            atomicAdd(&o[10], (unsigned long long)s_work[20]);
            atomicAdd(&o[11], (unsigned long long)s_work[18]);
            atomicAdd(&o[12], (unsigned long long)s_work[8]);
            atomicAdd(&o[13], (unsigned long long)s_work[14]);
            atomicAdd(&o[15], (unsigned long long)s_work[0]);
        }
    }
    for (int i = tid; i < n; i += T) {
        const int a = rec[i].A.w, b = rec[i].A.z;
        const int d = perm[i];
        p.next_out[d] = a >= 0 ? perm[a] : a;
        p.prev_out[d] = b >= 0 ? perm[b] : b;
    }
    __syncthreads();
    return true;
}
