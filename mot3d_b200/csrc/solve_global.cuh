// Block-per-component solver on a staging area in GLOBAL memory: components that fit no shared-memory staging area of
// solve_sm.cuh or whose costs need more than 32 bits. Any size, any cost: 80-byte records, 16-byte arcs, 64-bit costs,
// 32-bit slots. Included by solve.cu.
//
// After a path through S -> pre_a is flipped (several paths per re-labelling as long as their branches differ, see
// solve_sm.cuh), the branch of the shortest-path tree that hung from that arc is listed in frame-layer order and re-labelled by an exact Bellman-Ford fix-point in pull mode: ascending layer order settles the
// forward arcs (one warp, layer by layer, for the records that pass a fresh label on; everything else at once), a
// descending pass settles the reversed arcs (independent chain walks down the tracks); the two alternate until nothing
// moves (Graph::update_shortest_path_tree_recursive, Graph.cpp:394-503, without heap or tolerances). Work is
// proportional to the branch, which is what a 100 000-detection component needs.
#pragma once

constexpr int DL_FIRST = 1 << 28;  // first record of its frame layer in the branch list
struct __align__(16) Cost4 {
    int64_t x, y, z, w;
};
struct __align__(16) StArc {  // the first arc of a list also carries the length of the list in pad
    long long cost;
    int src;
    int pad;
};
struct GStaging {
    typedef Cost4 costs_t;
    typedef int64_t cost_t;
    typedef StArc arc_t;
    typedef int slot_t;
    static __device__ __forceinline__ int64_t conv(int64_t v) { return v; }
    static __device__ __forceinline__ bool has(int64_t v) { return v < INF_CUT; }
};
// A record, in 16-byte groups so that a relaxation is a handful of 128-bit loads:
//   A  x = DL_* flags of the current branch, y = arc list (see arcs_of), z = prev, w = next (slots / NONE / TERM)
//   F  x = parent of the pre node (slot / PAR_*), y = branch of pre, z = branch of post, w = first arc while staging,
//      then the stamp of the batch of paths that last went through the branch rooted here
//   E  exact distances from S of pre and post (the invariant the incremental update relies on)
//   W  costs: S -> pre, pre -> post, post -> T, matched arc prev -> this
struct __align__(16) GRec {
    int4 A, F;
    longlong2 E;
    Cost4 W;
};
static_assert(sizeof(GRec) == 80, "record layout");

struct GCounters {
    int64_t n_arc;            // arcs read while staging
    int64_t n_branch, n_rec;  // branches re-labelled, records listed for them
    int64_t n_cert;           // components finished by the optimality certificate instead of a last re-labelling
    long long t_phase[7];  // cycles (thread 0): arg-min, branch list, staging + initial labels, walk, fix-point, -,
                           // optimality certificate
    long long t_mark, t_start;
    __device__ __forceinline__ void tick(int k) {
        const long long now = clock64();
        t_phase[k] += now - t_mark;
        t_mark = now;
    }
};

// Solve one component: positions [c0, c0 + n) of graph g, staged in the global workspace.
__device__ __noinline__ void solve_component_global(const SolveParams &p, const int cg, const int c0, const int n, const int g,
                                             int *s_scan, KeyIdx *s_red, int *s_ok, long long *s_work,
                                             const int sweep_warp) {
    typedef GStaging S;
    typedef GRec Rec;
    typedef typename S::arc_t Arc;
    typedef typename S::costs_t Costs;
    typedef typename S::cost_t Cost;
    typedef typename S::slot_t Slot;
    const int T = blockDim.x, tid = threadIdx.x;
    const int sub = tid % LPN, grp = tid / LPN, ngrp = T / LPN;
    const unsigned gmask = ((1u << LPN) - 1u) << ((tid & 31) / LPN * LPN);
    // slot lists: bl = the branch in layer order, hop = detections whose matched in-arc changed, pl = the branch's
    // unused records with a post node to re-label ("producers" of the ascending sweep)
    Rec *const rec = (Rec *)(p.ws_rec + c0);
    int32_t *const tk = p.ws_tk + c0;
    Slot *const bl = (Slot *)(p.ws_bl + c0), *const hop = (Slot *)(p.ws_hop + c0), *const pl = (Slot *)(p.ws_pl + c0);
    // arc list of a record: A.y = ~offset in the global arc area (the list of detection d starts at in_ptr[d] + d, so
    // that lists of different detections never overlap); its first arc also carries the length of the list
    auto arcs_of = [&](int ay) -> Arc * { return (Arc *)(p.ws_arc + ~ay); };
    const int32_t *perm = p.comp_perm + c0;

    GCounters cnt;
    cnt.n_arc = cnt.n_branch = cnt.n_rec = cnt.n_cert = 0;
    for (int k = 0; k < 7; k++) cnt.t_phase[k] = 0;
    cnt.t_mark = cnt.t_start = clock64();

    // ---- stage the component. Records first: every load of a record is issued before any is used (one memory level
    //      after the permutation); its in-degree and first arc wait in A.x / F.w for the arc pass. Arc-list offsets by
    //      a block scan over contiguous chunks.
    const int chunk = (n + T - 1) / T;
    const int i0 = min(tid * chunk, n), i1 = min(i0 + chunk, n);
    {
        int my_arcs = 0;
        for (int i = i0; i < i1; i++) {
            const int q = c0 + i, d = perm[i];
            const int e0 = p.in_ptr[d], e1 = p.in_ptr[d + 1];
            const FtTmp ft = p.ft[q];
        const int par = ft.par, ap_ = ft.anc, ao_ = ft.anc;
            const int64_t dpre = ft.dpre, dpost = ft.dpost;
            const int64_t en = p.en[d], cn = p.cn[d], ex = p.ex[d];
            const int key = p.tkey[d];
            Rec r;
            r.A = make_int4(e1 - e0, 0, NONE, NONE);
            r.F = make_int4(par >= 0 ? par - c0 : par, ap_ >= 0 ? ap_ - c0 : -1, ao_ >= 0 ? ao_ - c0 : -1, e0);
            r.E = make_longlong2(dpre, dpost);
            r.W.x = S::conv(en);
            r.W.y = S::conv(cn);
            r.W.z = S::conv(ex);
            r.W.w = 0;
            rec[i] = r;
            tk[i] = key;
            my_arcs += e1 - e0 + 1;  // + 1: a list without arcs still carries its length
        }
        int tot;
        int a0 = block_excl_scan1(my_arcs, s_scan, 0, &tot);
        for (int i = i0; i < i1; i++) {
            const int deg = rec[i].A.x;
            int off = a0;
            a0 += deg + 1;
            off = ~(rec[i].F.w + perm[i]);  // the list lives at the detection's own place of the global arc area
            rec[i].A.y = off;
        }
        if (tid == 0) cnt.n_arc += tot - n;
    }
    __syncthreads();
    for (int ib = 0; ib < n; ib += ngrp) {  // arc lists: one lane group per record, two independent arcs per lane
        const int i = ib + grp;
        if (i < n) {
            const int e0 = rec[i].F.w, deg = rec[i].A.x;
            Arc *const ap = arcs_of(rec[i].A.y);
            for (int kb = 0; kb < deg; kb += 2 * LPN) {
                const int k0 = kb + sub, k1 = kb + LPN + sub;
                const int src0 = k0 < deg ? p.in_src[e0 + k0] : 0, src1 = k1 < deg ? p.in_src[e0 + k1] : 0;
                const int64_t w0 = k0 < deg ? p.in_cost[e0 + k0] : 0, w1 = k1 < deg ? p.in_cost[e0 + k1] : 0;
                const int q0 = p.comp_inv[src0], q1 = p.comp_inv[src1];
                if (k0 < deg) {
                    Arc x;
                    x.cost = (decltype(x.cost))w0;
                    x.src = (decltype(x.src))(q0 - c0);
                    x.pad = (decltype(x.pad))deg;  // only read from the first arc of the list
                    ap[k0] = x;
                }
                if (k1 < deg) {
                    Arc x;
                    x.cost = (decltype(x.cost))w1;
                    x.src = (decltype(x.src))(q1 - c0);
                    x.pad = (decltype(x.pad))deg;
                    ap[k1] = x;
                }
            }
            if (deg == 0 && sub == 0) {
                Arc x;
                x.cost = 0;
                x.src = 0;
                x.pad = 0;
                ap[0] = x;
            }
        }
    }
    __syncthreads();
    cnt.tick(2);

    int K = 0, n_used = 0;
    int64_t total = 0;
    for (int iter = 0; iter <= n; iter++) {  // at most one path per detection
        // ---- a batch of paths from ONE set of labels (solve_sm.cuh, "next path"): sinks in ascending order, each in a
        //      branch no earlier path of the batch went through. The root of a used branch carries the batch's stamp
        //      in F.w (free after staging; a first-arc index there is never negative).
        const int stamp = -2 - iter;
        bool finished = false;
#pragma unroll 1
        for (;;) {
        // ---- cheapest way into T
        KeyIdx best;
        best.key = INF;
        best.idx = -1;
        for (int i = tid; i < n; i += T) {
            const int64_t d = rec[i].E.y;
            const Cost x = rec[i].W.z;
            if (rec[i].A.w != TERM && d < INF_CUT && S::has(x)) {
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
        // (read before the barrier: thread 0 stamps the root right after it)
        const int root = (sink.idx >= 0 && sink.key < 0) ? rec[sink.idx].F.z : -1;
        const bool root_used = root >= 0 && rec[root].F.w == stamp;
        __syncthreads();  // s_red is rewritten next iteration
        cnt.tick(0);
        finished = true;
        if (sink.idx < 0) break;  // T unreachable
        if (sink.key >= 0) {      // see k_ssp_solve: the first-path rule belongs to the whole graph
            if (K == 0 && tid == 0) {
                p.first_cost[cg] = sink.key;
                p.first_sink[cg] = c0 + sink.idx;
            }
            break;
        }
        if (root < 0) break;  // cannot happen: a labelled node always has a branch
        finished = false;
        if (root_used) break;  // its branch already carries a path of this batch: stale labels, re-label first
        // ---- augment: walk the path back from T and rewrite next/prev (Graph::flip_path, Graph.cpp:266-335)
        if (tid == 0) {
            rec[root].F.w = stamp;
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
                if (q < 0) {  // PAR_S: the path starts here
                    rec[m].A.z = TERM;
                    ok = 1;
                    break;
                }
                rec[m].A.z = q;
                hop[nhop++] = (Slot)m;  // the cost of the newly matched arc q -> m is looked up below
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
            const Arc *const ap = arcs_of(rec[m].A.y);
            const int deg = ap[0].pad;
            for (int k = 0; k < deg; k++) {
                const Arc x = ap[k];
                if ((int)x.src == q) {
                    rec[m].W.w = x.cost;
                    break;
                }
            }
        }
        if (tid == 0 && p.path_cost) p.path_cost[c0 + K] = sink.key;
        K++;
        total += sink.key;
        __syncthreads();
        }  // (batch)
        if (finished) break;
        // ---- Is this component finished? Once every detection is on a track, a further S -> T path in the residual
        //      network has the shape  S -> pre_i, (pre -> post of the predecessor: cancels a matched arc, -mc), then
        //      any number of hops post_a -> pre_b -> post_prev(b) (an unused arc and the cancellation of b's matched
        //      arc: w_ab - mc_b) or post_a -> pre_a -> post_prev(a) (-cn_a - mc_a), then post -> T. If every matched arc
        //      is the cheapest in-arc of its head (w_ab >= mc_b), mc <= 0, cn + mc <= 0 and no entry / exit cost is
        //      negative, every piece is >= 0: no path can pay, exactly what the re-labelling + arg-min below would
        //      find out. One pass over the arcs instead.
        if (n_used == n) {
            bool good = true;
            for (int ib = 0; ib < n; ib += ngrp) {
                const int i = ib + grp;
                if (i < n) {
                    const int4 A = rec[i].A;
                    const Costs W = rec[i].W;
                    if ((S::has(W.x) && W.x < 0) || (S::has(W.z) && W.z < 0)) good = false;
                    if (A.z >= 0) {
                        const int64_t mc = W.w;
                        if (mc > 0 || (int64_t)W.y + mc > 0) good = false;
                        const Arc *const ap = arcs_of(A.y);
                        const int deg = ap[0].pad;
                        for (int k = sub; k < deg; k += LPN) {
                            const Arc x = ap[k];
                            if ((int)x.src != A.z && (int64_t)x.cost < mc) good = false;
                        }
                    }
                }
            }
            if (__syncthreads_and(good)) {
                cnt.n_cert++;
                cnt.tick(6);
                break;
            }
            cnt.tick(6);
        }
        // ---- the branch hanging from S -> pre_root, as a list of slots in layer order
        int nd;
        {
            int my_nd = 0;
            for (int i = i0; i < i1; i++) {
                const int4 F = rec[i].F;
                const int f = ((F.y >= 0 && rec[F.y].F.w == stamp) ? DL_PRE : 0) |
                              ((F.z >= 0 && rec[F.z].F.w == stamp) ? DL_POST : 0);
                rec[i].A.x = f;
                my_nd += f ? 1 : 0;
            }
            int lp = block_excl_scan1(my_nd, s_scan, 1, &nd);
            if (my_nd)
                for (int i = i0; i < i1; i++)
                    if (rec[i].A.x) bl[lp++] = (Slot)i;
        }
        __syncthreads();
        for (int lp = tid; lp < nd; lp += T)
            if (lp == 0 || tk[bl[lp]] != tk[bl[lp - 1]]) rec[bl[lp]].A.x |= DL_FIRST;
        cnt.n_branch++;
        cnt.n_rec += nd;
        cnt.tick(1);
        {  // the "producers" of the ascending sweeps (see below): every thread looks at a chunk of the list
            const int lchunk = (nd + T - 1) / T;
            const int lp0 = min(tid * lchunk, nd), lp1 = min(lp0 + lchunk, nd);
            int mine = 0;
            for (int lp = lp0; lp < lp1; lp++) {
                const int4 A = rec[bl[lp]].A;
                mine += ((A.x & DL_POST) && A.z == NONE) ? 1 : 0;
            }
            int np;
            int k = block_excl_scan1(mine, s_scan, 0, &np);
            if (mine)
                for (int lp = lp0; lp < lp1; lp++) {
                    const int4 A = rec[bl[lp]].A;
                    if ((A.x & DL_POST) && A.z == NONE) pl[k++] = bl[lp];
                }
            if (tid == 0) {
                s_ok[6] = np;
                s_ok[2] = s_ok[3] = s_ok[4] = s_ok[5] = 0;  // flags of the fix-point rounds
            }
        }
        __syncthreads();
        // ---- new labels of the branch start from what cannot change during the fix-point: in-arcs from outside the
        //      branch, the S arc, the own post node when it is outside. One lane group per record.
        for (int sb = 0; sb < nd; sb += ngrp) {
            const int lp = sb + grp;
            const bool valid = lp < nd;
            const int s = valid ? bl[lp] : 0;
            const int4 A = rec[s].A;
            KeyIdx cb;
            cb.key = INF;
            cb.idx = -1;
            if (valid && (A.x & DL_PRE)) {
                const Arc *const ap = arcs_of(A.y);
                const int deg = ap[0].pad;
                for (int k = sub; k < deg; k += LPN) {
                    const Arc x = ap[k];
                    const int64_t d = rec[x.src].E.y;
                    if (!(rec[x.src].A.x & DL_POST) && (int)x.src != A.z && d < INF_CUT) {
                        const int64_t cand = d + x.cost;
                        if (cand < cb.key || (cand == cb.key && (unsigned)x.src < (unsigned)cb.idx)) {
                            cb.key = cand;
                            cb.idx = x.src;
                        }
                    }
                }
            }
#pragma unroll
            for (int d = 1; d < LPN; d <<= 1) {
                const KeyIdx o = shfl_xor_ki(cb, d, gmask);
                const bool take = o.key < cb.key || (o.key == cb.key && (unsigned)o.idx < (unsigned)cb.idx);
                cb.key = take ? o.key : cb.key;
                cb.idx = take ? o.idx : cb.idx;
            }
            // every lane of every group has finished reading its sources' post labels before any of them is reset
            // (a record with DL_POST is never a source here, but its E/F words are rewritten as a whole)
            __syncwarp();
            if (valid && sub == 0) {
                int4 F = rec[s].F;
                const longlong2 E = rec[s].E;
                const Costs W = rec[s].W;
                int64_t key = INF, dpost_new = E.y;
                int par = PAR_NONE, anc = -1;
                if (A.x & DL_PRE) {
                    if (cb.idx >= 0) {
                        key = cb.key;
                        par = cb.idx;
                        anc = rec[cb.idx].F.z;
                    }
                    if (!(A.x & DL_POST) && A.z != NONE && E.y < INF_CUT && E.y - W.y < key) {
                        key = E.y - W.y;  // own post node is outside the branch: a constant candidate
                        par = PAR_POST;
                        anc = F.z;
                    }
                    if (A.z != TERM && S::has(W.x) && W.x < key) {
                        key = W.x;
                        par = PAR_S;
                        anc = s;
                    }
                    if (key >= INF_CUT) {
                        key = INF;
                        par = PAR_NONE;
                        anc = -1;
                    }
                    F.x = par;
                    F.y = anc;
                }
                if (A.x & DL_POST) {
                    const bool unused_labelled = (A.x & DL_PRE) && A.z == NONE && key < INF_CUT;
                    dpost_new = unused_labelled ? key + W.y : INF;
                    F.z = unused_labelled ? anc : -1;
                }
                rec[s].E = make_longlong2((A.x & DL_PRE) ? key : E.x, dpost_new);
                rec[s].F = F;
            }
        }
        __syncthreads();
        cnt.tick(2);
        // ---- fix-point (see relabel_branch). Ascending sweep in two phases: (A) the unused records whose post label
        //      is recomputed from their pre label are the only ones that pass a new label on to higher layers, they
        //      go first, layer by layer, by one warp; (B) every other record only reads post labels that are final
        //      for this sweep, so all of them are relaxed at once by the whole block, 8 lanes per record.
        //      Descending pass: independent chain walks, one thread per track segment.
        {
            const int lane = lane_id(), wid = warp_id(), nw = (T + 31) >> 5;
            const int r8 = lane >> 3, a8 = lane & 7;
            int n_node32 = 0, n_arc32 = 0, n_astep = 0, n_dstep = 0;
            long long t_asc = 0, t_desc = 0, t0 = clock64();
            const int np = s_ok[6];  // producers (list pl)
            // relax the pre node of record s over its in-arcs, S and its own post node, then its post node; 8 lanes.
            // Returns true if a change can travel along a reversed arc (to lower layers).
            auto relax = [&](const int s, const bool valid) -> bool {
                const Rec *rp = rec + s;
                const int4 A = rp->A;
                const Costs W = rp->W;
                int4 F = rp->F;
                const longlong2 E = rp->E;
                const int ent = valid ? A.x : 0;
                KeyIdx best;
                best.key = INF;
                best.idx = -1;
                int deg = 0;
                if (ent & DL_PRE) {
                    const Arc *const ap = arcs_of(A.y);
                    const Arc x0 = ap[a8];  // speculative: the first arc also carries the list length
                    deg = ap[0].pad;
                    if (a8 < deg) {
                        const int64_t d = rec[x0.src].E.y;
                        if ((int)x0.src != A.z && d < INF_CUT) {
                            best.key = d + x0.cost;
                            best.idx = x0.src;
                        }
                    }
                    for (int k = a8 + 8; k < deg; k += 8) {
                        const Arc x = ap[k];
                        const int64_t d = rec[x.src].E.y;
                        if ((int)x.src != A.z && d < INF_CUT &&
                            (d + x.cost < best.key || (d + x.cost == best.key && (unsigned)x.src < (unsigned)best.idx))) {
                            best.key = d + x.cost;
                            best.idx = x.src;
                        }
                    }
                }
#pragma unroll
                for (int d = 1; d < 8; d <<= 1) {  // lowest slot wins ties
                    const KeyIdx o = shfl_xor_ki(best, d);
                    const bool take = o.key < best.key || (o.key == best.key && (unsigned)o.idx < (unsigned)best.idx);
                    best.key = take ? o.key : best.key;
                    best.idx = take ? o.idx : best.idx;
                }
                bool down = false;
                if (ent != 0 && a8 == 0) {
                    int64_t dpre_s = E.x, dpost_s = E.y;
                    bool upd = false;
                    if (ent & DL_PRE) {
                        int64_t key = INF;
                        int par = PAR_NONE, anc = -1;
                        if (best.idx >= 0) {
                            key = best.key;
                            par = best.idx;
                            anc = rec[best.idx].F.z;
                        }
                        if (A.z != TERM && S::has(W.x) && W.x < key) {
                            key = W.x;
                            par = PAR_S;
                            anc = s;
                        }
                        if (A.z != NONE && dpost_s < INF_CUT && dpost_s - W.y < key) {
                            key = dpost_s - W.y;
                            par = PAR_POST;
                            anc = F.z;
                        }
                        if (key < dpre_s) {
                            dpre_s = key;
                            F.x = par;
                            F.y = anc;
                            upd = true;
                            down = A.z >= 0;  // pre_s feeds the post node of its predecessor on the track
                        }
                        n_arc32 += deg;
                    }
                    // The post node of a record on a track hangs from its successor's pre node: that is the descending
                    // pass's business, and it keeps every post label read during phase (B) stable.
                    if ((ent & DL_POST) && A.z == NONE && dpre_s < INF_CUT && dpre_s + W.y < dpost_s) {
                        dpost_s = dpre_s + W.y;
                        F.z = F.y;
                        upd = true;
                    }
                    if (upd) {
                        rec[s].E = make_longlong2(dpre_s, dpost_s);
                        rec[s].F = F;
                    }
                }
                return down;
            };
            for (int round = 0; round < 2 * nd + 4; round++) {
                const int pb = round & 1;
                if (round > 0) {  // round 0: the initial labels stand in for the ascending sweep
                    bool down = false;
                    if (wid == sweep_warp) {  // (A) producers, one layer at a time
                        for (int q = 0; q < np;) {
                            const int k = q + r8;
                            const bool in_range = k < np;
                            const int s = pl[in_range ? k : np - 1];
                            const bool same = in_range && tk[s] == tk[pl[q]];
                            const unsigned m = __ballot_sync(0xffffffffu, a8 == 0 && !same) & 0x01010100u;
                            const int cn_t = m ? (__ffs(m) - 1) >> 3 : 4;
                            down |= relax(s, r8 < cn_t);
                            n_node32 += cn_t;
                            n_astep++;
                            q += cn_t;
                            __syncwarp();
                        }
                    }
                    __syncthreads();
                    for (int base = 0; base < nd; base += 4 * nw) {  // (B) everything else, all at once
                        const int lp = base + 4 * wid + r8;
                        const bool in_range = lp < nd;
                        const int s = bl[in_range ? lp : nd - 1];
                        const int f = rec[s].A.x;
                        const bool producer = (f & DL_POST) && rec[s].A.z == NONE;
                        down |= relax(s, in_range && !producer);
                        n_astep++;
                    }
                    if (wid == 0) n_node32 += nd - np;
                    if (__any_sync(0xffffffffu, down) && lane == 0) s_ok[2 + pb] = 1;
                    if (wid == sweep_warp && lane == 0) s_work[0] += 1;
                    __syncthreads();
                    {
                        const long long t1 = clock64();
                        t_asc += t1 - t0;
                        t0 = t1;
                    }
                    if (!s_ok[2 + pb]) break;  // nothing can travel down: fix-point
                }
                // descending pass: every thread looks at its share of the list
                bool up = false;
                for (int lp = tid; lp < nd; lp += T) {
                    const int s0 = bl[lp];
                    int4 A = rec[s0].A;
                    if (A.z == NONE) {
                        if (A.x & DL_POST) {
                            const longlong2 E = rec[s0].E;
                            const Cost cn = rec[s0].W.y;
                            if (E.x < INF_CUT && E.x + cn < E.y) {
                                rec[s0].E.y = E.x + cn;
                                rec[s0].F.z = rec[s0].F.y;
                                up = true;
                            }
                        }
                        continue;
                    }
                    if ((A.x & DL_POST) && A.w >= 0) continue;  // reached from its successor's walk
                    // on a track, successor not in the branch: a segment starts here; walk it down
                    int cur = s0;
                    Costs W = rec[cur].W;
                    int4 F = rec[cur].F;
                    longlong2 E = rec[cur].E;
                    int64_t nx_dpre = INF, nx_mc = 0;
                    int nx_apre = -1;
                    for (;;) {
                        // the predecessor's post node hangs from pre_cur: it is in the branch when pre_cur is
                        const int prv = ((A.x & DL_PRE) && A.z >= 0) ? A.z : -1;
                        int4 An = A, Fn = F;
                        Costs Wn = W;
                        longlong2 En = E;
                        if (prv >= 0) {  // fetch the next record of the walk before working on this one
                            const Rec *rn = rec + prv;
                            An = rn->A;
                            Wn = rn->W;
                            Fn = rn->F;
                            En = rn->E;
                        }
                        const int ent = A.x;
                        int64_t dpre_s = E.x, dpost_s = E.y;
                        bool upd = false;
                        if ((ent & DL_POST) && nx_dpre < INF_CUT && nx_dpre - nx_mc < dpost_s) {
                            dpost_s = nx_dpre - nx_mc;
                            F.z = nx_apre;
                            upd = true;
                            up = true;  // forward arcs leave post_s: the ascending sweep has to come back
                        }
                        if ((ent & DL_PRE) && (ent & DL_POST) && dpost_s < INF_CUT && dpost_s - W.y < dpre_s) {
                            dpre_s = dpost_s - W.y;
                            F.x = PAR_POST;
                            F.y = F.z;
                            upd = true;
                        }
                        if (upd) {
                            rec[cur].E = make_longlong2(dpre_s, dpost_s);
                            rec[cur].F = F;
                        }
                        n_dstep++;
                        if (prv < 0) break;
                        nx_dpre = dpre_s;
                        nx_mc = W.w;
                        nx_apre = F.y;
                        cur = prv;
                        A = An;
                        W = Wn;
                        F = Fn;
                        E = En;
                    }
                }
                if (__any_sync(0xffffffffu, up) && lane == 0) s_ok[4 + pb] = 1;
                if (tid == 0) {
                    s_ok[2 + (pb ^ 1)] = 0;  // the next round's flags
                    s_ok[4 + (pb ^ 1)] = 0;
                    s_work[1] += 1;
                }
                __syncthreads();
                {
                    const long long t1 = clock64();
                    t_desc += t1 - t0;
                    t0 = t1;
                }
                if (!s_ok[4 + pb] && round > 0) break;  // nothing can travel up: fix-point
            }
            // work counters of this branch
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                n_arc32 += __shfl_xor_sync(0xffffffffu, n_arc32, d);
                n_dstep += __shfl_xor_sync(0xffffffffu, n_dstep, d);
            }
            if (lane == 0) {
                atomicAdd((unsigned long long *)&s_work[2], (unsigned long long)(n_node32 + n_dstep));
                atomicAdd((unsigned long long *)&s_work[3], (unsigned long long)n_arc32);
                atomicAdd((unsigned long long *)&s_work[7], (unsigned long long)n_dstep);
                if (wid == sweep_warp) {
                    s_work[4] += t_asc;
                    s_work[5] += t_desc;
                    atomicAdd((unsigned long long *)&s_work[6], (unsigned long long)n_astep);
                }
            }
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
            atomicAdd(&o[1], (unsigned long long)s_work[1]);
            atomicAdd(&o[2], (unsigned long long)s_work[2]);
            atomicAdd(&o[3], (unsigned long long)(cnt.n_arc + s_work[3]));
            for (int k = 0; k < 6; k++) atomicAdd(&o[4 + k], (unsigned long long)cnt.t_phase[k]);
            atomicAdd(&o[9], (unsigned long long)cnt.t_phase[6]);
            atomicAdd(&o[10], (unsigned long long)cnt.n_cert);
            atomicAdd(&o[11], (unsigned long long)cnt.n_branch);
            atomicAdd(&o[12], (unsigned long long)cnt.n_rec);
            for (int k = 0; k < 4; k++) atomicAdd(&o[13 + k], (unsigned long long)s_work[4 + k]);
            {  // [18..23]: the same four once more, the number of components solved here and their cycles
                for (int k = 0; k < 4; k++) atomicAdd(&o[18 + k], (unsigned long long)s_work[4 + k]);
                atomicAdd(&o[22], 1ull);
                atomicAdd(&o[23], (unsigned long long)(clock64() - cnt.t_start));
            }
        }
    }
    for (int i = tid; i < n; i += T) {
        const int a = rec[i].A.w, b = rec[i].A.z;
        const int d = perm[i];
        p.next_out[d] = a >= 0 ? perm[a] : a;
        p.prev_out[d] = b >= 0 ? perm[b] : b;
    }
    __syncthreads();
}

