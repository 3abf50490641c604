// Block-per-component solver for components that fit no shared-memory band (solve_sm.cuh) but whose costs fit 32
// bits: the staging area lives in GLOBAL memory (it stays in L2: 150 bytes per record, 12 per arc), slots are 32 bits,
// any size. Included by solve.cu. Costs beyond 32 bits, and thin components, go to solve_global.cuh (k_ssp_solve picks).
//
// Same successive shortest paths as solve_sm.cuh (arg-min over the sinks, several paths per re-labelling as long as
// their branches differ, path walk, certificate, exact re-labelling of the branches that hung from the paths' first
// arcs by synchronous pull rounds), with two differences that matter at this size:
//   pushed work lists   a round only touches the records whose inputs moved in the round before. Publishing a post label
//                       puts the heads of the node's out-arcs on the next round's work list (the out-arc lists are
//                       staged next to the in-arc lists); publishing a pre label on a track lists the track's previous
//                       record (reversed matched arc); bits per record keep a record from being listed twice. A round
//                       costs what its frontier costs, not what the component costs.
//   Dijkstra's order    the labels before the flip are potentials (every residual arc has a non-negative reduced cost,
//                       Johnson; what muSSP's heap works on, Graph.cpp:394-503), so a label that rises by less is
//                       final earlier. Records are listed with the smallest increase among the inputs they wait for;
//                       increases beyond the current threshold wait on a second list that is split again when the
//                       rounds of the bucket have reached their fix-point (threshold = smallest waiting increase +
//                       BIG_DELTA). Tentative labels that a better one overwrites anyway are never passed on: 10x fewer
//                       relaxations on a 98 000-detection component; any order reaches the same exact fix-point.
#pragma once

constexpr long long BIG_DELTA = 1ll << 26;  // width of a bucket of label increases, in cost units (1e-7): measured
                                            // on C5-whole, 2^18 4.3 s, 2^22 2.3 s, 2^24 1.9 s, 2^26 1.27 s, 1e9 1.40 s

__device__ __noinline__ void solve_component_big(const SolveParams &p, const int cg, const int c0, const int n, const int g,
                                                 int *s_scan, KeyIdx *s_red, int *s_ok, long long *s_work) {
    const int T = blockDim.x, tid = threadIdx.x;
    const int sub = tid % LPN, grp = tid / LPN, ngrp = T / LPN;
    const unsigned gmask = ((1u << LPN) - 1u) << ((tid & 31) / LPN * LPN);
    SRec *const rec = (SRec *)(p.ws_rec + c0);  // (an 80-byte slot per record of the area solve_global.cuh uses)
    longlong2 *const res_e = p.big_res_e + c0;  // staged results of a round: new labels ...
    int4 *const res_f = p.big_res_f + c0;       // ... parent, branches, what moved (1 = pre, 2 = post)
    int *const hop = p.ws_hop + c0;             // detections whose matched in-arc changed (path walk); second work list
    int *const pl = p.ws_pl + c0;               // work list of a relabelling round
    unsigned char *const flg = (unsigned char *)(p.ws_tk + c0);  // DL_* of the branch being re-labelled
    int *const out_off = p.ws_bl + c0;          // out-arcs of record i: out_dst[out_off[i] .. out_off[i + 1])
    unsigned *const listed = p.big_listed + (c0 >> 4) + cg;  // two bits per record: on the next round's work list /
                                                             // on the waiting list (every component has its own
                                                             // words: + component index)
    longlong2 *const old_e = p.big_old_e + c0;  // labels of the branch before the path was flipped: the potentials
    long long *const prio = p.big_prio + c0;    // smallest label increase among the inputs a listed record waits for
    int *const far0 = p.big_far + 2 * (size_t)c0, *const far1 = far0 + n;  // the waiting list, two buffers
    const int32_t *const perm = p.comp_perm + c0;

    // ---- stage the component. Arc offsets by a block scan over contiguous chunks of records.
    const int chunk = (n + T - 1) / T;
    const int i0 = min(tid * chunk, n), i1 = min(i0 + chunk, n);
    int my_arcs = 0;
    for (int i = i0; i < i1; i++) {
        const int d = perm[i];
        my_arcs += p.in_ptr[d + 1] - p.in_ptr[d];
    }
    int tot;
    int a0 = block_excl_scan1(my_arcs, s_scan, 0, &tot);
    // the component's arcs: three arrays (cost, source slot of every in-arc; head slot of every out-arc) in a block of
    // the arc area taken from a bump allocator (every component asks for exactly its number of in-arcs, once: the area
    // of arc_capacity entries per array cannot run out)
    if (tid == 0) s_ok[6] = (int)atomicAdd(p.big_arc_next, (unsigned long long)tot);
    for (int w = tid; w < (n >> 4) + 1; w += T) listed[w] = 0;
    for (int i = tid; i < n; i += T) prio[i] = INF;
    __syncthreads();
    int32_t *const arc_cost = p.big_arc + s_ok[6];
    int *const arc_src = arc_cost + p.big_arc_stride;
    int *const out_dst = arc_src + p.big_arc_stride;
    int *const out_cnt = (int *)res_e;  // while staging: out-degree, then fill cursor of every record
    PhaseClock cnt;
    cnt.w = s_work;
    cnt.t_mark = clock64();
    // records: every load of a record is issued before any is used (one memory level after the permutation)
    for (int i = i0; i < i1; i++) {
        const int q = c0 + i, d = perm[i];
        const int e0 = p.in_ptr[d], e1 = p.in_ptr[d + 1];
        const FtTmp ft = p.ft[q];
        const int par = ft.par, ap_ = ft.anc, ao_ = ft.anc;
        const int64_t dpre = ft.dpre, dpost = ft.dpost;
        const int64_t en = p.en[d], cn = p.cn[d], ex = p.ex[d];
        SRec r;
        r.A = make_int4(e1 - e0, a0, NONE, NONE);
        r.F = make_int4(par >= 0 ? par - c0 : par, ap_ >= 0 ? ap_ - c0 : -1, ao_ >= 0 ? ao_ - c0 : -1, e0);
        r.E = make_longlong2(dpre, dpost);
        r.W = make_int4(en >= INF_CUT ? NO_ARC32 : (int)en, (int)cn, ex >= INF_CUT ? NO_ARC32 : (int)ex, 0);
        rec[i] = r;
        a0 += e1 - e0;
        out_cnt[4 * i] = 0;  // (res_e is 16 bytes per record: one counter in each)
    }
    cnt.add(0, tot);
    __syncthreads();
    for (int ib = 0; ib < n; ib += ngrp) {  // arc lists: one lane group per record, two independent arcs per lane
        const int i = ib + grp;
        if (i < n) {
            const int e0 = rec[i].F.w, deg = rec[i].A.x, off = rec[i].A.y;
            for (int kb = sub; kb < deg; kb += 2 * LPN) {
                const int k1 = kb + LPN;
                const bool two = k1 < deg;
                const int src0 = p.in_src[e0 + kb], src1 = p.in_src[e0 + (two ? k1 : kb)];
                const int64_t w0 = p.in_cost[e0 + kb], w1 = p.in_cost[e0 + (two ? k1 : kb)];
                const int q0 = p.comp_inv[src0] - c0, q1 = p.comp_inv[src1] - c0;
                arc_cost[off + kb] = (int)w0;
                arc_src[off + kb] = q0;
                atomicAdd(&out_cnt[4 * q0], 1);
                if (two) {
                    arc_cost[off + k1] = (int)w1;
                    arc_src[off + k1] = q1;
                    atomicAdd(&out_cnt[4 * q1], 1);
                }
            }
        }
    }
    __syncthreads();
    {   // out-arc lists: offsets by a block scan of the out-degrees, then every in-arc is entered at its source
        int my_out = 0;
        for (int i = i0; i < i1; i++) my_out += out_cnt[4 * i];
        int tot_out;
        int o0 = block_excl_scan1(my_out, s_scan, 1, &tot_out);
        for (int i = i0; i < i1; i++) {
            const int c = out_cnt[4 * i];
            out_off[i] = o0;
            out_cnt[4 * i] = o0;  // fill cursor
            o0 += c;
        }
        __syncthreads();
        for (int ib = 0; ib < n; ib += ngrp) {
            const int i = ib + grp;
            if (i < n) {
                const int deg = rec[i].A.x, off = rec[i].A.y;
                for (int k = sub; k < deg; k += LPN) out_dst[atomicAdd(&out_cnt[4 * arc_src[off + k]], 1)] = i;
            }
        }
    }
    __syncthreads();
    cnt.tick(2);

    const unsigned lt = (1u << lane_id()) - 1u;
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
        // (read before the barrier: thread 0 stamps the root right after it)
        const int root = (sink.idx >= 0 && sink.key < 0) ? rec[sink.idx].F.z : -1;
        const bool root_used = root >= 0 && rec[root].F.w == stamp;
        __syncthreads();  // s_red is rewritten next iteration
        cnt.tick(0);
        finished = true;
        if (sink.idx < 0) break;  // T unreachable
        if (sink.key >= 0) {      // the first-path rule belongs to the whole graph (k_first_path_rule)
            if (K == 0 && tid == 0) {
                p.first_cost[cg] = sink.key;
                p.first_sink[cg] = c0 + sink.idx;
            }
            break;
        }
        if (root < 0) break;  // cannot happen: a labelled node always has a branch
        finished = false;
        if (root_used) break;  // its branch already carries a path of this batch: stale labels, re-label first
        // ---- augment: walk the path back from T and rewrite next / prev
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
                if (q < 0) {       // PAR_S: the path starts here
                    rec[m].A.z = TERM;
                    ok = 1;
                    break;
                }
                rec[m].A.z = q;
                hop[nhop++] = m;  // the cost of the newly matched arc q -> m is looked up below
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
        // ---- the branch hanging from S -> pre_root loses its labels; its records are the first round's work list
        int n_mine = 0;
        if (tid == 0) s_ok[3] = 0;
        __syncthreads();
        for (int base = 0; base < n; base += T) {
            const int s = base + tid;
            int f = 0;
            if (s < n) {
                int4 F = rec[s].F;
                // 1 = pre node, 2 = post node in one of the batch's branches
                f = ((F.y >= 0 && rec[F.y].F.w == stamp) ? 1 : 0) | ((F.z >= 0 && rec[F.z].F.w == stamp) ? 2 : 0);
                flg[s] = (unsigned char)f;
                if (f) {
                    longlong2 E = rec[s].E;
                    old_e[s] = E;
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
            const unsigned m = __ballot_sync(0xffffffffu, f != 0);
            if (m) {
                int at = 0;
                if (lane_id() == 0) at = atomicAdd(&s_ok[3], __popc(m));
                at = __shfl_sync(0xffffffffu, at, 0);
                if (f) pl[at + __popc(m & lt)] = s;
            }
        }
        cnt.add(1, 1);
        __syncthreads();
        cnt.tick(1);
        int n_arc_pull = 0, n_relax = 0, rounds = 0, n_far = 0, far_buf = 0;
        // Order of work. Labels only rise when a path is taken out (the old labels are potentials: every residual arc
        // has a non-negative reduced cost), and a label that has risen by less is final earlier - Dijkstra's order on
        // the increases. The rounds follow it bucket by bucket: a record is listed with the smallest increase among the
        // labels it waits for (prio); increases up to the threshold thr go on the next round's list, larger ones on a
        // waiting list. When the rounds of a bucket have reached their fix-point the threshold moves to the smallest
        // increase waiting (+ BIG_DELTA) and the waiting list is split again. Tentative labels that a better one would
        // overwrite anyway (a track started afresh at every node of the branch, detours over neighbouring tracks) are
        // then never passed on. Any order gives the same fix-point; this one keeps the rounds' frontiers thin.
        // two work lists (pl, hop) and their lengths (s_ok[3], s_ok[4]) alternate between the rounds; s_ok[5] = length
        // of the waiting list being written
        long long thr = -1;  // first round: everything the branch's records pass on waits
        if (tid == 0) s_ok[5] = 0;
#pragma unroll 1
        for (; rounds < 64 * n + 64; rounds++) {
            const int *const cur = (rounds & 1) ? hop : pl;
            int *const nxt = (rounds & 1) ? pl : hop;
            int *const n_next = &s_ok[3 + ((rounds + 1) & 1)];
            int n_work = s_ok[3 + (rounds & 1)];
            if (n_work == 0) {
                // ---- the bucket is done: next threshold, and the waiting records below it become the work list
                n_far = s_ok[5];
                if (n_far == 0) break;  // fix-point
                int *const fa = far_buf ? far1 : far0, *const fb = far_buf ? far0 : far1;
                long long lo = INF;
                for (int k = tid; k < n_far; k += T) lo = min(lo, prio[fa[k]]);
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
                __syncthreads();  // (everybody has read the list lengths)
                if (lane_id() == 0) s_red[warp_id()].key = lo;
                if (tid == 0) {
                    s_ok[5] = 0;
                    s_ok[3 + (rounds & 1)] = 0;
                }
                __syncthreads();
                lo = INF;
                for (int k = 0; k < (T + 31) >> 5; k++) lo = min(lo, (long long)s_red[k].key);
                thr = lo >= INF_CUT ? INF : lo + p.big_delta;
                int *const cur_w = (rounds & 1) ? hop : pl;
                for (int kb = 0; kb < n_far; kb += T) {
                    const int k = kb + tid;
                    int d = -1, where = 0;  // 1 = work list, 2 = keeps waiting, 0 = relaxed meanwhile: dropped
                    if (k < n_far) {
                        d = fa[k];
                        const long long pr = prio[d];
                        where = pr >= INF_CUT ? 0 : (pr <= thr ? 1 : 2);
                    }
                    if (where != 2 && d >= 0) {
                        const unsigned sh = (d & 15) * 2;
                        const unsigned old = atomicAnd(&listed[d >> 4], ~(2u << sh));  // off the waiting list
                        if (where == 1 && ((old | atomicOr(&listed[d >> 4], 1u << sh)) >> sh & 1u)) where = 0;  // listed already
                    }
                    const unsigned m1 = __ballot_sync(0xffffffffu, where == 1), m2 = __ballot_sync(0xffffffffu, where == 2);
                    int a1 = 0, a2 = 0;
                    if (lane_id() == 0) {
                        if (m1) a1 = atomicAdd(&s_ok[3 + (rounds & 1)], __popc(m1));
                        if (m2) a2 = atomicAdd(&s_ok[5], __popc(m2));
                    }
                    a1 = __shfl_sync(0xffffffffu, a1, 0);
                    a2 = __shfl_sync(0xffffffffu, a2, 0);
                    if (where == 1) cur_w[a1 + __popc(m1 & lt)] = d;
                    if (where == 2) fb[a2 + __popc(m2 & lt)] = d;
                }
                far_buf ^= 1;
                __syncthreads();
                n_work = s_ok[3 + (rounds & 1)];
                if (n_work == 0) continue;  // (everything waiting had been relaxed meanwhile: look again)
            }
            __syncthreads();  // (the list lengths have been read)
            if (tid == 0) *n_next = 0;
            // (A) relax the listed records from the labels of the round before, four lanes per record; the results are
            //     staged
#pragma unroll 1
            for (int wi = grp; wi < n_work; wi += ngrp) {
                const int s = cur[wi];
                // (its bit comes off the list here: set bits are exactly the records waiting on the next list, and the
                // words are all zero again when the rounds end)
                if (sub == 0) {
                    atomicAnd(&listed[s >> 4], ~(1u << ((s & 15) * 2)));
                    prio[s] = INF;
                }
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
                    const int *const as = arc_src + A.y;
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
                    res_f[wi] = make_int4(par, apre, apost, upd);
                    n_relax++;
                }
            }
            __syncthreads();  // every label of the round before has been read
#pragma unroll 1
            for (int wi = grp; wi < n_work; wi += ngrp) {  // (B) publish, and list what reads a label that moved: four
                const int4 rf = res_f[wi];                  //     lanes per record, one out-arc each
                if (rf.w) {
                    const int s = cur[wi];
                    // by how much the labels that moved have risen since the path was flipped
                    const longlong2 En = res_e[wi], Eo = old_e[s];
                    const int o0 = out_off[s], o1 = s + 1 < n ? (int)out_off[s + 1] : tot;
                    const int prv = rec[s].A.z;
                    if (sub == 0) {
                        rec[s].E = En;
                        int4 F = rec[s].F;
                        F.x = rf.x;
                        F.y = rf.y;
                        F.z = rf.z;
                        rec[s].F = F;
                    }
                    int *const far_w = far_buf ? far1 : far0;
                    auto list = [&](int d, long long inc) {
                        atomicMin(&prio[d], inc);
                        const unsigned sh = (d & 15) * 2;
                        if (inc <= thr) {
                            if (!(atomicOr(&listed[d >> 4], 1u << sh) >> sh & 1u)) nxt[atomicAdd(n_next, 1)] = d;
                        } else {
                            if (!(atomicOr(&listed[d >> 4], 2u << sh) >> sh & 2u)) far_w[atomicAdd(&s_ok[5], 1)] = d;
                        }
                    };
                    if (rf.w & 2) {  // post label: the heads of the out-arcs whose pre node is being re-labelled
                        const long long inc = En.y - Eo.y;
                        for (int k = o0 + sub; k < o1; k += LPN) {
                            const int d = out_dst[k];
                            if (flg[d] & 1) list(d, inc);
                        }
                    }
                    // pre label on a track: the post node of the record before it (reversed matched arc)
                    if ((rf.w & 1) && sub == LPN - 1 && prv >= 0 && (flg[prv] & 2)) list(prv, En.x - Eo.x);
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
}
