/* TEST INFRASTRUCTURE ONLY -- the product path never links, imports or calls this file.
 *
 * CPU restatement, in plain C on exact int64 costs, of the min-cost-flow solve the reference delegates to the
 * vendored muSSP solver (zip!/muSSP-master/muSSP, driven by mot3d/solvers/wrappers/muSSP/wrapper.cpp:171-317).
 *
 * What is restated (and where it lives in the reference):
 *   - the network: unit-capacity arcs, node 0 = source, node n-1 = sink (wrapper.cpp:57, Graph.cpp:13-19);
 *   - first shortest-path tree by one relaxation pass in topological order (Graph::shortest_path_dag,
 *     Graph.cpp:139-187), turned into node prices (Graph.cpp:183-186);
 *   - successive shortest paths on reduced costs c + price[u] - price[v] >= 0 with a Dijkstra search
 *     (Graph::update_shortest_path_tree_recursive, Graph.cpp:394-503 -- muSSP restricts the search to the branch
 *     of the tree the last path broke; that is a speed-up which does not change which path is found, so the
 *     restatement searches the whole residual graph);
 *   - residual update by reversing the arcs of the path (Graph::flip_path, Graph.cpp:266-335);
 *   - acceptance rule (wrapper.cpp:199-267): path 1 is always kept; path k>=2 is kept while its true cost
 *     cost_k = cost_{k-1} + dist(sink) is <= -1e-7, i.e. < 0 in integer units of 1e-7 (stop when cost_k >= 0);
 *   - flow read-out as one 0/1 flag per input arc (make_output2, wrapper.cpp:139-169).
 * Costs are the reference's own quantisation: graph_to_text prints weights with "{:0.7f}" (mot3d/mot3d.py:224),
 * so every cost muSSP sees is an integer multiple of 1e-7; the oracle works on those integers, which removes
 * muSSP's float tolerances (1e-7 / 1e-6, Graph.cpp:175,367,405) from the picture.
 *
 * Pinned against the real muSSP (oracle/_ref/libmussp_ref.so) by tests/test_oracle.py on every golden vector.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_INF ((int64_t)1 << 60)

typedef struct {
    int64_t key;
    int32_t node;
} heap_item;

typedef struct {
    heap_item *a;
    int n, cap;
} heap_t;

static void heap_push(heap_t *h, int64_t key, int32_t node) {
    if (h->n == h->cap) {
        h->cap = h->cap ? h->cap * 2 : 1024;
        h->a = (heap_item *)realloc(h->a, sizeof(heap_item) * (size_t)h->cap);
    }
    int i = h->n++;
    while (i > 0) {
        int p = (i - 1) >> 1;
        if (h->a[p].key <= key) break;
        h->a[i] = h->a[p];
        i = p;
    }
    h->a[i].key = key;
    h->a[i].node = node;
}

static heap_item heap_pop(heap_t *h) {
    heap_item top = h->a[0];
    heap_item last = h->a[--h->n];
    int i = 0;
    for (;;) {
        int c = 2 * i + 1;
        if (c >= h->n) break;
        if (c + 1 < h->n && h->a[c + 1].key < h->a[c].key) c++;
        if (h->a[c].key >= last.key) break;
        h->a[i] = h->a[c];
        i = c;
    }
    if (h->n > 0) h->a[i] = last;
    return top;
}

/* Solve. tail/head are 0-based. Returns 0, or a negative error code:
 *  -1 bad arguments, -2 out of memory, -3 the input graph has a directed cycle (not a tracking graph).
 * flow[e] in {0,1}; path_cost[k] = true cost of the k-th accepted path (first path_cap entries). */
int oracle_ssp_solve(int32_t n_nodes, int32_t n_arcs, const int32_t *tail, const int32_t *head, const int64_t *cost,
                     int32_t *flow, int32_t *n_paths, int64_t *path_cost, int32_t path_cap, int64_t *total_cost) {
    if (n_nodes < 2 || n_arcs < 0) return -1;
    const int32_t S = 0, T = n_nodes - 1;
    const int64_t m2 = 2 * (int64_t)n_arcs;
    int rc = 0;
    /* residual arcs: 2e = arc e forward (cap 1), 2e+1 = its reverse (cap 0) */
    int32_t *adj_ptr = (int32_t *)calloc((size_t)n_nodes + 1, sizeof(int32_t));
    int32_t *adj = (int32_t *)malloc(sizeof(int32_t) * (size_t)(m2 ? m2 : 1));
    int8_t *cap = (int8_t *)malloc((size_t)(m2 ? m2 : 1));
    int64_t *price = (int64_t *)malloc(sizeof(int64_t) * (size_t)n_nodes);
    int64_t *dist = (int64_t *)malloc(sizeof(int64_t) * (size_t)n_nodes);
    int32_t *par = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_nodes);
    int32_t *indeg = (int32_t *)calloc((size_t)n_nodes, sizeof(int32_t));
    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_nodes);
    int32_t *fill = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_nodes);
    heap_t heap = {0, 0, 0};
    if (!adj_ptr || !adj || !cap || !price || !dist || !par || !indeg || !order || !fill) {
        rc = -2;
        goto done;
    }
    for (int32_t e = 0; e < n_arcs; e++) {
        if (tail[e] < 0 || tail[e] >= n_nodes || head[e] < 0 || head[e] >= n_nodes) {
            rc = -1;
            goto done;
        }
        adj_ptr[tail[e] + 1]++;
        adj_ptr[head[e] + 1]++;
        indeg[head[e]]++;
    }
    for (int32_t v = 0; v < n_nodes; v++) adj_ptr[v + 1] += adj_ptr[v];
    memcpy(fill, adj_ptr, sizeof(int32_t) * (size_t)n_nodes);
    for (int32_t e = 0; e < n_arcs; e++) {
        adj[fill[tail[e]]++] = 2 * e;
        adj[fill[head[e]]++] = 2 * e + 1;
        cap[2 * e] = 1;
        cap[2 * e + 1] = 0;
    }
    /* topological order (Kahn) of the input DAG */
    {
        int32_t qh = 0, qt = 0;
        for (int32_t v = 0; v < n_nodes; v++)
            if (indeg[v] == 0) order[qt++] = v;
        while (qh < qt) {
            int32_t u = order[qh++];
            for (int32_t k = adj_ptr[u]; k < adj_ptr[u + 1]; k++) {
                int32_t a = adj[k];
                if (a & 1) continue;
                int32_t v = head[a >> 1];
                if (--indeg[v] == 0) order[qt++] = v;
            }
        }
        if (qt != n_nodes) {
            rc = -3;
            goto done;
        }
    }
    /* first tree: one pass in topological order (Graph.cpp:139-187) */
    for (int32_t v = 0; v < n_nodes; v++) {
        dist[v] = ORACLE_INF;
        par[v] = -1;
    }
    dist[S] = 0;
    for (int32_t i = 0; i < n_nodes; i++) {
        int32_t u = order[i];
        if (dist[u] >= ORACLE_INF) continue;
        for (int32_t k = adj_ptr[u]; k < adj_ptr[u + 1]; k++) {
            int32_t a = adj[k];
            if (a & 1) continue;
            int32_t v = head[a >> 1];
            int64_t d = dist[u] + cost[a >> 1];
            if (d < dist[v]) {
                dist[v] = d;
                par[v] = a;
            }
        }
    }
    int32_t K = 0;
    int64_t total = 0;
    int64_t max_finite = 0;
    if (flow) memset(flow, 0, sizeof(int32_t) * (size_t)n_arcs);
    while (dist[T] < ORACLE_INF) {
        /* true cost of this path = reduced distance + price difference; for k = 1 prices are 0 */
        int64_t true_cost = (K == 0) ? dist[T] : dist[T] + price[T] - price[S];
        if (K >= 1 && true_cost >= 0) break; /* wrapper.cpp:263-267 on integers */
        /* fold distances into prices (Graph.cpp:183-186; 547-556): unreachable nodes take the largest finite one */
        max_finite = 0;
        for (int32_t v = 0; v < n_nodes; v++)
            if (dist[v] < ORACLE_INF && dist[v] > max_finite) max_finite = dist[v];
        for (int32_t v = 0; v < n_nodes; v++) {
            int64_t d = dist[v] < ORACLE_INF ? dist[v] : max_finite;
            price[v] = (K == 0 ? 0 : price[v]) + d;
        }
        /* augment: reverse every arc of the path (Graph.cpp:266-335) */
        for (int32_t v = T; v != S;) {
            int32_t a = par[v];
            cap[a] = 0;
            cap[a ^ 1] = 1;
            if (flow) flow[a >> 1] = (a & 1) ? 0 : 1;
            v = (a & 1) ? head[a >> 1] : tail[a >> 1];
        }
        if (K < path_cap && path_cost) path_cost[K] = true_cost;
        total += true_cost;
        K++;
        /* next search: Dijkstra on reduced costs (Graph.cpp:394-503) */
        for (int32_t v = 0; v < n_nodes; v++) {
            dist[v] = ORACLE_INF;
            par[v] = -1;
        }
        dist[S] = 0;
        heap.n = 0;
        heap_push(&heap, 0, S);
        while (heap.n) {
            heap_item it = heap_pop(&heap);
            int32_t u = it.node;
            if (it.key != dist[u]) continue;
            for (int32_t k = adj_ptr[u]; k < adj_ptr[u + 1]; k++) {
                int32_t a = adj[k];
                if (!cap[a]) continue;
                int32_t e = a >> 1;
                int32_t v = (a & 1) ? tail[e] : head[e];
                int64_t c = (a & 1) ? -cost[e] : cost[e];
                int64_t rcost = c + price[u] - price[v];
                int64_t d = it.key + rcost;
                if (d < dist[v]) {
                    dist[v] = d;
                    par[v] = a;
                    heap_push(&heap, d, v);
                }
            }
        }
    }
    if (n_paths) *n_paths = K;
    if (total_cost) *total_cost = total;
done:
    free(adj_ptr);
    free(adj);
    free(cap);
    free(price);
    free(dist);
    free(par);
    free(indeg);
    free(order);
    free(fill);
    free(heap.a);
    return rc;
}

/* The same solve with SEVERAL paths per search - the rule the CUDA solver uses (csrc/solve_sm.cuh, "next path"),
 * restated here so that the CPU suite can check it against the one-path-per-search solve above on every golden graph
 * and on random ones. After a search the labels L are exact distances from S. Sinks (arcs u -> T with capacity) are
 * taken in ascending (L[u] + c, u) order; one is accepted when its path pays (or it is the very first path) and the
 * first-level branch of the search tree it lies in (the arc out of S its tree path starts with) carries no earlier
 * path of the batch: flipping a path leaves every reduced cost non-negative under L, so labels outside the flipped
 * branches stay exact and the ones inside are lower bounds. The first sink in a used branch ends the batch (search
 * again); a next sink that does not pay ends the solve. */
typedef struct {
    int64_t key;
    int32_t node, arc;
} sink_item;

static int sink_cmp(const void *a, const void *b) {
    const sink_item *x = (const sink_item *)a, *y = (const sink_item *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->node < y->node ? -1 : (x->node > y->node ? 1 : 0);
}

int oracle_ssp_solve_batched(int32_t n_nodes, int32_t n_arcs, const int32_t *tail, const int32_t *head,
                             const int64_t *cost, int32_t *flow, int32_t *n_paths, int64_t *path_cost, int32_t path_cap,
                             int64_t *total_cost, int32_t *n_searches) {
    if (n_nodes < 2 || n_arcs < 0) return -1;
    const int32_t S = 0, T = n_nodes - 1;
    const int64_t m2 = 2 * (int64_t)n_arcs;
    int rc = 0;
    int32_t *adj_ptr = (int32_t *)calloc((size_t)n_nodes + 1, sizeof(int32_t));
    int32_t *adj = (int32_t *)malloc(sizeof(int32_t) * (size_t)(m2 ? m2 : 1));
    int8_t *cap = (int8_t *)malloc((size_t)(m2 ? m2 : 1));
    int64_t *price = (int64_t *)calloc((size_t)n_nodes, sizeof(int64_t));
    int64_t *dist = (int64_t *)malloc(sizeof(int64_t) * (size_t)n_nodes);
    int32_t *par = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_nodes);
    int32_t *branch = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_nodes);
    int8_t *used = (int8_t *)malloc((size_t)n_nodes);
    int32_t *fill = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_nodes);
    int32_t *stack = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_nodes);
    sink_item *sinks = (sink_item *)malloc(sizeof(sink_item) * (size_t)(n_arcs ? n_arcs : 1));
    heap_t heap = {0, 0, 0};
    if (!adj_ptr || !adj || !cap || !price || !dist || !par || !branch || !used || !fill || !stack || !sinks) {
        rc = -2;
        goto done;
    }
    for (int32_t e = 0; e < n_arcs; e++) {
        if (tail[e] < 0 || tail[e] >= n_nodes || head[e] < 0 || head[e] >= n_nodes) {
            rc = -1;
            goto done;
        }
        adj_ptr[tail[e] + 1]++;
        adj_ptr[head[e] + 1]++;
    }
    for (int32_t v = 0; v < n_nodes; v++) adj_ptr[v + 1] += adj_ptr[v];
    memcpy(fill, adj_ptr, sizeof(int32_t) * (size_t)n_nodes);
    for (int32_t e = 0; e < n_arcs; e++) {
        adj[fill[tail[e]]++] = 2 * e;
        adj[fill[head[e]]++] = 2 * e + 1;
        cap[2 * e] = 1;
        cap[2 * e + 1] = 0;
    }
    int32_t K = 0, searches = 0;
    int64_t total = 0;
    if (flow) memset(flow, 0, sizeof(int32_t) * (size_t)n_arcs);
    /* first labels: Bellman-Ford passes over the input arcs (a DAG: settles in at most n passes; costs may be
     * negative, so no Dijkstra yet) */
    for (int32_t v = 0; v < n_nodes; v++) {
        dist[v] = ORACLE_INF;
        par[v] = -1;
    }
    dist[S] = 0;
    for (int32_t pass = 0; pass < n_nodes; pass++) {
        int moved = 0;
        for (int32_t e = 0; e < n_arcs; e++) {
            if (dist[tail[e]] >= ORACLE_INF) continue;
            int64_t d = dist[tail[e]] + cost[e];
            if (d < dist[head[e]]) {
                dist[head[e]] = d;
                par[head[e]] = 2 * e;
                moved = 1;
            }
        }
        if (!moved) break;
        if (pass == n_nodes - 1) {
            rc = -3;
            goto done;
        }
    }
    for (;;) {
        searches++;
        /* exact labels L = dist + price (price[S] stays 0), folded into the prices as the solve above does */
        int64_t max_finite = 0;
        for (int32_t v = 0; v < n_nodes; v++)
            if (dist[v] < ORACLE_INF && dist[v] > max_finite) max_finite = dist[v];
        /* first-level branch of every labelled node: the node behind S on its tree path */
        for (int32_t v = 0; v < n_nodes; v++) branch[v] = -2;
        branch[S] = -1;
        for (int32_t v = 0; v < n_nodes; v++) {
            if (branch[v] != -2 || dist[v] >= ORACLE_INF) continue;
            int32_t sp = 0, u = v;
            while (branch[u] == -2) {
                stack[sp++] = u;
                int32_t a = par[u];
                u = (a & 1) ? head[a >> 1] : tail[a >> 1];
            }
            int32_t b = (u == S) ? stack[sp - 1] : branch[u];
            while (sp) branch[stack[--sp]] = b;
        }
        /* sinks in ascending order of the true cost of their path */
        int32_t ns = 0;
        for (int32_t k = adj_ptr[T]; k < adj_ptr[T + 1]; k++) {
            int32_t a = adj[k];
            if (!(a & 1)) continue;       /* out-arcs of T (none in a tracking graph) */
            int32_t f = a ^ 1, e = a >> 1; /* the forward arc u -> T */
            if (!cap[f] || dist[tail[e]] >= ORACLE_INF) continue;
            sinks[ns].key = dist[tail[e]] + price[tail[e]] + cost[e];
            sinks[ns].node = tail[e];
            sinks[ns].arc = f;
            ns++;
        }
        qsort(sinks, (size_t)ns, sizeof(sink_item), sink_cmp);
        for (int32_t v = 0; v < n_nodes; v++) price[v] += dist[v] < ORACLE_INF ? dist[v] : max_finite;
        memset(used, 0, (size_t)n_nodes);
        int finished = 1, taken = 0;
        for (int32_t i = 0; i < ns; i++) {
            if (K >= 1 && sinks[i].key >= 0) break; /* wrapper.cpp:263-267; no later sink pays either */
            int32_t b = branch[sinks[i].node];
            if (b < 0) b = S; /* an arc S -> T: a branch of its own (not a tracking graph, but harmless) */
            if (used[b]) {
                finished = 0; /* stale labels in that branch: search again */
                break;
            }
            used[b] = 1;
            int32_t a = sinks[i].arc;
            cap[a] = 0;
            cap[a ^ 1] = 1;
            if (flow) flow[a >> 1] = 1;
            for (int32_t v = sinks[i].node; v != S;) {
                a = par[v];
                cap[a] = 0;
                cap[a ^ 1] = 1;
                if (flow) flow[a >> 1] = (a & 1) ? 0 : 1;
                v = (a & 1) ? head[a >> 1] : tail[a >> 1];
            }
            if (K < path_cap && path_cost) path_cost[K] = sinks[i].key;
            total += sinks[i].key;
            K++;
            taken++;
        }
        if (finished || !taken) break;
        /* next search: Dijkstra on reduced costs */
        for (int32_t v = 0; v < n_nodes; v++) {
            dist[v] = ORACLE_INF;
            par[v] = -1;
        }
        dist[S] = 0;
        heap.n = 0;
        heap_push(&heap, 0, S);
        while (heap.n) {
            heap_item it = heap_pop(&heap);
            int32_t u = it.node;
            if (it.key != dist[u]) continue;
            for (int32_t k = adj_ptr[u]; k < adj_ptr[u + 1]; k++) {
                int32_t a = adj[k];
                if (!cap[a]) continue;
                int32_t e = a >> 1;
                int32_t v = (a & 1) ? tail[e] : head[e];
                int64_t c = (a & 1) ? -cost[e] : cost[e];
                int64_t d = it.key + c + price[u] - price[v];
                if (d < dist[v]) {
                    dist[v] = d;
                    par[v] = a;
                    heap_push(&heap, d, v);
                }
            }
        }
    }
    if (n_paths) *n_paths = K;
    if (total_cost) *total_cost = total;
    if (n_searches) *n_searches = searches;
done:
    free(adj_ptr);
    free(adj);
    free(cap);
    free(price);
    free(dist);
    free(par);
    free(branch);
    free(used);
    free(fill);
    free(stack);
    free(sinks);
    free(heap.a);
    return rc;
}

const char *oracle_version(void) { return "mot3d_b200 oracle: int64 SSP restatement of muSSP/wrapper.cpp"; }
