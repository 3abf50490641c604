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

const char *oracle_version(void) { return "mot3d_b200 oracle: int64 SSP restatement of muSSP/wrapper.cpp"; }
