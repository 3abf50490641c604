// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or called from the product path.
//
// extern "C" host for the *unmodified* vendored muSSP solver
// (/root/reference/mussp-master-6cf61b8.zip!/muSSP-master/muSSP/{Graph,Node,Sink}.{h,cpp}).
// It stands in for the reference's Boost.Python module, which cannot be built here (no Boost.Python):
//   - mussp_ref_solve        : reference mot3d/solvers/wrappers/muSSP/wrapper.cpp:171-317 (solve) fed arc arrays,
//                              flow read-out as in wrapper.cpp:139-169 (make_output2)
//   - mussp_ref_solve_text   : same, but parsing "a tail head w" lines with sscanf as wrapper.cpp:39-88 (init) does,
//                              so the text round-trip the reference pays can be timed too.
// Node ids are 1-based on this interface exactly like the text the reference hands to wrapper.solve
// (mot3d/mot3d.py:218-226); the shim subtracts 1 as wrapper.cpp:75 does.
// Built by oracle/build_ref.py into oracle/_ref/libmussp_ref.so (git-ignored). No reference source is copied.
#include <cstdio>
#include <cstring>
#include <ctime>
#include <iostream>
#include <memory>
#include <streambuf>
#include <string>
#include <vector>

#include "Graph.h"

namespace {

// Graph::invalid_edge_rm (Graph.cpp:110) and the parsers print unconditionally; keep stdout clean.
struct CoutSilencer {
    std::streambuf *saved;
    CoutSilencer() : saved(std::cout.rdbuf(nullptr)) {}
    ~CoutSilencer() { std::cout.rdbuf(saved); }
};

struct RefResult {
    std::vector<double> path_cost;
    std::vector<std::vector<int>> path_set;
    double core_seconds = 0;
};

// The iteration of wrapper.cpp:186-290: clip dummy edges, DAG shortest-path tree, reduced weights,
// first path always kept, then tree update / sink refresh / extract until the next path cost > -1e-7.
void run_reference_iteration(Graph &g, RefResult &out) {
    clock_t t0 = clock();
    g.invalid_edge_rm();
    g.shortest_path_dag();
    out.path_cost.push_back(g.distance2src[g.sink_id_]);
    g.cur_path_max_cost = -g.distance2src[g.sink_id_];
    g.update_allgraph_weights();
    g.extract_shortest_path();
    out.path_set.push_back(g.shortest_path);
    std::vector<int> nodes4update;
    g.find_node_set4update(nodes4update);
    g.flip_path();
    for (;;) {
        g.update_shortest_path_tree_recursive(nodes4update);
        g.update_sink_info(nodes4update);
        g.extract_shortest_path();
        double cur = out.path_cost.back() + g.distance2src[g.sink_id_];
        if (cur > -0.0000001) break;
        out.path_cost.push_back(cur);
        g.cur_path_max_cost = -cur;
        out.path_set.push_back(g.shortest_path);
        g.update_subgraph_weights(nodes4update);
        g.find_node_set4update(nodes4update);
        g.flip_path();
    }
    out.core_seconds = double(clock() - t0) / CLOCKS_PER_SEC;
}

// make_output2 (wrapper.cpp:139-169): toggle a flag per arc for every hop of every stored path.
void read_out_flow(Graph &g, const RefResult &r, int *flow) {
    std::vector<bool> flag(g.num_edges_);
    for (const auto &path : r.path_set)
        for (size_t j = 0; j + 1 < path.size(); j++) {
            int e = g.node_id2edge_id[Graph::node_key(path[j + 1], path[j])];
            flag[e] = !flag[e];
        }
    for (int e = 0; e < g.num_edges_; e++) flow[e] = flag[e] ? 1 : 0;
}

int finish(Graph &g, int *flow, int *n_paths, double *path_cost, int path_cap, double *total_cost,
           double *core_seconds) {
    RefResult r;
    run_reference_iteration(g, r);
    if (flow) read_out_flow(g, r, flow);
    double sum = 0;
    for (double c : r.path_cost) sum += c;
    if (n_paths) *n_paths = (int)r.path_cost.size();
    if (total_cost) *total_cost = sum;
    if (core_seconds) *core_seconds = r.core_seconds;
    if (path_cost)
        for (int i = 0; i < path_cap && i < (int)r.path_cost.size(); i++) path_cost[i] = r.path_cost[i];
    return 0;
}

}  // namespace

extern "C" {

int mussp_ref_solve(int n_nodes, int n_arcs, const int *tail, const int *head, const double *weight, int *flow,
                    int *n_paths, double *path_cost, int path_cap, double *total_cost, double *core_seconds) {
    CoutSilencer quiet;
    Graph g(n_nodes, n_arcs, 0, n_nodes - 1, 0.0, 0.0);
    for (int e = 0; e < n_arcs; e++) g.add_edge(tail[e] - 1, head[e] - 1, e, weight[e]);
    return finish(g, flow, n_paths, path_cost, path_cap, total_cost, core_seconds);
}

// `text` holds '\n'-separated lines, first line "p min V E".
int mussp_ref_solve_text(const char *text, int *flow, int flow_cap, int *n_paths, double *path_cost, int path_cap,
                         double *total_cost, double *core_seconds, double *parse_seconds) {
    CoutSilencer quiet;
    clock_t t0 = clock();
    const char *p = text;
    const char *nl = strchr(p, '\n');
    std::string first(p, nl ? (size_t)(nl - p) : strlen(p));
    int n = 0, m = 0;
    char pr_type[8];
    if (sscanf(first.c_str(), "%*c %3s %d %d", pr_type, &n, &m) != 3) return -1;
    if (flow && flow_cap < m) return -2;
    Graph g(n, m, 0, n - 1, 0.0, 0.0);
    int edge_id = 0;
    p = nl ? nl + 1 : p + strlen(p);
    std::string line;
    while (*p) {
        nl = strchr(p, '\n');
        size_t len = nl ? (size_t)(nl - p) : strlen(p);
        line.assign(p, len);
        if (!line.empty() && (line[0] == 'a' || line[0] == 'p')) {
            int t = 0, h = 0;
            double w = 0;
            sscanf(line.c_str(), "%*c %d %d %lf", &t, &h, &w);
            g.add_edge(t - 1, h - 1, edge_id, w);
            edge_id++;
        }
        p += len + (nl ? 1 : 0);
    }
    if (parse_seconds) *parse_seconds = double(clock() - t0) / CLOCKS_PER_SEC;
    if (edge_id != m) return -3;
    return finish(g, flow, n_paths, path_cost, path_cap, total_cost, core_seconds);
}

const char *mussp_ref_version() { return "muSSP 6cf61b8 (vendored zip), g++ -O2 -std=c++14"; }

}  // extern "C"
