// Stand-in for "pmc/pmc.h" (TEST INFRASTRUCTURE, see oracle/refshim/README.md).  PMC (LimHyungTae/pmc, tag libpmc) is an
// un-vendored dependency of the reference (3rdparty/pmc/pmc.cmake:28-29) and is not in this image.  The reference's
// graph.cc drives it as: k-core bound, heuristic clique, exact search -- "a maximum clique" comes back, which one is
// PMC's business.  NOT A PIN of A13: the exact search here is handed to a callback (the oracle's own max-clique routine,
// registered by oracle/refsolver.py), so that both sides pick the same clique among equal-size maxima and the rest of
// the reference's solver can be compared on identical inlier sets.  The k-core numbers are computed for real.
#ifndef REFSHIM_PMC_H
#define REFSHIM_PMC_H
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

extern "C" {
// adjacency as N x N bytes (row-major, symmetric, zero diagonal) -> clique vertices in out (room for N), returns its size
typedef int (*RefMaxCliqueFn)(const uint8_t *adj, int n, int *out);
extern RefMaxCliqueFn g_ref_max_clique;
}

using namespace std;          // the real pmc headers do this at global scope, and the reference's graph.cc relies on it

namespace pmc {

struct input {
    int algorithm = 0, threads = 1, experiment = 0, lb = 0, ub = 0, param_ub = 0, adj_limit = 0;
    double time_limit = 0, remove_time = 0;
    bool graph_stats = false, verbose = false, help = false, MCE = false, decreasing_order = false;
    std::string heu_strat, vertex_search_order;
};

class pmc_graph {
public:
    pmc_graph(const std::vector<long long> &vs, const std::vector<int> &es) : vertices(vs), edges(es) {}
    int num_vertices() const { return (int)vertices.size() - 1; }
    void create_adj() {}
    // core numbers by repeated minimum-degree removal; kcore has one leading slot like PMC's (graph.cc:76-81)
    void compute_cores()
    {
        const int n = num_vertices();
        std::vector<int> deg(n), core(n, 0);
        std::vector<char> gone(n, 0);
        for (int v = 0; v < n; ++v) deg[v] = (int)(vertices[v + 1] - vertices[v]);
        int k = 0;
        for (int it = 0; it < n; ++it) {
            int best = -1;
            for (int v = 0; v < n; ++v) if (!gone[v] && (best < 0 || deg[v] < deg[best])) best = v;
            k = std::max(k, deg[best]);
            core[best] = k; gone[best] = 1;
            for (long long e = vertices[best]; e < vertices[best + 1]; ++e) if (!gone[edges[e]]) --deg[edges[e]];
        }
        kcore.assign(1, 0);
        kcore.insert(kcore.end(), core.begin(), core.end());
        max_core = k;
    }
    int get_max_core() const { return max_core; }
    std::vector<int> *get_kcores() { return &kcore; }
    int solve(std::vector<int> &C) const
    {
        const int n = num_vertices();
        std::vector<uint8_t> adj((size_t)n * n, 0);
        for (int v = 0; v < n; ++v)
            for (long long e = vertices[v]; e < vertices[v + 1]; ++e) if (edges[e] != v) adj[(size_t)v * n + edges[e]] = adj[(size_t)edges[e] * n + v] = 1;
        std::vector<int> out((size_t)std::max(n, 1));
        const int m = g_ref_max_clique(adj.data(), n, out.data());
        C.assign(out.begin(), out.begin() + m);
        return m;
    }
    std::vector<long long> vertices;
    std::vector<int> edges, kcore;
    int max_core = 0;
};

class pmc_heu {
public:
    pmc_heu(pmc_graph &, input &) {}
    int search(pmc_graph &G, std::vector<int> &C) { return G.solve(C); }
};

class pmcx_maxclique {
public:
    pmcx_maxclique(pmc_graph &, input &) {}
    int search(pmc_graph &G, std::vector<int> &C) { return G.solve(C); }
    int search_dense(pmc_graph &G, std::vector<int> &C) { return G.solve(C); }
};

}  // namespace pmc
#endif
