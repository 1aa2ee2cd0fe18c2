// refshim: mapMAP stand-in that RECORDS the model view_selection.cpp builds and hands back a trivial
// solution (per node: the offset of its cheapest label).  mapMAP itself is absent from the reference
// checkout, so the solver stays unpinned; what this pins is everything view_selection.cpp does around
// it: which edges enter the graph, label sets and their order, unary costs, the Potts weight, the
// termination and control parameters, and the decoding of offsets back to labels (see ../README.md).
#pragma once
#include <cstdint>
#include <functional>
#include <vector>

namespace mapmap {

typedef std::uint64_t luint_t;
template <typename C, unsigned W> using _iv_st = std::int32_t;   // index / label scalar
template <typename C, unsigned W> using _s_t = C;                // cost scalar
template <typename C> constexpr unsigned sys_max_simd_width() { return 1; }

enum TREE_SAMPLER_ALGORITHM { OPTIMISTIC_TREE_SAMPLER, LOCK_FREE_TREE_SAMPLER };

struct mapMAP_control {
    bool use_multilevel, use_spanning_tree, use_acyclic;
    unsigned spanning_tree_multilevel_after_n_iterations;
    bool force_acyclic;
    unsigned min_acyclic_iterations;
    bool relax_acyclic_maximal;
    TREE_SAMPLER_ALGORITHM tree_algorithm;
    bool sample_deterministic;
    std::uint64_t initial_seed;
    mapMAP_control() : use_multilevel(false), use_spanning_tree(false), use_acyclic(false), spanning_tree_multilevel_after_n_iterations(0),
        force_acyclic(false), min_acyclic_iterations(0), relax_acyclic_maximal(false), tree_algorithm(OPTIMISTIC_TREE_SAMPLER),
        sample_deterministic(false), initial_seed(0) {}
};

// what the glue reads back after tex::view_selection returned
struct Capture {
    std::uint64_t num_nodes;
    std::vector<std::uint32_t> edges;       // pairs, in add_edge order
    std::vector<float> edge_weight;
    std::vector<std::vector<std::int32_t> > labels;
    std::vector<std::vector<float> > costs;
    std::vector<int> unary_set;             // set_unary(i, ..) called
    float potts;
    unsigned window; double ratio;
    mapMAP_control ctr;
    bool components_updated, compress;
};
Capture& refshim_capture();   // defined in ref_glue.cpp

template <typename C>
class Graph {
public:
    explicit Graph(luint_t num_nodes) { Capture& c = refshim_capture(); c = Capture(); c.num_nodes = num_nodes; c.components_updated = false; }
    void add_edge(luint_t a, luint_t b, C w) {
        Capture& c = refshim_capture();
        c.edges.push_back(static_cast<std::uint32_t>(a)); c.edges.push_back(static_cast<std::uint32_t>(b)); c.edge_weight.push_back(w);
    }
    void update_components() { refshim_capture().components_updated = true; }
};

template <typename C, unsigned W>
class LabelSet {
public:
    LabelSet(luint_t num_nodes, bool compress) { Capture& c = refshim_capture(); c.labels.assign(num_nodes, std::vector<std::int32_t>()); c.compress = compress; }
    void set_label_set_for_node(luint_t node, std::vector<_iv_st<C, W> > const& l) { refshim_capture().labels[node] = l; }
    _iv_st<C, W> label_from_offset(luint_t node, _iv_st<C, W> offset) const { return refshim_capture().labels[node][offset]; }
};

template <typename C, unsigned W>
class UnaryTable {
    luint_t node;
public:
    UnaryTable(luint_t n, LabelSet<C, W>*) : node(n) { Capture& c = refshim_capture(); if (c.costs.size() <= n) c.costs.resize(n + 1); }
    void set_costs(std::vector<_s_t<C, W> > const& v) { refshim_capture().costs[node] = v; }
    luint_t node_id() const { return node; }
};

template <typename C, unsigned W>
class PairwisePotts {
public:
    explicit PairwisePotts(C w) { refshim_capture().potts = w; }
};

template <typename C, unsigned W>
class StopWhenReturnsDiminish {
public:
    StopWhenReturnsDiminish(unsigned window, double ratio) { Capture& c = refshim_capture(); c.window = window; c.ratio = ratio; }
};

template <typename C, unsigned W>
class mapMAP {
    std::function<void(const luint_t, const _iv_st<C, W>)> log;
public:
    void set_graph(Graph<C>*) {}
    void set_label_set(LabelSet<C, W>*) {}
    void set_unary(luint_t i, UnaryTable<C, W>* u) {
        Capture& c = refshim_capture();
        if (c.unary_set.size() < c.num_nodes) c.unary_set.assign(c.num_nodes, 0);
        c.unary_set[i] += (u->node_id() == i) ? 1 : 1000;
    }
    void set_pairwise(PairwisePotts<C, W>*) {}
    template <typename F> void set_logging_callback(F f) { log = f; }
    void set_termination_criterion(StopWhenReturnsDiminish<C, W>*) {}
    C optimize(std::vector<_iv_st<C, W> >& solution, mapMAP_control const& ctr) {
        Capture& c = refshim_capture();
        c.ctr = ctr;
        solution.assign(c.num_nodes, 0);
        for (luint_t i = 0; i < c.num_nodes; ++i) {
            std::vector<float> const& u = c.costs[i];
            std::size_t best = 0;
            for (std::size_t k = 1; k < u.size(); ++k) if (u[k] < u[best]) best = k;   // first minimum
            solution[i] = static_cast<_iv_st<C, W> >(best);
        }
        if (log) log(0, 0);
        return C(0);
    }
};

}  // namespace mapmap
