// mrf.cu -- K5-K7: pairwise-Potts MRF view selection on the device.
//
// Replaces tex::view_selection's call into mapMAP (libs/tex/view_selection.cpp:84-118); the model
// (label sets, unaries, Potts edges between seen faces) follows view_selection.cpp:26-82.
// Solver = block coordinate descent over induced forests with exact min-sum DP, the algorithm
// defined in oracle/mrf.c; this file reproduces it bit for bit:
//   * forest sampling is order independent (hash priorities, level-synchronous rounds)
//   * messages are summed in adjacency order in fp32 (no FMA: additions and minima only)
//   * the energy used for termination is 32.32 fixed point, summed with integer atomics.
//
// Execution (one iteration = 4 launches, NO host round trip: the stop rule is evaluated on the device
// and the host only polls a pinned flag a few iterations behind the launches it has queued):
//   k_forest   persistent cooperative kernel: root selection, `rounds` growth rounds on a frontier,
//              separated by grid.sync().  Every node that joins records (tree, slot) -- the tree of its
//              parent and its arrival number in that tree -- and adds its label count to the tree's
//              totals, so that the kernel can lay the forest out TREE BY TREE (levels ascending inside
//              a tree) without any sort: one block-aggregated allocation pass + one scatter pass.
//   k_tree<G>  the trees of an induced forest do not touch each other (every edge that leaves a tree
//              ends at a node whose label is fixed in this iteration), so the whole min-sum DP of a
//              tree -- bottom-up messages AND top-down assignment -- runs inside ONE warp on data staged
//              in shared memory: a CTA claims a chunk of trees, packs as many as fit into its shared
//              memory pool, every warp stages its trees with bulk async copies (cp.async.bulk = TMA,
//              one copy per node for the cost row and one for the view row, completion on an mbarrier),
//              builds per-node label bitmasks in shared memory (O(1) merge-join of the sorted label
//              lists by popcount), sweeps the levels up and down with __syncwarp() between levels, and
//              writes the new labels.  The messages H never exist in global memory: DRAM traffic per
//              forest node is 6 B per label (cost + view, read once) + ~60 B of node metadata, against
//              the 14 B per label of a sweep through global memory.  Trees that do not fit
//              (or contain a node of degree > 3) take the same recursion through global memory.
//   k_energy   fixed-point energy -> efix[t] on the device
//   k_stop     StopWhenReturnsDiminish (view_selection.cpp:84) on the device; once it fires, the
//              launches the host has already queued return immediately
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <memory>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace b2 {

namespace {

constexpr uint32_t LVL_NONE = 0xFFFFFFFFu;
constexpr uint32_t LVL_DEAD = 0xFFFFFFFEu;
constexpr uint32_t NO_NODE = 0xFFFFFFFFu;
constexpr int MAX_LEVELS = 1024;    // rounds + 1 must fit
constexpr int MAX_MASK_WORDS = 64;  // label bitmasks up to K = 2047 views, else binary search

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__host__ __device__ __forceinline__ uint32_t iter_seed(uint32_t seed, uint32_t t)
{
    return mix32(seed + 0x9E3779B9u * (t + 1u));
}
__device__ __forceinline__ uint32_t prio(uint32_t v, uint32_t seed_t) { return mix32(v ^ seed_t); }
__device__ __forceinline__ bool root_cand(uint32_t v, uint32_t seed_t, uint32_t rdiv)
{
    return mix32(prio(v, seed_t) ^ 0x68E31DA4u) % rdiv == 0;
}

struct Mrf {
    uint32_t F, nb, ne;          // nodes, owned node range
    const uint32_t *adj_ptr, *adj_idx;
    const uint4 *adj4;           // compact adjacency: x,y,z = neighbours, w = degree (CSR if > 3)
    const uint64_t *ptr;
    const uint16_t *view;
    const float *cost;
    float *H, *hminp1;           // global-memory DP scratch (oversize / high-degree trees only)
    uint32_t *amin;
    uint32_t *level, *labels, *lidx;   // lidx = position of the current label in the node's list
    uint32_t *order;             // forest nodes, tree by tree, levels ascending inside a tree
    uint16_t *olev;              // level of order[i]
    uint32_t *pos;               // index of a node in order, NO_NODE outside the forest
    uint2 *tjoin;                // per node: (tree, arrival number in the tree)
    uint4 *ttab;                 // per tree: (nodes, labels, first order index, flags: 1 = has a node of degree > 3)
    uint32_t *ctl;               // per-iteration control block (zeroed before k_forest), see CTL_*
    uint32_t *state;             // per-run state, see ST_*
    uint32_t *queue, *qstamp;    // frontier lists [2][F] and push de-duplication stamps [F]
    unsigned long long *efix;    // [max_iterations + 1] fixed-point energies
    uint32_t K, mask_words;
    uint32_t part_size, rounds, rdiv, seed, iter;
    uint32_t tree_smem;          // dynamic shared memory of k_tree (bytes)
};
// control block layout (uint32 words)
constexpr int CTL_QN = 0;                        // [MAX_LEVELS+1] frontier sizes per round
constexpr int CTL_NROOTS = MAX_LEVELS + 8;       // trees of this iteration
constexpr int CTL_CURSOR = MAX_LEVELS + 9;       // forest nodes laid out so far
constexpr int CTL_CLAIM = MAX_LEVELS + 10;       // k_tree: next unclaimed tree
constexpr int CTL_MAXPRIO = MAX_LEVELS + 12;     // 64-bit, 8-byte aligned
constexpr int CTL_WORDS = MAX_LEVELS + 32;
// run state (uint32 words)
constexpr int ST_STOP = 0;      // 0 while running, else the iteration the stop rule fired in
constexpr int ST_DONE = 1;      // last iteration whose energy is final
constexpr int ST_BAD = 2;       // labels > K found by k_label_check
constexpr int ST_UNSEEN = 3;    // faces with label 0
constexpr int ST_FNODES = 4;    // 64-bit: forest nodes summed over the iterations (roofline accounting)
constexpr int ST_FNNZ = 6;      // 64-bit: labels of forest nodes summed over the iterations
constexpr int ST_SLOW = 8;      // trees that went through global memory
constexpr int ST_WORDS = 16;

__device__ __forceinline__ bool same_part(const Mrf &m, uint32_t a, uint32_t b)
{
    if (m.part_size >= m.F) return true;  // single partition: no integer divisions on the hot path
    return a / m.part_size == b / m.part_size;
}
__device__ __forceinline__ bool owned(const Mrf &m, uint32_t v) { return v >= m.nb && v < m.ne; }
__device__ __forceinline__ bool local_pair(const Mrf &m, uint32_t v, uint32_t w)
{
    return owned(m, w) && same_part(m, v, w);
}

// Neighbour list of one node.  Face graphs of manifold meshes have degree <= 3: one 16-byte load
// replaces the adj_ptr -> adj_idx dependent chain; larger degrees fall back to the CSR arrays.
struct Nb { uint32_t deg, x, y, z, base; };
__device__ __forceinline__ Nb load_nb(const Mrf &m, uint32_t v)
{
    const uint4 a = __ldg(m.adj4 + v);
    Nb n; n.deg = a.w; n.x = a.x; n.y = a.y; n.z = a.z;
    n.base = a.w > 3 ? m.adj_ptr[v] : 0;
    return n;
}
__device__ __forceinline__ uint32_t nb_at(const Mrf &m, const Nb &n, uint32_t i)
{
    if (n.deg <= 3) return i == 0 ? n.x : (i == 1 ? n.y : n.z);
    return m.adj_idx[n.base + i];
}

__global__ void __launch_bounds__(256) k_build_adj4(uint32_t F, const uint32_t *__restrict__ adj_ptr,
                                                    const uint32_t *__restrict__ adj_idx, uint4 *adj4)
{
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= F) return;
    uint32_t a0 = adj_ptr[v], d = adj_ptr[v + 1] - a0;
    uint4 r = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, d);
    if (d <= 3) {
        if (d > 0) r.x = adj_idx[a0];
        if (d > 1) r.y = adj_idx[a0 + 1];
        if (d > 2) r.z = adj_idx[a0 + 2];
    }
    adj4[v] = r;
}

// ---- shared-memory / async-copy primitives (sm_90+ PTX; the host emulation replaces this block) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy accesses to shared memory before / async-proxy (bulk copy) writes after
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// ---- end of primitives ----

template <int G>
__global__ void __launch_bounds__(256) k_init_labels(Mrf m)
{
    const uint32_t lane = threadIdx.x & (G - 1);
    const uint32_t gpw = blockDim.x / G;
    for (uint32_t base = m.nb + blockIdx.x * gpw; base < m.ne; base += gridDim.x * gpw) {
        uint32_t v = base + threadIdx.x / G;
        bool act = v < m.ne;
        uint64_t p0 = act ? m.ptr[v] : 0, p1 = act ? m.ptr[v + 1] : 0;
        float bh = INFINITY;
        uint32_t bk = 0xFFFFFFFFu;
        for (uint64_t k = p0 + lane; k < p1; k += G) {
            float c = m.cost[k];
            if (c < bh) { bh = c; bk = (uint32_t)(k - p0); }
        }
        __syncwarp();
        for (int s = G / 2; s; s >>= 1) {
            float oh = __shfl_xor_sync(0xffffffffu, bh, s);
            uint32_t ok = __shfl_xor_sync(0xffffffffu, bk, s);
            if (oh < bh || (oh == bh && ok < bk)) { bh = oh; bk = ok; }
        }
        if (act && lane == 0) {
            m.labels[v] = (p1 > p0) ? (uint32_t)m.view[p0 + bk] + 1u : 0u;
            m.lidx[v] = (p1 > p0) ? bk : 0u;
        }
    }
}

// position of label `lab` (= view+1) in node w's sorted list, or -1
__device__ __forceinline__ long long find_label(const Mrf &m, uint32_t w, uint32_t lab)
{
    uint64_t lo = m.ptr[w], end = m.ptr[w + 1], hi = end;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        uint32_t l = (uint32_t)m.view[mid] + 1u;
        if (l < lab) lo = mid + 1; else hi = mid;
    }
    if (lo < end && (uint32_t)m.view[lo] + 1u == lab) return (long long)lo;
    return -1;
}

// ---- forest sampling + tree layout, one persistent cooperative launch -----------------------------
__device__ __forceinline__ uint32_t count_in_forest(const Mrf &m, uint32_t v, uint32_t r)
{
    uint32_t c = 0;
    const Nb nb = load_nb(m, v);
    for (uint32_t i = 0; i < nb.deg; ++i) {
        uint32_t w = nb_at(m, nb, i);
        if (local_pair(m, v, w) && __ldcg(m.level + w) < r) ++c;
    }
    return c;
}

constexpr int FOREST_THREADS = 1024;
__global__ void __launch_bounds__(FOREST_THREADS, 1) k_forest(Mrf m, int build_trees)
{
    cg::grid_group grid = cg::this_grid();
    if (__ldcg(m.state + ST_STOP)) return;  // the stop rule fired in an earlier iteration (grid-uniform)
    __shared__ uint32_t s_cnt, s_base, s_warp[FOREST_THREADS / 32];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const uint32_t seed_t = iter_seed(m.seed, m.iter);
    const uint32_t n_own = m.ne - m.nb;
    unsigned long long *maxprio = reinterpret_cast<unsigned long long *>(m.ctl + CTL_MAXPRIO);

    if (m.rdiv == 0) {  // single-root mode: the seen node with the largest priority
        unsigned long long key = 0;
        for (uint32_t v = m.nb + tid; v < m.ne; v += nth)
            if (m.labels[v] != 0) {
                unsigned long long k = (((unsigned long long)prio(v, seed_t)) << 1) | 1ull;
                key = k > key ? k : key;
            }
        for (int s = 16; s; s >>= 1) {
            unsigned long long o = __shfl_xor_sync(0xffffffffu, key, s);
            key = o > key ? o : key;
        }
        if ((threadIdx.x & 31) == 0 && key) atomicMax(maxprio, key);
        grid.sync();
    }
    // round 0: eligibility and roots (full scan, once).  Roots open a tree: compact tree numbers come from one
    // global atomic per block and pass (block-aggregated), not one per root.
    for (uint32_t base = blockIdx.x * blockDim.x; base < n_own; base += nth) {  // block-uniform trip count
        const uint32_t v = m.nb + base + threadIdx.x;
        const bool valid = v < m.ne;
        uint32_t lvl = LVL_DEAD;
        if (valid && m.labels[v] != 0) {
            const uint32_t pv = prio(v, seed_t);
            bool eligible = true, is_root;
            if (m.rdiv) is_root = root_cand(v, seed_t, m.rdiv);
            else is_root = __ldcg(maxprio) == ((((unsigned long long)pv) << 1) | 1ull);
            const Nb nb = load_nb(m, v);
            for (uint32_t i = 0; i < nb.deg; ++i) {
                uint32_t w = nb_at(m, nb, i);
                if (m.labels[w] == 0) continue;
                if (!local_pair(m, v, w)) { if (prio(w, seed_t) > pv) eligible = false; continue; }
                if (m.rdiv && is_root && root_cand(w, seed_t, m.rdiv) && prio(w, seed_t) > pv) is_root = false;
            }
            lvl = !eligible ? LVL_DEAD : (is_root ? 0u : LVL_NONE);
        }
        if (valid) { m.level[v] = lvl; if (build_trees) m.pos[v] = NO_NODE; }
        if (build_trees) {
            const bool is_root = valid && lvl == 0u;
            if (threadIdx.x == 0) s_cnt = 0;
            __syncthreads();
            uint32_t my = 0;
            if (is_root) my = atomicAdd(&s_cnt, 1u);
            __syncthreads();
            if (threadIdx.x == 0 && s_cnt) s_base = atomicAdd(&m.ctl[CTL_NROOTS], s_cnt);
            __syncthreads();
            if (is_root) {
                const uint32_t j = s_base + my;
                const uint32_t nl = (uint32_t)(m.ptr[v + 1] - m.ptr[v]);
                m.ttab[j] = make_uint4(1u, nl, 0u, __ldg(&m.adj4[v].w) > 3u ? 1u : 0u);
                m.tjoin[v] = make_uint2(j, 0u);
            }
        }
    }
    grid.sync();
    // Growth rounds on a FRONTIER instead of full scans.  Only an undecided node with >= 1 forest
    // neighbour can change state in a round, and such a node is either newly adjacent to a node that
    // joined in the previous round (pushed by that node) or a candidate that lost and re-queues
    // itself.  The evaluated set equals the set a full scan would act on, so levels are identical to
    // oracle/mrf.c; list order is irrelevant.  qstamp de-duplicates pushes (unique per iteration+round).
    const uint32_t stamp_base = m.iter * 2048u;
    auto push = [&](uint32_t w, uint32_t round) {
        if (atomicExch(m.qstamp + w, stamp_base + round) == stamp_base + round) return;
        cg::coalesced_group g = cg::coalesced_threads();
        uint32_t base = 0;
        if (g.thread_rank() == 0) base = atomicAdd(&m.ctl[CTL_QN + round], g.size());
        base = g.shfl(base, 0);
        m.queue[(size_t)(round & 1u) * m.F + base + g.thread_rank()] = w;
    };
    for (uint32_t v = m.nb + tid; v < m.ne; v += nth) {  // seed: undecided neighbours of the roots
        if (__ldcg(m.level + v) != 0u) continue;
        const Nb nb = load_nb(m, v);
        for (uint32_t i = 0; i < nb.deg; ++i) {
            uint32_t w = nb_at(m, nb, i);
            if (local_pair(m, v, w) && __ldcg(m.level + w) == LVL_NONE) push(w, 1u);
        }
    }
    grid.sync();
    for (uint32_t r = 1; r <= m.rounds; ++r) {
        const uint32_t n = __ldcg(m.ctl + CTL_QN + r);
        const uint32_t *q = m.queue + (size_t)(r & 1u) * m.F;
        for (uint32_t qi = tid; qi < n; qi += nth) {
            const uint32_t v = __ldcg(q + qi);
            if (__ldcg(m.level + v) != LVL_NONE) continue;
            const Nb nb = load_nb(m, v);
            uint32_t c = 0, parent = NO_NODE;
            for (uint32_t i = 0; i < nb.deg; ++i) {
                uint32_t w = nb_at(m, nb, i);
                if (local_pair(m, v, w) && __ldcg(m.level + w) < r) { ++c; parent = w; }
            }
            if (c >= 2) { m.level[v] = LVL_DEAD; continue; }
            if (c != 1) continue;
            const uint32_t pv = prio(v, seed_t);
            bool win = true;
            for (uint32_t i = 0; i < nb.deg && win; ++i) {
                uint32_t w = nb_at(m, nb, i);
                if (!local_pair(m, v, w)) continue;
                uint32_t lw = __ldcg(m.level + w);
                if (!(lw == LVL_NONE || lw == r)) continue;
                if (prio(w, seed_t) < pv) continue;
                if (count_in_forest(m, w, r) == 1) win = false;  // a stronger adjacent candidate: wait
            }
            if (win) {
                m.level[v] = r;
                if (build_trees) {   // the parent joined in an earlier round: its (tree, slot) is final
                    const uint2 pj = __ldcg(m.tjoin + parent);
                    uint4 *te = m.ttab + pj.x;
                    const uint32_t slot = atomicAdd(&te->x, 1u);
                    atomicAdd(&te->y, (uint32_t)(m.ptr[v + 1] - m.ptr[v]));
                    if (nb.deg > 3) atomicOr(&te->w, 1u);
                    m.tjoin[v] = make_uint2(pj.x, slot);
                }
                if (r < m.rounds)
                    for (uint32_t i = 0; i < nb.deg; ++i) {
                        uint32_t w = nb_at(m, nb, i);
                        if (local_pair(m, v, w) && __ldcg(m.level + w) == LVL_NONE) push(w, r + 1u);
                    }
            } else if (r < m.rounds) {
                push(v, r + 1u);
            }
        }
        grid.sync();
    }
    if (!build_trees) return;

    // ---- lay the forest out tree by tree: arrival numbers grow with the rounds, so the nodes of a tree are
    // ---- sorted by level; which tree comes first in memory is irrelevant (block-aggregated allocation)
    const uint32_t nroots = __ldcg(m.ctl + CTL_NROOTS);
    for (uint32_t base = blockIdx.x * blockDim.x; base < nroots; base += nth) {
        const uint32_t j = base + threadIdx.x;
        const uint32_t cnt = j < nroots ? __ldcg(&m.ttab[j].x) : 0u;
        uint32_t incl = cnt;
        for (int s = 1; s < 32; s <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, incl, s);
            if ((int)(threadIdx.x & 31) >= s) incl += o;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = threadIdx.x < (uint32_t)(FOREST_THREADS / 32) ? s_warp[threadIdx.x] : 0u, wi = w;
            for (int s = 1; s < 32; s <<= 1) {
                uint32_t o = __shfl_up_sync(0xffffffffu, wi, s);
                if ((int)threadIdx.x >= s) wi += o;
            }
            if (threadIdx.x < (uint32_t)(FOREST_THREADS / 32)) s_warp[threadIdx.x] = wi - w;  // exclusive warp offsets
            if (threadIdx.x == 31) s_base = wi ? atomicAdd(&m.ctl[CTL_CURSOR], wi) : 0u;
        }
        __syncthreads();
        if (j < nroots) m.ttab[j].z = s_base + s_warp[threadIdx.x >> 5] + incl - cnt;
        __syncthreads();
    }
    grid.sync();
    unsigned long long fn = 0, fz = 0;
    for (uint32_t v = m.nb + tid; v < m.ne; v += nth) {
        const uint32_t l = __ldcg(m.level + v);
        if (l > m.rounds) continue;
        const uint2 tj = __ldcg(m.tjoin + v);
        const uint32_t idx = __ldcg(&m.ttab[tj.x].z) + tj.y;
        m.order[idx] = v;
        m.olev[idx] = (uint16_t)l;
        m.pos[v] = idx;
        fn += 1ull;
        fz += m.ptr[v + 1] - m.ptr[v];
    }
    for (int s = 16; s; s >>= 1) { fn += __shfl_xor_sync(0xffffffffu, fn, s); fz += __shfl_xor_sync(0xffffffffu, fz, s); }
    if ((threadIdx.x & 31) == 0 && fn) {
        atomicAdd(reinterpret_cast<unsigned long long *>(m.state + ST_FNODES), fn);
        atomicAdd(reinterpret_cast<unsigned long long *>(m.state + ST_FNNZ), fz);
    }
}

// ---- min-sum DP of whole trees inside one warp ---------------------------------------------------------------
constexpr int TREE_THREADS = 128;
constexpr int TREE_WARPS = TREE_THREADS / 32;
constexpr int TREE_CHUNK = 16;       // trees claimed per global atomic
constexpr uint32_t NBR_SKIP = 0u, NBR_CHILD = 1u << 30, NBR_FIXED = 2u << 30, NBR_PARENT = 3u << 30;
constexpr uint32_t NBR_KIND = 3u << 30, NBR_ARG = ~NBR_KIND;

// upper bound of the shared memory one tree needs (the exact rows are padded to the 16-byte granules of the bulk
// copies: <= 6 extra floats and <= 14 extra u16 per node)
__host__ __device__ __forceinline__ uint32_t tree_hcap(uint32_t cnt, uint32_t nnz) { return (nnz + 6u * cnt + 3u) & ~3u; }
__host__ __device__ __forceinline__ uint32_t tree_vcap(uint32_t cnt, uint32_t nnz) { return (nnz + 14u * cnt + 7u) & ~7u; }
__host__ __device__ __forceinline__ uint32_t tree_node_bytes(uint32_t W) { return 40u + 6u * W; }

struct TreeStatic {
    unsigned long long mbar[TREE_CHUNK];
    uint32_t mphase[TREE_CHUNK];
    uint32_t t_cnt[TREE_CHUNK], t_nnz[TREE_CHUNK], t_start[TREE_CHUNK], t_flags[TREE_CHUNK];
    uint32_t t_node0[TREE_CHUNK], t_h0[TREE_CHUNK], t_v0[TREE_CHUNK], t_slow[TREE_CHUNK];
    uint32_t chunk_first, sb_n, sb_nodes, sb_hcap, sb_vcap;
};

// pointers into the dynamic shared memory of one sub-batch
struct TreePool {
    float *H;          // [hcap]   cost rows, turned into the min-sum tables in place
    uint16_t *V;       // [vcap]   view rows
    uint32_t *gid, *hoff, *voff, *nbr, *am, *lab, *mask;
    float *hm;
    uint16_t *nlab, *lev, *mpre;
};
__device__ __forceinline__ TreePool carve_pool(unsigned char *base, uint32_t nodes, uint32_t hcap, uint32_t vcap, uint32_t W)
{
    TreePool p;
    p.H = reinterpret_cast<float *>(base); base += (size_t)hcap * 4;
    p.V = reinterpret_cast<uint16_t *>(base); base += (size_t)vcap * 2;
    p.gid = reinterpret_cast<uint32_t *>(base); base += (size_t)nodes * 4;
    p.hoff = reinterpret_cast<uint32_t *>(base); base += (size_t)nodes * 4;
    p.voff = reinterpret_cast<uint32_t *>(base); base += (size_t)nodes * 4;
    p.nbr = reinterpret_cast<uint32_t *>(base); base += (size_t)nodes * 12;
    p.am = reinterpret_cast<uint32_t *>(base); base += (size_t)nodes * 4;
    p.lab = reinterpret_cast<uint32_t *>(base); base += (size_t)nodes * 4;
    p.hm = reinterpret_cast<float *>(base); base += (size_t)nodes * 4;
    p.mask = reinterpret_cast<uint32_t *>(base); base += (size_t)nodes * 4 * W;
    p.nlab = reinterpret_cast<uint16_t *>(base); base += (size_t)nodes * 2;
    p.lev = reinterpret_cast<uint16_t *>(base); base += (size_t)nodes * 2;
    p.mpre = reinterpret_cast<uint16_t *>(base);
    return p;
}

// first index of the run of equal levels that ends at `end` (exclusive), searched inside [a, end)
template <typename LevPtr>
__device__ __forceinline__ uint32_t level_run_begin(LevPtr lev, uint32_t a, uint32_t end, uint32_t lane)
{
    const uint32_t L = lev[end - 1];
    uint32_t s = end;
    for (;;) {
        const bool p = s > a + lane && lev[s - 1 - lane] == L;
        const uint32_t b = __ballot_sync(0xffffffffu, p);
        const uint32_t run = b == 0xFFFFFFFFu ? 32u : (uint32_t)__ffs((int)~b) - 1u;
        s -= run;
        if (run < 32u) break;
    }
    return s;
}
// end (exclusive) of the run of equal levels that starts at `s`, searched inside [s, b)
template <typename LevPtr>
__device__ __forceinline__ uint32_t level_run_end(LevPtr lev, uint32_t s, uint32_t b, uint32_t lane)
{
    const uint32_t L = lev[s];
    uint32_t e = s;
    for (;;) {
        const bool p = e + lane < b && lev[e + lane] == L;
        const uint32_t bits = __ballot_sync(0xffffffffu, p);
        const uint32_t run = bits == 0xFFFFFFFFu ? 32u : (uint32_t)__ffs((int)~bits) - 1u;
        e += run;
        if (run < 32u) break;
    }
    return e;
}

// position of label `lab` in a shared-memory row: bitmask + prefix popcount, or binary search without masks
__device__ __forceinline__ int row_find(const TreePool &p, uint32_t W, uint32_t li, uint32_t lab)
{
    if (W) {
        const uint32_t word = lab >> 5, bit = lab & 31u;
        if (word >= W) return -1;
        const uint32_t bits = p.mask[(size_t)li * W + word];
        if (!((bits >> bit) & 1u)) return -1;
        return (int)((uint32_t)p.mpre[(size_t)li * W + word] + __popc(bits & ((1u << bit) - 1u)));
    }
    const uint16_t *row = p.V + p.voff[li];
    int lo = 0, n = (int)p.nlab[li], hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if ((uint32_t)row[mid] + 1u < lab) lo = mid + 1; else hi = mid;
    }
    return (lo < n && (uint32_t)row[lo] + 1u == lab) ? lo : -1;
}

// stage one tree: node tables + one bulk copy per node and array; the copies complete on `bar`
__device__ void tree_stage(const Mrf &m, const TreePool &p, uint32_t start, uint32_t cnt, uint32_t node0, uint32_t h0,
                           uint32_t v0, unsigned long long *bar, uint32_t lane)
{
    uint32_t hcarry = h0, vcarry = v0, tx = 0;
    for (uint32_t c = 0; c < cnt; c += 32) {
        const uint32_t i = c + lane;
        const bool valid = i < cnt;
        uint32_t v = 0, lv = 0, n = 0, hsz = 0, vsz = 0;
        uint64_t p0 = 0;
        uint4 a4 = make_uint4(NO_NODE, NO_NODE, NO_NODE, 0u);
        if (valid) {
            v = m.order[start + i];
            lv = m.olev[start + i];
            p0 = m.ptr[v];
            n = (uint32_t)(m.ptr[v + 1] - p0);
            a4 = __ldg(m.adj4 + v);
            hsz = ((uint32_t)(p0 & 3u) + n + 3u) & ~3u;
            vsz = ((uint32_t)(p0 & 7u) + n + 7u) & ~7u;
        }
        uint32_t hi = hsz, vi = vsz;
        for (int s = 1; s < 32; s <<= 1) {
            const uint32_t oh = __shfl_up_sync(0xffffffffu, hi, s), ov = __shfl_up_sync(0xffffffffu, vi, s);
            if ((int)lane >= s) { hi += oh; vi += ov; }
        }
        const uint32_t ho = hcarry + hi - hsz, vo = vcarry + vi - vsz;
        hcarry += __shfl_sync(0xffffffffu, hi, 31);
        vcarry += __shfl_sync(0xffffffffu, vi, 31);
        if (valid) {
            bulk_g2s(p.H + ho, m.cost + (p0 & ~(uint64_t)3), hsz * 4u, bar);
            bulk_g2s(p.V + vo, m.view + (p0 & ~(uint64_t)7), vsz * 2u, bar);
            tx += hsz * 4u + vsz * 2u;
            const uint32_t li = node0 + i;
            p.gid[li] = v;
            p.hoff[li] = ho + (uint32_t)(p0 & 3u);
            p.voff[li] = vo + (uint32_t)(p0 & 7u);
            p.nlab[li] = (uint16_t)n;
            p.lev[li] = (uint16_t)lv;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const uint32_t w = a == 0 ? a4.x : (a == 1 ? a4.y : a4.z);
                uint32_t enc = NBR_SKIP;
                if ((uint32_t)a < a4.w && w != NO_NODE) {
                    const uint32_t x = m.labels[w];
                    if (x != 0u) {  // unseen faces carry no edges (view_selection.cpp:30,35)
                        const uint32_t pw = m.pos[w];
                        if (pw != NO_NODE) {   // a forest neighbour is in the same tree: deeper = child
                            const uint32_t lj = pw - start;
                            enc = (lj > i ? NBR_CHILD : NBR_PARENT) | (node0 + lj);
                        } else enc = NBR_FIXED | x;
                    }
                }
                p.nbr[3 * (size_t)li + a] = enc;
            }
        }
    }
    for (int s = 16; s; s >>= 1) tx += __shfl_xor_sync(0xffffffffu, tx, s);
    // the only arrival of this phase comes after every copy was issued, so the phase cannot complete early
    if (lane == 0) mbar_arrive_expect_tx(bar, tx);
}

// the DP of one staged tree: nodes [a, b) of the pool, sorted by level
template <int G>
__device__ void tree_solve_smem(const Mrf &m, const TreePool &p, uint32_t a, uint32_t b, uint32_t W, uint32_t lane)
{
    constexpr uint32_t GPW = 32 / G;
    const uint32_t glane = lane & (G - 1), sub = lane / G;
    // label bitmasks of this tree's rows
    if (W) {
        for (uint32_t i = a * W + lane; i < b * W; i += 32) p.mask[i] = 0u;
        __syncwarp();
        for (uint32_t base = a; base < b; base += GPW) {
            const uint32_t li = base + sub;
            if (li < b) {
                const uint16_t *row = p.V + p.voff[li];
                const uint32_t n = p.nlab[li];
                for (uint32_t k = glane; k < n; k += G) {
                    const uint32_t lab = (uint32_t)row[k] + 1u;
                    atomicOr(&p.mask[(size_t)li * W + (lab >> 5)], 1u << (lab & 31u));
                }
            }
        }
        __syncwarp();
        for (uint32_t li = a + lane; li < b; li += 32) {
            uint32_t seen = 0;
            for (uint32_t w = 0; w < W; ++w) {
                p.mpre[(size_t)li * W + w] = (uint16_t)seen;
                seen += __popc(p.mask[(size_t)li * W + w]);
            }
        }
        __syncwarp();
    }
    // bottom-up: deepest level first
    for (uint32_t end = b; end > a;) {
        const uint32_t s = level_run_begin(p.lev, a, end, lane);
        for (uint32_t base = s; base < end; base += GPW) {
            const uint32_t li = base + sub;
            const bool act = li < end;
            float bh = INFINITY;
            uint32_t bk = 0xFFFFFFFFu;
            if (act) {
                const uint32_t n = p.nlab[li];
                float *Hv = p.H + p.hoff[li];
                const uint16_t *viewv = p.V + p.voff[li];
                uint32_t kind[3], arg[3], cho[3] = {0, 0, 0};
                float chm[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const uint32_t e = p.nbr[3 * (size_t)li + i];
                    kind[i] = e & NBR_KIND; arg[i] = e & NBR_ARG;
                    if (kind[i] == NBR_CHILD) { chm[i] = p.hm[arg[i]]; cho[i] = p.hoff[arg[i]]; }
                }
                for (uint32_t k = glane; k < n; k += G) {
                    const uint32_t lab = (uint32_t)viewv[k] + 1u;
                    float h = Hv[k];
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        if (kind[i] == NBR_CHILD) {  // Potts message min(h_w(lab), hmin_w + 1)
                            float msg = chm[i];
                            const int j = row_find(p, W, arg[i], lab);
                            if (j >= 0) { const float hw = p.H[cho[i] + (uint32_t)j]; if (hw < msg) msg = hw; }
                            h = h + msg;
                        } else if (kind[i] == NBR_FIXED) {
                            h = h + (lab != arg[i] ? 1.0f : 0.0f);
                        }
                    }
                    Hv[k] = h;
                    if (h < bh) { bh = h; bk = k; }
                }
            }
            for (int sft = G / 2; sft; sft >>= 1) {
                const float oh = __shfl_xor_sync(0xffffffffu, bh, sft);
                const uint32_t ok = __shfl_xor_sync(0xffffffffu, bk, sft);
                if (oh < bh || (oh == bh && ok < bk)) { bh = oh; bk = ok; }
            }
            if (act && glane == 0) { p.hm[li] = bh + 1.0f; p.am[li] = bk; }
        }
        __syncwarp();
        end = s;
    }
    // top-down: shallowest level first, one lane per node
    for (uint32_t s = a; s < b;) {
        const uint32_t e = level_run_end(p.lev, s, b, lane);
        for (uint32_t li = s + lane; li < e; li += 32) {
            uint32_t bk = p.am[li];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const uint32_t en = p.nbr[3 * (size_t)li + i];
                if ((en & NBR_KIND) == NBR_PARENT) {
                    const uint32_t xp = p.lab[en & NBR_ARG];
                    const int j = row_find(p, W, li, xp);
                    if (j >= 0 && p.H[p.hoff[li] + (uint32_t)j] <= p.hm[li]) bk = (uint32_t)j;
                }
            }
            p.am[li] = bk;
            p.lab[li] = (uint32_t)p.V[p.voff[li] + bk] + 1u;
        }
        __syncwarp();
        s = e;
    }
    for (uint32_t li = a + lane; li < b; li += 32) {
        const uint32_t v = p.gid[li];
        m.labels[v] = p.lab[li];
        m.lidx[v] = p.am[li];
    }
}

// the same recursion through global memory: any degree, any size (one node at a time, 32 lanes over its labels)
__device__ void tree_solve_global(const Mrf &m, uint32_t start, uint32_t cnt, uint32_t lane)
{
    const uint16_t *lev = m.olev + start;
    for (uint32_t end = cnt; end > 0;) {
        const uint32_t s = level_run_begin(lev, 0u, end, lane);
        for (uint32_t i = s; i < end; ++i) {
            const uint32_t v = m.order[start + i];
            const uint64_t p0 = m.ptr[v], p1 = m.ptr[v + 1];
            const Nb nb = load_nb(m, v);
            float bh = INFINITY;
            uint32_t bk = 0xFFFFFFFFu;
            for (uint64_t k = p0 + lane; k < p1; k += 32) {
                const uint32_t lab = (uint32_t)m.view[k] + 1u;
                float h = m.cost[k];
                for (uint32_t q = 0; q < nb.deg; ++q) {
                    const uint32_t w = nb_at(m, nb, q);
                    const uint32_t x = m.labels[w];
                    if (x == 0) continue;
                    const uint32_t pw = m.pos[w];
                    if (pw != NO_NODE) {
                        if (pw > start + i) {  // child
                            float msg = __ldcg(m.hminp1 + w);
                            const long long j = find_label(m, w, lab);
                            if (j >= 0) { const float hw = __ldcg(m.H + j); if (hw < msg) msg = hw; }
                            h = h + msg;
                        }
                    } else {
                        h = h + (lab != x ? 1.0f : 0.0f);
                    }
                }
                m.H[k] = h;
                if (h < bh) { bh = h; bk = (uint32_t)(k - p0); }
            }
            for (int sft = 16; sft; sft >>= 1) {
                const float oh = __shfl_xor_sync(0xffffffffu, bh, sft);
                const uint32_t ok = __shfl_xor_sync(0xffffffffu, bk, sft);
                if (oh < bh || (oh == bh && ok < bk)) { bh = oh; bk = ok; }
            }
            if (lane == 0) { m.hminp1[v] = bh + 1.0f; m.amin[v] = bk; }
        }
        __syncwarp();
        end = s;
    }
    for (uint32_t s = 0; s < cnt;) {
        const uint32_t e = level_run_end(lev, s, cnt, lane);
        for (uint32_t i = s + lane; i < e; i += 32) {
            const uint32_t v = m.order[start + i];
            uint32_t bk = __ldcg(m.amin + v);
            const Nb nb = load_nb(m, v);
            for (uint32_t q = 0; q < nb.deg; ++q) {
                const uint32_t w = nb_at(m, nb, q);
                const uint32_t pw = m.pos[w];
                if (m.labels[w] != 0 && pw != NO_NODE && pw < start + i) {  // the parent: assigned one level earlier
                    const uint32_t xp = __ldcg(m.labels + w);
                    const long long j = find_label(m, v, xp);
                    if (j >= 0 && __ldcg(m.H + j) <= __ldcg(m.hminp1 + v)) bk = (uint32_t)(j - (long long)m.ptr[v]);
                    break;
                }
            }
            m.labels[v] = (uint32_t)m.view[m.ptr[v] + bk] + 1u;
            m.lidx[v] = bk;
        }
        __syncwarp();
        s = e;
    }
}

template <int G>
__global__ void __launch_bounds__(TREE_THREADS) k_tree(Mrf m)
{
    if (__ldcg(m.state + ST_STOP)) return;
    extern __shared__ __align__(16) unsigned char tree_dyn[];
    __shared__ TreeStatic ts;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t W = m.mask_words;
    const uint32_t nroots = m.ctl[CTL_NROOTS];
    if (threadIdx.x < TREE_CHUNK) { mbar_init(&ts.mbar[threadIdx.x], 1u); ts.mphase[threadIdx.x] = 0u; }
    if (threadIdx.x == 0) fence_mbar_init();
    __syncthreads();
    for (;;) {
        if (threadIdx.x == 0) ts.chunk_first = atomicAdd(&m.ctl[CTL_CLAIM], (uint32_t)TREE_CHUNK);
        __syncthreads();
        const uint32_t first = ts.chunk_first;
        if (first >= nroots) break;
        const uint32_t nchunk = min((uint32_t)TREE_CHUNK, nroots - first);
        if (threadIdx.x < nchunk) {
            const uint4 t = m.ttab[first + threadIdx.x];
            ts.t_cnt[threadIdx.x] = t.x; ts.t_nnz[threadIdx.x] = t.y; ts.t_start[threadIdx.x] = t.z; ts.t_flags[threadIdx.x] = t.w;
        }
        __syncthreads();
        for (uint32_t done = 0; done < nchunk;) {
            if (threadIdx.x == 0) {   // pack the next trees of the chunk into the pool, in order
                uint32_t n = 0, nodes = 0, hcap = 0, vcap = 0, slow = 0;
                while (done + n < nchunk) {
                    const uint32_t i = done + n, cnt = ts.t_cnt[i], nnz = ts.t_nnz[i];
                    const uint32_t hc = tree_hcap(cnt, nnz), vc = tree_vcap(cnt, nnz);
                    const uint64_t alone = 4ull * hc + 2ull * vc + (uint64_t)cnt * tree_node_bytes(W);
                    if ((ts.t_flags[i] & 1u) || alone > m.tree_smem || cnt > 16384u) {   // through global memory
                        ts.t_slow[i] = 1u; ts.t_node0[i] = 0; ts.t_h0[i] = 0; ts.t_v0[i] = 0;
                        ++n; ++slow;
                        continue;
                    }
                    const uint64_t bytes = 4ull * (hcap + hc) + 2ull * (vcap + vc) + (uint64_t)(nodes + cnt) * tree_node_bytes(W);
                    if (bytes > m.tree_smem) break;
                    ts.t_slow[i] = 0u; ts.t_node0[i] = nodes; ts.t_h0[i] = hcap; ts.t_v0[i] = vcap;
                    nodes += cnt; hcap += hc; vcap += vc;
                    ++n;
                }
                ts.sb_n = n; ts.sb_nodes = nodes; ts.sb_hcap = hcap; ts.sb_vcap = vcap;
                if (slow) atomicAdd(m.state + ST_SLOW, slow);
            }
            __syncthreads();
            const uint32_t sb_n = ts.sb_n;
            const TreePool pool = carve_pool(tree_dyn, ts.sb_nodes, ts.sb_hcap, ts.sb_vcap, W);
            fence_proxy_async();   // the previous sub-batch's generic accesses to the pool precede the bulk copies
            // every warp first stages all of its trees (the copies of the later ones overlap the DP of the earlier)
            for (uint32_t i = done + warp; i < done + sb_n; i += TREE_WARPS)
                if (!ts.t_slow[i])
                    tree_stage(m, pool, ts.t_start[i], ts.t_cnt[i], ts.t_node0[i], ts.t_h0[i], ts.t_v0[i], &ts.mbar[i - done], lane);
            for (uint32_t i = done + warp; i < done + sb_n; i += TREE_WARPS) {
                if (ts.t_slow[i]) { tree_solve_global(m, ts.t_start[i], ts.t_cnt[i], lane); continue; }
                const uint32_t slot = i - done;
                mbar_wait(&ts.mbar[slot], ts.mphase[slot]);
                __syncwarp();
                if (lane == 0) ts.mphase[slot] ^= 1u;
                tree_solve_smem<G>(m, pool, ts.t_node0[i], ts.t_node0[i] + ts.t_cnt[i], W, lane);
            }
            __syncthreads();
            done += sb_n;
        }
    }
}

// 32.32 fixed-point energy of the owned nodes: unaries + edges counted by their lower endpoint -> efix[slot]
__global__ void __launch_bounds__(256) k_energy(Mrf m, uint32_t slot)
{
    if (__ldcg(m.state + ST_STOP)) return;
    unsigned long long e = 0;
    for (uint32_t v = m.nb + blockIdx.x * blockDim.x + threadIdx.x; v < m.ne; v += gridDim.x * blockDim.x) {
        uint32_t x = m.labels[v];
        if (x == 0) { e += 1ull << 32; continue; }
        e += (unsigned long long)(long long)((double)m.cost[m.ptr[v] + m.lidx[v]] * 4294967296.0);
        const Nb nb = load_nb(m, v);
        for (uint32_t i = 0; i < nb.deg; ++i) {
            uint32_t w = nb_at(m, nb, i);
            uint32_t xw = m.labels[w];
            if (w > v && xw != 0 && xw != x) e += 1ull << 32;
        }
    }
    for (int s = 16; s; s >>= 1) e += __shfl_xor_sync(0xffffffffu, e, s);
    if ((threadIdx.x & 31) == 0 && e) atomicAdd(m.efix + slot, e);
}

// StopWhenReturnsDiminish(window, ratio) (view_selection.cpp:84), the double arithmetic of oracle/mrf.c
__global__ void k_stop(Mrf m, uint32_t t, uint32_t window, float ratio, uint32_t max_iterations)
{
    if (m.state[ST_STOP]) return;
    m.state[ST_DONE] = t;
    bool stop = t >= max_iterations;
    if (t >= window) {
        const double e0 = (double)(long long)m.efix[t - window], e1 = (double)(long long)m.efix[t];
        if (e0 <= 0.0 || (e0 - e1) / e0 < (double)ratio) stop = true;
    }
    if (stop) m.state[ST_STOP] = t;
}

// label range check + unseen count (view_selection.cpp:121-132)
__global__ void __launch_bounds__(256) k_label_check(Mrf m)
{
    uint32_t bad = 0, unseen = 0;
    for (uint32_t v = m.nb + blockIdx.x * blockDim.x + threadIdx.x; v < m.ne; v += gridDim.x * blockDim.x) {
        const uint32_t x = m.labels[v];
        if (m.K && x > m.K) ++bad;
        if (x == 0) ++unseen;
    }
    for (int s = 16; s; s >>= 1) { bad += __shfl_xor_sync(0xffffffffu, bad, s); unseen += __shfl_xor_sync(0xffffffffu, unseen, s); }
    if ((threadIdx.x & 31) == 0) {
        if (bad) atomicAdd(m.state + ST_BAD, bad);
        if (unseen) atomicAdd(m.state + ST_UNSEEN, unseen);
    }
}

Mrf make_mrf(b2tex_ctx *c, uint32_t iter)
{
    Mrf m;
    m.F = c->F; m.nb = c->face_begin; m.ne = c->face_end;
    m.adj_ptr = c->adj_ptr.p; m.adj_idx = c->adj_idx.p;
    m.adj4 = c->mrf_adj4.p;
    m.ptr = c->dc_ptr.p; m.view = c->dc_view.p; m.cost = c->dc_cost.p;
    m.H = c->mrf_H.p; m.hminp1 = c->mrf_hminp1.p; m.amin = c->mrf_amin.p; m.level = c->mrf_level.p;
    m.labels = c->labels.p; m.lidx = c->mrf_lidx.p;
    m.order = c->mrf_order.p; m.olev = c->mrf_olev.p; m.pos = c->mrf_pos.p;
    m.tjoin = c->mrf_tjoin.p; m.ttab = c->mrf_ttab.p;
    m.ctl = c->mrf_ctl.p; m.state = c->mrf_state.p;
    m.queue = c->mrf_queue.p; m.qstamp = c->mrf_queue.p + 2 * (size_t)c->F;
    m.efix = c->mrf_energy.p;
    m.K = c->K;
    m.mask_words = c->mrf_mask_words;
    const b2tex_mrf_params &p = c->mrf_params;
    uint32_t P = p.num_parts ? p.num_parts : 1;
    m.part_size = (c->F + P - 1) / P; if (!m.part_size) m.part_size = 1;
    m.rounds = p.rounds;
    if (p.root_div == 0) m.rdiv = 0;
    else { uint32_t cap = c->F / 8u; if (cap < 1u) cap = 1u; m.rdiv = p.root_div < cap ? p.root_div : cap; }
    m.seed = p.seed;
    m.iter = iter;
    m.tree_smem = c->mrf_tree_smem;
    return m;
}

template <typename K>
int coop_grid(b2tex_ctx *c, K kernel, size_t smem, int *grid, int threads = 256)
{
    int per_sm = 0;
    B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    if (per_sm < 1) { set_error("mrf kernel cannot be resident"); return B2TEX_ERR_CUDA; }
    *grid = c->num_sms * per_sm;
    return B2TEX_OK;
}

int launch_forest(b2tex_ctx *c, Mrf &m, int build_trees)
{
    cudaStream_t s = c->stream;
    int grid = 0;
    // few fat blocks: the cost of grid.sync() grows with the number of blocks
    B2_TRY(coop_grid(c, k_forest, 0, &grid, FOREST_THREADS));
    if (grid > c->num_sms) grid = c->num_sms;  // one fat block per SM: cheapest grid.sync()
    if (const char *e = getenv("B2TEX_FOREST_BLOCKS_PER_SM")) grid = c->num_sms * std::max(1, atoi(e));
    uint32_t n = m.ne - m.nb;
    int need = (int)((n + FOREST_THREADS - 1) / FOREST_THREADS);
    if (grid > need) grid = need > 0 ? need : 1;
    B2_CUDA(cudaMemsetAsync(m.ctl, 0, CTL_WORDS * sizeof(uint32_t), s));
    void *args[] = {&m, &build_trees};
    B2_CUDA(cudaLaunchCooperativeKernel((void *)k_forest, dim3(grid), dim3(FOREST_THREADS), args, 0, s));
    return B2TEX_OK;
}

template <int G>
int launch_tree(b2tex_ctx *c, Mrf &m)
{
    static bool attr_set = false;   // opt in to > 48 KB of dynamic shared memory (per function, once)
    if (!attr_set) {
        B2_CUDA(cudaFuncSetAttribute(k_tree<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    int per_sm = 0;
    B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tree<G>, TREE_THREADS, m.tree_smem));
    if (per_sm < 1) { set_error("k_tree cannot be resident with %u bytes of shared memory", m.tree_smem); return B2TEX_ERR_CUDA; }
    const int grid = c->num_sms * per_sm;
    k_tree<G><<<grid, TREE_THREADS, m.tree_smem, c->stream>>>(m);
    B2_KERNEL_CHECK();
    return B2TEX_OK;
}

int enqueue_iteration(b2tex_ctx *c, Mrf &m)
{
    cudaStream_t s = c->stream;
    if (m.ne <= m.nb) return B2TEX_OK;
    {
        ScopedTimer tf(c, "mrf.k_forest", 20.0 * (double)(m.ne - m.nb));
        B2_TRY(launch_forest(c, m, 1));
    }
    {
        ScopedTimer tu(c, "mrf.k_tree");   // bytes are filled in after the run (forest coverage is known then)
        switch (c->mrf_group) {
            case 4: B2_TRY(launch_tree<4>(c, m)); break;
            case 8: B2_TRY(launch_tree<8>(c, m)); break;
            case 16: B2_TRY(launch_tree<16>(c, m)); break;
            default: B2_TRY(launch_tree<32>(c, m)); break;
        }
    }
    ScopedTimer te(c, "mrf.k_energy", 12.0 * (double)(m.ne - m.nb));
    k_energy<<<std::max(1, c->num_sms * 8), 256, 0, s>>>(m, m.iter);
    B2_KERNEL_CHECK();
    return B2TEX_OK;
}

int alloc_mrf(b2tex_ctx *c, const b2tex_mrf_params *p)
{
    if (!c->have_costs) { set_error("view selection: data costs missing"); return B2TEX_ERR_ARG; }
    if (!c->have_adj) { set_error("view selection: adjacency missing"); return B2TEX_ERR_ARG; }
    if (p->rounds + 2 > (uint32_t)MAX_LEVELS) { set_error("mrf rounds too large"); return B2TEX_ERR_ARG; }
    c->mrf_params = *p;
    const size_t F = c->F;
    B2_TRY(c->mrf_H.alloc(c->nnz));
    B2_TRY(c->mrf_hminp1.alloc(F));
    B2_TRY(c->mrf_amin.alloc(F));
    B2_TRY(c->mrf_level.alloc(F));
    B2_TRY(c->mrf_lidx.alloc(F));
    B2_TRY(c->mrf_order.alloc(F));
    B2_TRY(c->mrf_olev.alloc(F));
    B2_TRY(c->mrf_pos.alloc(F));
    B2_CUDA(cudaMemsetAsync(c->mrf_pos.p, 0xFF, F * sizeof(uint32_t), c->stream));  // nodes of other ranks never enter a tree
    B2_TRY(c->mrf_tjoin.alloc(F));
    B2_TRY(c->mrf_ttab.alloc(F));
    B2_TRY(c->mrf_queue.alloc(3 * F));   // frontier lists [2][F] | qstamp [F]
    B2_TRY(c->mrf_queue.zero(c->stream));
    B2_TRY(c->mrf_ctl.alloc(CTL_WORDS));
    B2_TRY(c->mrf_state.alloc(ST_WORDS));
    B2_TRY(c->mrf_state.zero(c->stream));
    B2_TRY(c->mrf_energy.alloc((size_t)p->max_iterations + 2));
    B2_TRY(c->mrf_energy.zero(c->stream));
    B2_TRY(c->mrf_adj4.alloc(F));
    if (F) k_build_adj4<<<(unsigned)((F + 255) / 256), 256, 0, c->stream>>>((uint32_t)F, c->adj_ptr.p, c->adj_idx.p, c->mrf_adj4.p);
    if (!c->have_labels || c->labels.n != F) { B2_TRY(c->labels.alloc(F)); B2_TRY(c->labels.zero(c->stream)); }
    uint32_t nodes = c->face_end - c->face_begin;
    double rho = nodes ? (double)c->nnz / nodes : 0.0;
    // lanes per node: a level of one tree holds only a few nodes, so wide groups idle on short label lists
    c->mrf_group = rho >= 40 ? 32 : rho >= 12 ? 16 : rho >= 6 ? 8 : 4;
    if (const char *g = getenv("B2TEX_MRF_GROUP")) {
        int v = atoi(g);
        if (v == 4 || v == 8 || v == 16 || v == 32) c->mrf_group = v;
    }
    // label bitmasks: labels are view+1 <= K
    uint32_t words = (c->K + 1 + 31) / 32;
    static const bool no_masks = getenv("B2TEX_NO_MASKS") != nullptr;
    c->mrf_mask_words = (c->K == 0 || words > (uint32_t)MAX_MASK_WORDS || no_masks) ? 0 : words;
    // shared-memory pool of k_tree: room for a few average trees (~ rounds * 2.7 nodes) per CTA
    uint32_t smem = 72 * 1024;
    if (const char *e = getenv("B2TEX_TREE_SMEM_KB")) smem = (uint32_t)std::max(8, std::min(200, atoi(e))) * 1024u;
    c->mrf_tree_smem = smem;
    return B2TEX_OK;
}

int read_energy(b2tex_ctx *c, const Mrf &m, uint32_t slot, int64_t *efix)
{
    unsigned long long e = 0;
    B2_CUDA(cudaMemcpyAsync(&e, m.efix + slot, sizeof(e), cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    *efix = (int64_t)e;
    return B2TEX_OK;
}

}  // namespace

int mrf_init(b2tex_ctx *c, const b2tex_mrf_params *p, int64_t *efix)
{
    B2_TRY(alloc_mrf(c, p));
    Mrf m = make_mrf(c, 0);
    cudaStream_t s = c->stream;
    const int grid = std::max(1, c->num_sms * 8);
    if (m.ne > m.nb) {
        ScopedTimer tm(c, "mrf_init");
        switch (c->mrf_group) {
            case 4: k_init_labels<4><<<grid, 256, 0, s>>>(m); break;
            case 8: k_init_labels<8><<<grid, 256, 0, s>>>(m); break;
            case 16: k_init_labels<16><<<grid, 256, 0, s>>>(m); break;
            default: k_init_labels<32><<<grid, 256, 0, s>>>(m); break;
        }
        B2_KERNEL_CHECK();
    }
    c->have_labels = true;
    c->mrf_ready = true;
    if (m.ne > m.nb) k_energy<<<grid, 256, 0, s>>>(m, 0u);
    B2_KERNEL_CHECK();
    return read_energy(c, m, 0u, efix);
}

// one iteration, energy read back (the building block a host-driven sharded loop uses)
int mrf_iterate(b2tex_ctx *c, uint32_t t, int64_t *efix)
{
    if (!c->mrf_ready) { set_error("mrf_iterate before mrf_init"); return B2TEX_ERR_ARG; }
    if (t == 0 || t > c->mrf_params.max_iterations) { set_error("mrf_iterate: iterations are numbered from 1 to max_iterations"); return B2TEX_ERR_ARG; }
    Mrf m = make_mrf(c, t);
    B2_CUDA(cudaMemsetAsync(m.efix + t, 0, sizeof(unsigned long long), c->stream));
    B2_TRY(enqueue_iteration(c, m));
    return read_energy(c, m, t, efix);
}

// The whole run without a host round trip per iteration: the host queues iterations ahead of the device; k_stop
// evaluates the stop rule on the device and turns the launches that are already queued behind it into no-ops.
int mrf_run(b2tex_ctx *c, const b2tex_mrf_params *p, b2tex_mrf_info *info, double *trace)
{
    int64_t e0 = 0;
    B2_TRY(mrf_init(c, p, &e0));
    cudaStream_t s = c->stream;
    const uint32_t max_it = p->max_iterations, window = p->window ? p->window : 1u;
    info->sweep_bytes = 14ull * c->nnz + 20ull * c->F;
    if (c->face_end <= c->face_begin) {   // nothing owned: the energy is constant, the stop rule fires at `window`
        const uint32_t t_end = std::min(window, max_it);
        info->iterations = t_end; info->unseen = 0;
        info->energy_initial = info->energy_final = (double)e0 / 4294967296.0;
        if (trace) for (uint32_t t = 0; t <= t_end; ++t) trace[t] = info->energy_initial;
        return B2TEX_OK;
    }
    constexpr int LAG = 3;   // iterations queued beyond the last one whose stop flag the host has seen
    if (!c->mrf_host_flags) B2_CUDA(cudaHostAlloc((void **)&c->mrf_host_flags, 64 * sizeof(uint32_t), cudaHostAllocDefault));
    cudaEvent_t ev[LAG + 1];
    for (auto &e : ev) B2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    volatile uint32_t *hf = c->mrf_host_flags;
    int rc = B2TEX_OK;
    for (uint32_t t = 1; t <= max_it; ++t) {
        Mrf m = make_mrf(c, t);
        rc = enqueue_iteration(c, m);
        if (rc != B2TEX_OK) break;
        k_stop<<<1, 1, 0, s>>>(m, t, window, p->ratio, max_it);
        const int slot = (int)(t % (LAG + 1));
        if (cudaMemcpyAsync((void *)&hf[slot], m.state + ST_STOP, 4, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaEventRecord(ev[slot], s) != cudaSuccess) {
            set_error("view selection: %s", cudaGetErrorString(cudaGetLastError()));
            rc = B2TEX_ERR_CUDA;
            break;
        }
        if (t > (uint32_t)LAG) {
            const int w = (int)((t - LAG) % (LAG + 1));
            if (cudaEventSynchronize(ev[w]) != cudaSuccess) { set_error("view selection: %s", cudaGetErrorString(cudaGetLastError())); rc = B2TEX_ERR_CUDA; break; }
            if (hf[w]) break;   // the rule fired LAG iterations ago; what is queued behind it returns at once
        }
    }
    if (cudaStreamSynchronize(s) != cudaSuccess && rc == B2TEX_OK) {
        set_error("view selection: %s", cudaGetErrorString(cudaGetLastError()));
        rc = B2TEX_ERR_CUDA;
    }
    for (auto &e : ev) cudaEventDestroy(e);
    B2_TRY(rc);
    Mrf m = make_mrf(c, 0);
    k_label_check<<<std::max(1, c->num_sms * 4), 256, 0, s>>>(m);
    B2_KERNEL_CHECK();
    uint32_t st[ST_WORDS];
    B2_TRY(c->mrf_state.download(st, ST_WORDS, s));
    std::vector<unsigned long long> efix((size_t)max_it + 2, 0ull);
    B2_TRY(c->mrf_energy.download(efix.data(), (size_t)max_it + 1, s));
    B2_CUDA(cudaStreamSynchronize(s));
    const uint32_t t_end = st[ST_STOP] ? st[ST_STOP] : st[ST_DONE];   // ST_STOP == 0 only for max_iterations == 0
    info->iterations = t_end;
    info->energy_initial = (double)(int64_t)efix[0] / 4294967296.0;
    info->energy_final = (double)(int64_t)efix[t_end] / 4294967296.0;
    info->unseen = st[ST_UNSEEN];
    if (trace) for (uint32_t t = 0; t <= t_end; ++t) trace[t] = (double)(int64_t)efix[t] / 4294967296.0;
    // roofline accounting of k_tree: SURVEY 8d's sweep formula restricted to the nodes the launches processed
    unsigned long long fn, fz;
    memcpy(&fn, st + ST_FNODES, 8); memcpy(&fz, st + ST_FNNZ, 8);
    c->mrf_forest_nodes = fn; c->mrf_forest_nnz = fz; c->mrf_slow_trees = st[ST_SLOW];
    if (c->profile && t_end) {
        const double per = (14.0 * (double)fz + 20.0 * (double)fn) / (double)t_end;
        for (auto &k : c->timers) if (!strcmp(k.name, "mrf.k_tree") && k.bytes == 0.0) k.bytes = per;
    }
    if (st[ST_BAD]) { set_error("Incorrect labeling"); return B2TEX_ERR_LABELING; }
    return B2TEX_OK;
}

// fixed-point energy of the owned nodes with the labels currently in the context (a sharded run calls
// this after the boundary-label exchange, so that cut edges see the neighbours' NEW labels)
int mrf_energy_only(b2tex_ctx *c, int64_t *efix)
{
    if (!c->mrf_ready) { set_error("mrf_energy before mrf_init"); return B2TEX_ERR_ARG; }
    Mrf m = make_mrf(c, 1);
    const uint32_t slot = c->mrf_params.max_iterations + 1;   // scratch slot
    B2_CUDA(cudaMemsetAsync(m.efix + slot, 0, sizeof(unsigned long long), c->stream));
    if (m.ne > m.nb) k_energy<<<std::max(1, c->num_sms * 8), 256, 0, c->stream>>>(m, slot);
    B2_KERNEL_CHECK();
    return read_energy(c, m, slot, efix);
}

int mrf_sample_only(b2tex_ctx *c, const b2tex_mrf_params *p, uint32_t t, uint32_t *level_host)
{
    if (!c->mrf_ready) { int64_t e; B2_TRY(mrf_init(c, p, &e)); }
    c->mrf_params = *p;
    Mrf m = make_mrf(c, t);
    B2_TRY(c->mrf_queue.zero(c->stream));  // the same (iteration, round) stamps may be replayed
    if (m.ne > m.nb) B2_TRY(launch_forest(c, m, 0));
    B2_TRY(c->mrf_level.download(level_host, c->F, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return B2TEX_OK;
}

}  // namespace b2
