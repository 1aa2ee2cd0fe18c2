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
// Execution (one iteration = 4 launches, no host round trips except the energy read-back):
//   k_forest   persistent cooperative kernel: root selection, `rounds` growth rounds separated by
//              grid.sync(), then bucketing of the forest nodes by level (deepest level first,
//              every level padded to a multiple of 32 entries) -- all in one launch
//   k_up<G>    persistent DATAFLOW kernel: warps claim 32/G nodes at a time in bucket order; a node
//              waits on per-node flags of its children instead of a grid-wide level barrier, so the
//              sweep is bounded by tree depth x node latency, not by 33 launches/barriers.  G lanes
//              (4..32, from the mean label count) stride over the node's sorted label list; child
//              messages are looked up through per-node label bitmasks + prefix popcounts (O(1)
//              merge-join) or by binary search when K is large; min/argmin by warp shuffles
//   k_down     persistent dataflow kernel, one thread per node waiting on its parent's flag
//   k_energy   fixed-point energy
#include <cooperative_groups.h>
#include <cub/cub.cuh>
#include <stdlib.h>

#include <memory>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace b2 {

namespace {

constexpr uint32_t LVL_NONE = 0xFFFFFFFFu;
constexpr uint32_t LVL_DEAD = 0xFFFFFFFEu;
constexpr uint32_t NO_NODE = 0xFFFFFFFFu;
constexpr int MAX_LEVELS = 1024;  // rounds + 1 must fit (shared-memory bucketing)
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
    float *H, *hminp1;
    uint32_t *amin, *level, *labels, *order;
    uint32_t *ctl;               // control block, see CTL_* offsets
    uint32_t *flag_up, *flag_dn;
    uint32_t *queue, *qstamp;    // frontier lists [2][F] and push de-duplication stamps [F]
    uint32_t *sort_key_in, *sort_key_out, *sort_val_in, *sort_val_out;
    const uint32_t *mask;        // [F][mask_words] label bitmask (bit = label), or null
    const uint16_t *mpre;        // [F][mask_words] labels in lower words
    unsigned long long *energy;
    uint32_t mask_words;
    uint32_t part_size, rounds, rdiv, seed, iter;
};
// control block layout (uint32 words)
constexpr int CTL_CNT = 0;                      // [MAX_LEVELS] per-level counts
constexpr int CTL_CUR = MAX_LEVELS;             // [MAX_LEVELS] per-level fill cursors
constexpr int CTL_OFF = 2 * MAX_LEVELS;         // [MAX_LEVELS+1] padded bucket offsets (deepest first)
constexpr int CTL_TOTAL = 3 * MAX_LEVELS + 8;   // padded number of order entries
constexpr int CTL_UTOTAL = 3 * MAX_LEVELS + 9;  // forest nodes (unpadded)
constexpr int CTL_MAXPRIO = 3 * MAX_LEVELS + 12;  // 64-bit, 8-byte aligned
constexpr int CTL_QN = 3 * MAX_LEVELS + 16;      // [MAX_LEVELS+1] frontier sizes per round
constexpr int CTL_WORDS = 4 * MAX_LEVELS + 32;

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

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// polling load without acquire semantics (no L1 invalidation per poll); pair with ONE acquire fence
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(uint32_t *p, uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

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
        if (act && lane == 0) m.labels[v] = (p1 > p0) ? (uint32_t)m.view[p0 + bk] + 1u : 0u;
    }
}

// per-node label bitmask + number of labels in lower words (one thread per node)
__global__ void __launch_bounds__(256) k_build_masks(Mrf m, uint32_t *mask, uint16_t *mpre)
{
    uint32_t v = m.nb + blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= m.ne) return;
    const uint32_t W = m.mask_words;
    uint32_t *mw = mask + (size_t)v * W;
    uint16_t *mp = mpre + (size_t)v * W;
    uint64_t k = m.ptr[v], end = m.ptr[v + 1];
    uint32_t seen = 0;
    for (uint32_t w = 0; w < W; ++w) {
        uint32_t bits = 0;
        while (k < end) {
            uint32_t lab = (uint32_t)m.view[k] + 1u;
            if ((lab >> 5) != w) break;
            bits |= 1u << (lab & 31);
            ++k;
        }
        mw[w] = bits;
        mp[w] = (uint16_t)seen;
        seen += __popc(bits);
    }
}

// position of label `lab` (= view+1) in node w's sorted list, or -1
__device__ __forceinline__ long long find_label(const Mrf &m, uint32_t w, uint32_t lab)
{
    if (m.mask) {
        const uint32_t word = lab >> 5, bit = lab & 31;
        if (word >= m.mask_words) return -1;
        const uint32_t mw = __ldg(m.mask + (size_t)w * m.mask_words + word);
        if (!((mw >> bit) & 1u)) return -1;
        const uint32_t pre = __ldg(m.mpre + (size_t)w * m.mask_words + word);
        return (long long)(m.ptr[w] + pre + __popc(mw & ((1u << bit) - 1u)));
    }
    uint64_t lo = m.ptr[w], end = m.ptr[w + 1], hi = end;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        uint32_t l = (uint32_t)m.view[mid] + 1u;
        if (l < lab) lo = mid + 1; else hi = mid;
    }
    if (lo < end && (uint32_t)m.view[lo] + 1u == lab) return (long long)lo;
    return -1;
}

// ---- forest sampling + bucketing, one persistent cooperative launch ------------------------------
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

__global__ void __launch_bounds__(1024, 1) k_forest(Mrf m, int do_bucket)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ uint32_t sm[];  // [rounds+1] level counts
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const uint32_t seed_t = iter_seed(m.seed, m.iter);
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
    // round 0: eligibility and roots (full scan, once)
    for (uint32_t v = m.nb + tid; v < m.ne; v += nth) {
        if (m.labels[v] == 0) { m.level[v] = LVL_DEAD; continue; }
        const uint32_t pv = prio(v, seed_t);
        bool eligible = true, is_root;
        if (m.rdiv) is_root = root_cand(v, seed_t, m.rdiv);
        else is_root = *maxprio == ((((unsigned long long)pv) << 1) | 1ull);
        const Nb nb = load_nb(m, v);
        for (uint32_t i = 0; i < nb.deg; ++i) {
            uint32_t w = nb_at(m, nb, i);
            if (m.labels[w] == 0) continue;
            if (!local_pair(m, v, w)) { if (prio(w, seed_t) > pv) eligible = false; continue; }
            if (m.rdiv && is_root && root_cand(w, seed_t, m.rdiv) && prio(w, seed_t) > pv) is_root = false;
        }
        m.level[v] = !eligible ? LVL_DEAD : (is_root ? 0u : LVL_NONE);
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
            uint32_t c = 0;
            for (uint32_t i = 0; i < nb.deg; ++i) {
                uint32_t w = nb_at(m, nb, i);
                if (local_pair(m, v, w) && __ldcg(m.level + w) < r) ++c;
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
    if (!do_bucket) return;

    // ---- bucket the forest nodes by level: deepest level first, levels padded to 32 entries ----
    uint32_t *cnt = sm;
    for (uint32_t i = threadIdx.x; i <= m.rounds; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    for (uint32_t v = m.nb + tid; v < m.ne; v += nth) {
        uint32_t l = m.level[v];
        if (l <= m.rounds) atomicAdd(&cnt[l], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i <= m.rounds; i += blockDim.x)
        if (cnt[i]) atomicAdd(&m.ctl[CTL_CNT + i], cnt[i]);
    grid.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        uint32_t acc = 0, uacc = 0;
        for (int l = (int)m.rounds; l >= 0; --l) {
            m.ctl[CTL_OFF + l] = acc;     // padded start of level l (deepest level first)
            m.ctl[CTL_CUR + l] = uacc;    // unpadded start of level l in the sorted list
            acc += (m.ctl[CTL_CNT + l] + 31u) & ~31u;
            uacc += m.ctl[CTL_CNT + l];
        }
        m.ctl[CTL_TOTAL] = acc;
        m.ctl[CTL_UTOTAL] = uacc;
    }
    // sort keys: a stable radix sort by (rounds - level) keeps node ids ascending inside a level, so
    // that consecutive order entries touch neighbouring rows of every per-node array (the sweeps are
    // bound by scattered DRAM sectors, not by arithmetic)
    for (uint32_t v = m.nb + tid; v < m.ne; v += nth) {
        uint32_t l = m.level[v];
        m.sort_key_in[v - m.nb] = l <= m.rounds ? m.rounds - l : m.rounds + 1u;
        m.sort_val_in[v - m.nb] = v;
    }
}

// sorted (level-major, node-ascending) list -> order array with every level padded to 32 entries
__global__ void __launch_bounds__(256) k_scatter_order(Mrf m)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.ctl[CTL_UTOTAL]) return;
    const uint32_t l = m.rounds - m.sort_key_out[i];
    m.order[m.ctl[CTL_OFF + l] + (i - m.ctl[CTL_CUR + l])] = m.sort_val_out[i];
}

// ---- bottom-up min-sum messages, dataflow over the forest -------------------------------------------
template <int G>
__global__ void __launch_bounds__(256, 8) k_up(Mrf m)
{
    constexpr uint32_t GPW = 32 / G;  // nodes per warp
    const uint32_t lane = threadIdx.x & (G - 1);
    const uint32_t sub = (threadIdx.x & 31) / G;
    const uint32_t total = m.ctl[CTL_TOTAL];
    const uint32_t stamp = m.iter;
    // Static round-robin over warps, chunks in bucket order (deepest level first).  A chunk only
    // depends on EARLIER chunks and every warp walks its chunks in increasing order, so the earliest
    // unfinished chunk can always run: no deadlock as long as all warps are resident (cooperative
    // launch), and no claim atomics on the critical path.
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);; chunk += nwarps) {
        const uint32_t oi = chunk * GPW + sub;
        if ((uint64_t)chunk * GPW >= total) return;
        const uint32_t v = oi < total ? m.order[oi] : NO_NODE;
        const bool act = v != NO_NODE;
        uint64_t p0 = 0, p1 = 0;
        uint32_t lv = 0;
        Nb nb; nb.deg = 0; nb.x = nb.y = nb.z = nb.base = 0;
        if (act) { p0 = m.ptr[v]; p1 = m.ptr[v + 1]; lv = m.level[v]; nb = load_nb(m, v); }
        float bh = INFINITY;
        uint32_t bk = 0xFFFFFFFFu;
        if (nb.deg <= 3 && m.mask) {
            // ---- fast path: manifold degree, label bitmasks; 32-bit offsets inside the node ----
            const uint32_t n = (uint32_t)(p1 - p0);
            const float *__restrict__ costv = m.cost + p0;
            const uint16_t *__restrict__ viewv = m.view + p0;
            float *Hv = m.H + p0;
            // issue the loads of the first two label chunks now: they overlap the neighbour
            // classification and the flag waits below
            uint32_t lab0 = 0, lab1 = 0;
            float c0 = 0.0f, c1 = 0.0f;
            if (lane < n) { lab0 = (uint32_t)viewv[lane] + 1u; c0 = costv[lane]; }
            if (lane + G < n) { lab1 = (uint32_t)viewv[lane + G] + 1u; c1 = costv[lane + G]; }
            // classify the (at most three) neighbours once: 0 = skip (unseen / parent), 1 = child, 2 = fixed
            uint32_t kind[3] = {0, 0, 0}, xw[3] = {0, 0, 0}, wv[3] = {0, 0, 0};
            float hm[3] = {0.0f, 0.0f, 0.0f};
            const float *Hw[3] = {nullptr, nullptr, nullptr};
            const uint32_t *mw[3] = {nullptr, nullptr, nullptr};
            const uint16_t *mp[3] = {nullptr, nullptr, nullptr};
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                if ((uint32_t)i >= nb.deg) continue;
                const uint32_t w = i == 0 ? nb.x : (i == 1 ? nb.y : nb.z);
                const uint32_t x = m.labels[w];
                wv[i] = w; xw[i] = x;
                if (x == 0) continue;  // unseen faces carry no edges (view_selection.cpp:30,35)
                const uint32_t lw = local_pair(m, v, w) ? m.level[w] : LVL_DEAD;
                if (lw <= m.rounds) {
                    if (lw > lv) {
                        kind[i] = 1;
                        Hw[i] = m.H + m.ptr[w];
                        mw[i] = m.mask + (size_t)w * m.mask_words;
                        mp[i] = m.mpre + (size_t)w * m.mask_words;
                    }
                } else kind[i] = 2;
            }
            bool any_child = false;
#pragma unroll
            for (int i = 0; i < 3; ++i)
                if (kind[i] == 1) {  // wait for the child's messages
                    while (ld_relaxed(m.flag_up + wv[i]) != stamp) __nanosleep(20);
                    any_child = true;
                }
            if (any_child) fence_acq_rel();  // one acquire fence for all children of this node
#pragma unroll
            for (int i = 0; i < 3; ++i)
                if (kind[i] == 1) hm[i] = __ldcg(m.hminp1 + wv[i]);
            auto eval = [&](uint32_t lab, float h) -> float {
                const uint32_t word = lab >> 5, bit = lab & 31u, below = (1u << bit) - 1u;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    if (kind[i] == 1) {  // Potts message min(h_w(lab), hmin_w + 1)
                        float msg = hm[i];
                        const uint32_t bits = __ldg(mw[i] + word);
                        if ((bits >> bit) & 1u) {
                            const uint32_t j = (uint32_t)__ldg(mp[i] + word) + __popc(bits & below);
                            const float hw = __ldcg(Hw[i] + j);
                            if (hw < msg) msg = hw;
                        }
                        h = h + msg;
                    } else if (kind[i] == 2) {
                        h = h + (lab != xw[i] ? 1.0f : 0.0f);
                    }
                }
                return h;
            };
            if (lane < n) {
                const float h0 = eval(lab0, c0);
                float h1 = 0.0f;
                if (lane + G < n) h1 = eval(lab1, c1);
                Hv[lane] = h0;
                if (h0 < bh) { bh = h0; bk = lane; }
                if (lane + G < n) {
                    Hv[lane + G] = h1;
                    if (h1 < bh) { bh = h1; bk = lane + G; }
                }
            }
            for (uint32_t k = lane + 2 * G; k < n; k += G) {
                const float h = eval((uint32_t)viewv[k] + 1u, costv[k]);
                Hv[k] = h;
                if (h < bh) { bh = h; bk = k; }
            }
        } else {
            // generic degree (non-manifold edges): neighbour loop inside the label loop
            for (uint32_t i = 0; i < nb.deg; ++i) {
                const uint32_t w = nb_at(m, nb, i);
                if (m.labels[w] == 0 || !local_pair(m, v, w)) continue;
                const uint32_t lw = m.level[w];
                if (lw <= m.rounds && lw > lv)
                    while (ld_acquire(m.flag_up + w) != stamp) __nanosleep(32);
            }
            for (uint64_t k = p0 + lane; k < p1; k += G) {
                const uint32_t lab = (uint32_t)m.view[k] + 1u;
                float h = m.cost[k];
                for (uint32_t i = 0; i < nb.deg; ++i) {
                    const uint32_t w = nb_at(m, nb, i);
                    const uint32_t x = m.labels[w];
                    if (x == 0) continue;
                    const uint32_t lw = local_pair(m, v, w) ? m.level[w] : LVL_DEAD;
                    if (lw <= m.rounds) {
                        if (lw > lv) {
                            float msg = __ldcg(m.hminp1 + w);
                            long long j = find_label(m, w, lab);
                            if (j >= 0) { float hw = __ldcg(m.H + j); if (hw < msg) msg = hw; }
                            h = h + msg;
                        }
                    } else {
                        h = h + (lab != x ? 1.0f : 0.0f);
                    }
                }
                m.H[k] = h;
                if (h < bh) { bh = h; bk = (uint32_t)(k - p0); }
            }
        }
        __syncwarp();  // orders the H stores of all lanes before lane 0's fence + release below
        for (int s = G / 2; s; s >>= 1) {
            float oh = __shfl_xor_sync(0xffffffffu, bh, s);
            uint32_t ok = __shfl_xor_sync(0xffffffffu, bk, s);
            if (oh < bh || (oh == bh && ok < bk)) { bh = oh; bk = ok; }
        }
        if (act && lane == 0) {
            m.hminp1[v] = bh + 1.0f;
            m.amin[v] = bk;
            fence_acq_rel();  // cumulative: covers the H stores of the other lanes (ordered by __syncwarp)
            st_relaxed(m.flag_up + v, stamp);
        }
    }
}

// ---- top-down assignment, dataflow: one thread per node waits for its parent ---------------------------
__global__ void __launch_bounds__(256, 8) k_down(Mrf m)
{
    const uint32_t total = m.ctl[CTL_TOTAL];
    const uint32_t stamp = m.iter;
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);; chunk += nwarps) {
        if ((uint64_t)chunk * 32u >= total) return;
        // shallowest level first: walk the order array backwards (a warp never straddles two levels)
        const uint32_t oi = total - 1u - (chunk * 32u + (threadIdx.x & 31));
        const uint32_t v = m.order[oi];
        if (v == NO_NODE) continue;
        const uint32_t lv = m.level[v];
        uint32_t best = (uint32_t)m.view[m.ptr[v] + m.amin[v]] + 1u;
        if (lv > 0) {
            const Nb nb = load_nb(m, v);
            for (uint32_t i = 0; i < nb.deg; ++i) {
                const uint32_t w = nb_at(m, nb, i);
                if (local_pair(m, v, w) && m.level[w] < lv) {
                    while (ld_acquire(m.flag_dn + w) != stamp) __nanosleep(32);
                    const uint32_t xp = __ldcg(m.labels + w);
                    long long j = find_label(m, v, xp);
                    if (j >= 0 && m.H[j] <= m.hminp1[v]) best = xp;
                    break;
                }
            }
        }
        m.labels[v] = best;
        st_release(m.flag_dn + v, stamp);
    }
}

// 32.32 fixed-point energy of the owned nodes: unaries + edges counted by their lower endpoint
__global__ void __launch_bounds__(256) k_energy(Mrf m)
{
    unsigned long long e = 0;
    for (uint32_t v = m.nb + blockIdx.x * blockDim.x + threadIdx.x; v < m.ne; v += gridDim.x * blockDim.x) {
        uint32_t x = m.labels[v];
        if (x == 0) { e += 1ull << 32; continue; }
        long long j = find_label(m, v, x);
        if (j >= 0) e += (unsigned long long)(long long)((double)m.cost[j] * 4294967296.0);
        const Nb nb = load_nb(m, v);
        for (uint32_t i = 0; i < nb.deg; ++i) {
            uint32_t w = nb_at(m, nb, i);
            uint32_t xw = m.labels[w];
            if (w > v && xw != 0 && xw != x) e += 1ull << 32;
        }
    }
    for (int s = 16; s; s >>= 1) e += __shfl_xor_sync(0xffffffffu, e, s);
    if ((threadIdx.x & 31) == 0 && e) atomicAdd(m.energy, e);
}

Mrf make_mrf(b2tex_ctx *c, uint32_t iter)
{
    Mrf m;
    m.F = c->F; m.nb = c->face_begin; m.ne = c->face_end;
    m.adj_ptr = c->adj_ptr.p; m.adj_idx = c->adj_idx.p;
    m.adj4 = c->mrf_adj4.p;
    m.ptr = c->dc_ptr.p; m.view = c->dc_view.p; m.cost = c->dc_cost.p;
    m.H = c->mrf_H.p; m.hminp1 = c->mrf_hminp1.p; m.amin = c->mrf_amin.p; m.level = c->mrf_level.p;
    m.labels = c->labels.p; m.order = c->mrf_order.p; m.ctl = c->mrf_ctl.p;
    m.flag_up = c->mrf_flags.p; m.flag_dn = c->mrf_flags.p + c->F;
    m.queue = c->mrf_queue.p; m.qstamp = c->mrf_queue.p + 2 * (size_t)c->F;
    m.sort_key_in = c->mrf_sort.p; m.sort_key_out = c->mrf_sort.p + c->F;
    m.sort_val_in = c->mrf_sort.p + 2 * (size_t)c->F; m.sort_val_out = c->mrf_sort.p + 3 * (size_t)c->F;
    m.mask = c->mrf_mask_words ? c->mrf_mask.p : nullptr;
    m.mpre = c->mrf_mask_words ? c->mrf_mpre.p : nullptr;
    m.mask_words = c->mrf_mask_words;
    m.energy = c->mrf_energy.p;
    const b2tex_mrf_params &p = c->mrf_params;
    uint32_t P = p.num_parts ? p.num_parts : 1;
    m.part_size = (c->F + P - 1) / P; if (!m.part_size) m.part_size = 1;
    m.rounds = p.rounds;
    if (p.root_div == 0) m.rdiv = 0;
    else { uint32_t cap = c->F / 8u; if (cap < 1u) cap = 1u; m.rdiv = p.root_div < cap ? p.root_div : cap; }
    m.seed = p.seed;
    m.iter = iter;
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

int launch_forest(b2tex_ctx *c, Mrf &m, int do_bucket)
{
    cudaStream_t s = c->stream;
    size_t sh = (size_t)(m.rounds + 1) * sizeof(uint32_t);
    int grid = 0;
    // few fat blocks: the cost of grid.sync() grows with the number of blocks
    B2_TRY(coop_grid(c, k_forest, sh, &grid, 1024));
    if (grid > c->num_sms) grid = c->num_sms;  // one fat block per SM: cheapest grid.sync()
    if (const char *e = getenv("B2TEX_FOREST_BLOCKS_PER_SM")) grid = c->num_sms * std::max(1, atoi(e));
    uint32_t n = m.ne - m.nb;
    int need = (int)((n + 1023) / 1024);
    if (grid > need) grid = need > 0 ? need : 1;
    B2_CUDA(cudaMemsetAsync(m.ctl, 0, CTL_WORDS * sizeof(uint32_t), s));
    if (do_bucket) B2_CUDA(cudaMemsetAsync(m.order, 0xFF, c->mrf_order.n * sizeof(uint32_t), s));
    void *args[] = {&m, &do_bucket};
    B2_CUDA(cudaLaunchCooperativeKernel((void *)k_forest, dim3(grid), dim3(1024), args, sh, s));
    if (do_bucket) {
        int bits = 1;
        while ((1u << bits) < m.rounds + 2u) ++bits;
        size_t bytes = 0;
        B2_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, m.sort_key_in, m.sort_key_out, m.sort_val_in,
                                                m.sort_val_out, (int)n, 0, bits, s));
        B2_TRY(c->cub_tmp.alloc(bytes));
        B2_CUDA(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, bytes, m.sort_key_in, m.sort_key_out, m.sort_val_in,
                                                m.sort_val_out, (int)n, 0, bits, s));
        k_scatter_order<<<(n + 255) / 256, 256, 0, s>>>(m);
        B2_KERNEL_CHECK();
    }
    return B2TEX_OK;
}

template <int G>
int launch_up(b2tex_ctx *c, Mrf &m)
{
    int grid = 0;
    B2_TRY(coop_grid(c, k_up<G>, 0, &grid));
    void *args[] = {&m};
    // cooperative launch only to guarantee co-residency of the spinning warps (no grid.sync inside)
    B2_CUDA(cudaLaunchCooperativeKernel((void *)k_up<G>, dim3(grid), dim3(256), args, 0, c->stream));
    return B2TEX_OK;
}

int enqueue_iteration(b2tex_ctx *c, Mrf &m)
{
    cudaStream_t s = c->stream;
    if (m.ne <= m.nb) return B2TEX_OK;
    {
        ScopedTimer tf(c, "mrf.k_forest+sort", 20.0 * (double)(m.ne - m.nb));
        B2_TRY(launch_forest(c, m, 1));
    }
    std::unique_ptr<ScopedTimer> tu(new ScopedTimer(c, "mrf.k_up", 14.0 * (double)c->nnz));
    switch (c->mrf_group) {
        case 4: B2_TRY(launch_up<4>(c, m)); break;
        case 8: B2_TRY(launch_up<8>(c, m)); break;
        case 16: B2_TRY(launch_up<16>(c, m)); break;
        default: B2_TRY(launch_up<32>(c, m)); break;
    }
    tu.reset();
    static const bool repeat_up = getenv("B2TEX_MRF_REPEAT_UP") != nullptr;
    if (repeat_up) {  // experiment: second sweep finds every flag already set (no dataflow waits)
        ScopedTimer tu2(c, "mrf.k_up(again)");
        switch (c->mrf_group) {
            case 4: B2_TRY(launch_up<4>(c, m)); break;
            case 8: B2_TRY(launch_up<8>(c, m)); break;
            case 16: B2_TRY(launch_up<16>(c, m)); break;
            default: B2_TRY(launch_up<32>(c, m)); break;
        }
    }
    {
        ScopedTimer td(c, "mrf.k_down");
        int grid = 0;
        B2_TRY(coop_grid(c, k_down, 0, &grid));
        void *args[] = {&m};
        B2_CUDA(cudaLaunchCooperativeKernel((void *)k_down, dim3(grid), dim3(256), args, 0, s));
    }
    ScopedTimer te(c, "mrf.k_energy");
    B2_CUDA(cudaMemsetAsync(m.energy, 0, sizeof(unsigned long long), s));
    k_energy<<<std::max(1, c->num_sms * 8), 256, 0, s>>>(m);
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
    B2_TRY(c->mrf_order.alloc(F + 32 * ((size_t)p->rounds + 2)));
    B2_TRY(c->mrf_sort.alloc(4 * F));    // sort keys/values in/out
    B2_TRY(c->mrf_queue.alloc(3 * F));   // frontier lists [2][F] | qstamp [F]
    B2_TRY(c->mrf_queue.zero(c->stream));
    B2_TRY(c->mrf_flags.alloc(2 * F));  // flag_up | flag_dn (iteration stamps)
    B2_TRY(c->mrf_flags.zero(c->stream));
    B2_TRY(c->mrf_ctl.alloc(CTL_WORDS));
    B2_TRY(c->mrf_energy.alloc(4));
    B2_TRY(c->mrf_adj4.alloc(F));
    if (F) k_build_adj4<<<(unsigned)((F + 255) / 256), 256, 0, c->stream>>>((uint32_t)F, c->adj_ptr.p, c->adj_idx.p, c->mrf_adj4.p);
    if (!c->have_labels || c->labels.n != F) { B2_TRY(c->labels.alloc(F)); B2_TRY(c->labels.zero(c->stream)); }
    uint32_t nodes = c->face_end - c->face_begin;
    double rho = nodes ? (double)c->nnz / nodes : 0.0;
    // lanes per node: fewer lanes = more nodes in flight per SM (the sweep is latency bound)
    c->mrf_group = rho >= 56 ? 32 : rho >= 12 ? 16 : rho >= 6 ? 8 : 4;
    if (const char *g = getenv("B2TEX_MRF_GROUP")) {
        int v = atoi(g);
        if (v == 4 || v == 8 || v == 16 || v == 32) c->mrf_group = v;
    }
    // label bitmasks: labels are view+1 <= K
    uint32_t words = (c->K + 1 + 31) / 32;
    static const bool no_masks = getenv("B2TEX_NO_MASKS") != nullptr;
    c->mrf_mask_words = (c->K == 0 || words > (uint32_t)MAX_MASK_WORDS || no_masks) ? 0 : words;
    if (c->mrf_mask_words) {
        B2_TRY(c->mrf_mask.alloc(F * c->mrf_mask_words));
        B2_TRY(c->mrf_mpre.alloc(F * c->mrf_mask_words));
    }
    return B2TEX_OK;
}

int read_energy(b2tex_ctx *c, const Mrf &m, int64_t *efix)
{
    unsigned long long e = 0;
    B2_CUDA(cudaMemcpyAsync(&e, m.energy, sizeof(e), cudaMemcpyDeviceToHost, c->stream));
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
        if (c->mrf_mask_words)
            k_build_masks<<<(m.ne - m.nb + 255) / 256, 256, 0, s>>>(m, c->mrf_mask.p, c->mrf_mpre.p);
        B2_KERNEL_CHECK();
    }
    c->have_labels = true;
    c->mrf_ready = true;
    B2_CUDA(cudaMemsetAsync(m.energy, 0, sizeof(unsigned long long), s));
    if (m.ne > m.nb) k_energy<<<grid, 256, 0, s>>>(m);
    B2_KERNEL_CHECK();
    return read_energy(c, m, efix);
}

int mrf_iterate(b2tex_ctx *c, uint32_t t, int64_t *efix)
{
    if (!c->mrf_ready) { set_error("mrf_iterate before mrf_init"); return B2TEX_ERR_ARG; }
    if (t == 0) { set_error("mrf_iterate: iterations are numbered from 1"); return B2TEX_ERR_ARG; }
    Mrf m = make_mrf(c, t);
    {
        B2_TRY(enqueue_iteration(c, m));
    }
    return read_energy(c, m, efix);
}

// fixed-point energy of the owned nodes with the labels currently in the context (a sharded run calls
// this after the boundary-label exchange, so that cut edges see the neighbours' NEW labels)
int mrf_energy_only(b2tex_ctx *c, int64_t *efix)
{
    if (!c->mrf_ready) { set_error("mrf_energy before mrf_init"); return B2TEX_ERR_ARG; }
    Mrf m = make_mrf(c, 1);
    B2_CUDA(cudaMemsetAsync(m.energy, 0, sizeof(unsigned long long), c->stream));
    if (m.ne > m.nb) k_energy<<<std::max(1, c->num_sms * 8), 256, 0, c->stream>>>(m);
    B2_KERNEL_CHECK();
    return read_energy(c, m, efix);
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
