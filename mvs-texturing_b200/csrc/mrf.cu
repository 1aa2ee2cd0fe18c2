// mrf.cu -- K5-K7: pairwise-Potts MRF view selection on the device.
//
// Replaces tex::view_selection's call into mapMAP (libs/tex/view_selection.cpp:84-118); the model
// (label sets, unaries, Potts edges between seen faces) follows view_selection.cpp:26-82.
// Solver = block coordinate descent over induced forests with exact min-sum DP, the algorithm
// defined in oracle/mrf.c; this file reproduces it bit for bit:
//   * forest sampling is order independent (hash priorities, level-synchronous rounds)
//   * messages are summed in adjacency order in fp32 (no FMA: additions and minima only)
//   * the energy used for termination is 32.32 fixed point, summed with integer atomics.
// Work mapping: a group of G lanes (G = 4..32, chosen from the mean label count) owns one node and
// strides over its sorted label list; min/argmin by warp shuffles; one launch per forest level,
// the whole iteration captured in a CUDA graph.
#include "common.cuh"

namespace b2 {

namespace {

constexpr uint32_t LVL_NONE = 0xFFFFFFFFu;
constexpr uint32_t LVL_DEAD = 0xFFFFFFFEu;

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
    const uint64_t *ptr;
    const uint16_t *view;
    const float *cost;
    float *H, *hminp1;
    uint32_t *amin, *level, *labels, *order, *lvl_ptr, *cursor;
    const uint32_t *iter;        // device scalar: current iteration
    uint32_t *max_prio;          // device scalar for single-root mode
    unsigned long long *energy;
    uint32_t part_size, rounds, rdiv, seed;
};

__device__ __forceinline__ bool same_part(const Mrf &m, uint32_t a, uint32_t b)
{
    return a / m.part_size == b / m.part_size;
}
__device__ __forceinline__ bool owned(const Mrf &m, uint32_t v) { return v >= m.nb && v < m.ne; }

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

__global__ void k_max_prio(Mrf m)
{
    uint32_t v = m.nb + blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t p = 0;
    bool have = false;
    if (v < m.ne && m.labels[v] != 0) { p = prio(v, iter_seed(m.seed, *m.iter)); have = true; }
    // encode "have" so that prio 0 still wins over "none": store prio as 33-bit? use p|1 trick
    // (prio is a bijection; collisions of p|1 only pair two nodes -> both may become roots only if
    // non adjacent; the oracle uses the exact maximum, so keep exactness with a 64-bit key)
    unsigned long long key = have ? (((unsigned long long)p << 1) | 1ull) : 0ull;
    for (int s = 16; s; s >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xffffffffu, key, s);
        key = o > key ? o : key;
    }
    if ((threadIdx.x & 31) == 0 && key) atomicMax(reinterpret_cast<unsigned long long *>(m.max_prio), key);
}

__global__ void __launch_bounds__(256) k_forest_init(Mrf m)
{
    uint32_t v = m.nb + blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= m.ne) return;
    const uint32_t seed_t = iter_seed(m.seed, *m.iter);
    if (m.labels[v] == 0) { m.level[v] = LVL_DEAD; return; }
    const uint32_t pv = prio(v, seed_t);
    bool eligible = true, is_root;
    if (m.rdiv) is_root = root_cand(v, seed_t, m.rdiv);
    else {
        unsigned long long key = *reinterpret_cast<unsigned long long *>(m.max_prio);
        is_root = key == ((((unsigned long long)pv) << 1) | 1ull);
    }
    for (uint32_t a = m.adj_ptr[v]; a < m.adj_ptr[v + 1]; ++a) {
        uint32_t w = m.adj_idx[a];
        if (m.labels[w] == 0) continue;
        if (!(owned(m, w) && same_part(m, v, w))) { if (prio(w, seed_t) > pv) eligible = false; continue; }
        if (m.rdiv && is_root && root_cand(w, seed_t, m.rdiv) && prio(w, seed_t) > pv) is_root = false;
    }
    m.level[v] = !eligible ? LVL_DEAD : (is_root ? 0u : LVL_NONE);
}

__device__ __forceinline__ uint32_t count_in_forest(const Mrf &m, uint32_t v, uint32_t r)
{
    uint32_t c = 0;
    for (uint32_t a = m.adj_ptr[v]; a < m.adj_ptr[v + 1]; ++a) {
        uint32_t w = m.adj_idx[a];
        if (owned(m, w) && same_part(m, v, w) && ((volatile uint32_t *)m.level)[w] < r) ++c;
    }
    return c;
}

__global__ void __launch_bounds__(256) k_forest_round(Mrf m, uint32_t r)
{
    uint32_t v = m.nb + blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= m.ne) return;
    if (m.level[v] != LVL_NONE) return;
    uint32_t c = count_in_forest(m, v, r);
    if (c >= 2) { m.level[v] = LVL_DEAD; return; }
    if (c != 1) return;
    const uint32_t seed_t = iter_seed(m.seed, *m.iter);
    const uint32_t pv = prio(v, seed_t);
    for (uint32_t a = m.adj_ptr[v]; a < m.adj_ptr[v + 1]; ++a) {
        uint32_t w = m.adj_idx[a];
        if (!(owned(m, w) && same_part(m, v, w))) continue;
        uint32_t lw = ((volatile uint32_t *)m.level)[w];
        if (!(lw == LVL_NONE || lw == r)) continue;
        if (prio(w, seed_t) < pv) continue;
        if (count_in_forest(m, w, r) == 1) return;  // a stronger adjacent candidate: wait
    }
    m.level[v] = r;
}

// ---- bucket nodes by level -------------------------------------------------------------------
constexpr int MAX_LEVELS = 1024;  // rounds + 1 must fit for the shared-memory bucketing

__global__ void __launch_bounds__(256) k_bucket_count(Mrf m)
{
    extern __shared__ uint32_t sc[];
    for (uint32_t i = threadIdx.x; i <= m.rounds; i += blockDim.x) sc[i] = 0;
    __syncthreads();
    for (uint32_t v = m.nb + blockIdx.x * blockDim.x + threadIdx.x; v < m.ne; v += gridDim.x * blockDim.x) {
        uint32_t l = m.level[v];
        if (l <= m.rounds) atomicAdd(&sc[l], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i <= m.rounds; i += blockDim.x)
        if (sc[i]) atomicAdd(&m.cursor[i], sc[i]);
}

__global__ void k_bucket_scan(Mrf m)
{
    // single thread: rounds+1 entries
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        uint32_t acc = 0;
        for (uint32_t i = 0; i <= m.rounds; ++i) {
            uint32_t c = m.cursor[i];
            m.lvl_ptr[i] = acc;
            m.cursor[i] = acc;
            acc += c;
        }
        m.lvl_ptr[m.rounds + 1] = acc;
    }
}

__global__ void __launch_bounds__(256) k_bucket_fill(Mrf m)
{
    extern __shared__ uint32_t sc[];  // [rounds+1] counts, then [rounds+1] bases
    uint32_t *cnt = sc, *base = sc + (m.rounds + 1);
    for (uint32_t start = m.nb + blockIdx.x * blockDim.x; start < m.ne; start += gridDim.x * blockDim.x) {
        for (uint32_t i = threadIdx.x; i <= m.rounds; i += blockDim.x) cnt[i] = 0;
        __syncthreads();
        uint32_t v = start + threadIdx.x;
        uint32_t l = v < m.ne ? m.level[v] : LVL_DEAD;
        uint32_t rank = 0;
        if (l <= m.rounds) rank = atomicAdd(&cnt[l], 1u);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i <= m.rounds; i += blockDim.x)
            if (cnt[i]) base[i] = atomicAdd(&m.cursor[i], cnt[i]);
        __syncthreads();
        if (l <= m.rounds) m.order[base[l] + rank] = v;
        __syncthreads();
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

// bottom-up min-sum messages for forest level r
template <int G>
__global__ void __launch_bounds__(256) k_up(Mrf m, uint32_t r)
{
    const uint32_t beg = m.lvl_ptr[r], end = m.lvl_ptr[r + 1];
    const uint32_t lane = threadIdx.x & (G - 1);
    const uint32_t gpb = blockDim.x / G;
    for (uint32_t base = beg + blockIdx.x * gpb; base < end; base += gridDim.x * gpb) {
        uint32_t oi = base + threadIdx.x / G;
        bool act = oi < end;
        uint32_t v = act ? m.order[oi] : 0;
        uint64_t p0 = act ? m.ptr[v] : 0, p1 = act ? m.ptr[v + 1] : 0;
        uint32_t a0 = act ? m.adj_ptr[v] : 0, a1 = act ? m.adj_ptr[v + 1] : 0;
        float bh = INFINITY;
        uint32_t bk = 0xFFFFFFFFu;
        for (uint64_t k = p0 + lane; k < p1; k += G) {
            const uint32_t lab = (uint32_t)m.view[k] + 1u;
            float h = m.cost[k];
            for (uint32_t a = a0; a < a1; ++a) {
                const uint32_t w = m.adj_idx[a];
                const uint32_t xw = m.labels[w];
                if (xw == 0) continue;  // unseen faces carry no edges (view_selection.cpp:30,35)
                const uint32_t lw = (owned(m, w) && same_part(m, v, w)) ? m.level[w] : LVL_DEAD;
                if (lw <= m.rounds) {
                    if (lw > r) {  // child: Potts message min(h_w(lab), hmin_w + 1)
                        float msg = m.hminp1[w];
                        long long j = find_label(m, w, lab);
                        if (j >= 0) { float hw = m.H[j]; if (hw < msg) msg = hw; }
                        h = h + msg;
                    }              // parent: skipped
                } else {
                    h = h + (lab != xw ? 1.0f : 0.0f);
                }
            }
            m.H[k] = h;
            if (h < bh) { bh = h; bk = (uint32_t)(k - p0); }
        }
        __syncwarp();
        for (int s = G / 2; s; s >>= 1) {
            float oh = __shfl_xor_sync(0xffffffffu, bh, s);
            uint32_t ok = __shfl_xor_sync(0xffffffffu, bk, s);
            if (oh < bh || (oh == bh && ok < bk)) { bh = oh; bk = ok; }
        }
        if (act && lane == 0) { m.hminp1[v] = bh + 1.0f; m.amin[v] = bk; }
    }
}

// top-down assignment for forest level r
__global__ void __launch_bounds__(256) k_down(Mrf m, uint32_t r)
{
    const uint32_t beg = m.lvl_ptr[r], end = m.lvl_ptr[r + 1];
    for (uint32_t oi = beg + blockIdx.x * blockDim.x + threadIdx.x; oi < end; oi += gridDim.x * blockDim.x) {
        uint32_t v = m.order[oi];
        uint32_t best = (uint32_t)m.view[m.ptr[v] + m.amin[v]] + 1u;
        if (r > 0) {
            for (uint32_t a = m.adj_ptr[v]; a < m.adj_ptr[v + 1]; ++a) {
                uint32_t w = m.adj_idx[a];
                if (owned(m, w) && same_part(m, v, w) && m.level[w] < r) {
                    uint32_t xp = m.labels[w];
                    long long j = find_label(m, v, xp);
                    if (j >= 0 && m.H[j] <= m.hminp1[v]) best = xp;
                    break;
                }
            }
        }
        m.labels[v] = best;
    }
}

// 32.32 fixed-point energy of the owned nodes: unaries + edges (i<j, counted by the lower id if both
// owned, else by the owner of the lower... every edge is counted by its lower endpoint's owner)
__global__ void __launch_bounds__(256) k_energy(Mrf m)
{
    unsigned long long e = 0;
    for (uint32_t v = m.nb + blockIdx.x * blockDim.x + threadIdx.x; v < m.ne; v += gridDim.x * blockDim.x) {
        uint32_t x = m.labels[v];
        if (x == 0) { e += 1ull << 32; continue; }
        long long j = find_label(m, v, x);
        if (j >= 0) e += (unsigned long long)(long long)((double)m.cost[j] * 4294967296.0);
        for (uint32_t a = m.adj_ptr[v]; a < m.adj_ptr[v + 1]; ++a) {
            uint32_t w = m.adj_idx[a];
            uint32_t xw = m.labels[w];
            if (w > v && xw != 0 && xw != x) e += 1ull << 32;
        }
    }
    for (int s = 16; s; s >>= 1) e += __shfl_xor_sync(0xffffffffu, e, s);
    if ((threadIdx.x & 31) == 0 && e) atomicAdd(m.energy, e);
}

Mrf make_mrf(b2tex_ctx *c)
{
    Mrf m;
    m.F = c->F; m.nb = c->face_begin; m.ne = c->face_end;
    m.adj_ptr = c->adj_ptr.p; m.adj_idx = c->adj_idx.p;
    m.ptr = c->dc_ptr.p; m.view = c->dc_view.p; m.cost = c->dc_cost.p;
    m.H = c->mrf_H.p; m.hminp1 = c->mrf_hminp1.p; m.amin = c->mrf_amin.p; m.level = c->mrf_level.p;
    m.labels = c->labels.p; m.order = c->mrf_order.p; m.lvl_ptr = c->mrf_lvlptr.p; m.cursor = c->mrf_cursor.p;
    m.iter = c->mrf_cursor.p + MAX_LEVELS + 8;
    m.max_prio = c->mrf_cursor.p + MAX_LEVELS + 16;
    m.energy = c->mrf_energy.p;
    const b2tex_mrf_params &p = c->mrf_params;
    uint32_t P = p.num_parts ? p.num_parts : 1;
    m.part_size = (c->F + P - 1) / P; if (!m.part_size) m.part_size = 1;
    m.rounds = p.rounds;
    if (p.root_div == 0) m.rdiv = 0;
    else { uint32_t cap = c->F / 8u; if (cap < 1u) cap = 1u; m.rdiv = p.root_div < cap ? p.root_div : cap; }
    m.seed = p.seed;
    return m;
}

template <int G>
void launch_up(const Mrf &m, uint32_t r, int grid, cudaStream_t s) { k_up<G><<<grid, 256, 0, s>>>(m, r); }

int enqueue_forest(b2tex_ctx *c, const Mrf &m)
{
    cudaStream_t s = c->stream;
    const uint32_t n = m.ne - m.nb;
    const uint32_t nblocks = (n + 255) / 256;
    if (!n) return B2TEX_OK;
    if (m.rdiv == 0) {
        B2_CUDA(cudaMemsetAsync(m.max_prio, 0, 8, s));
        k_max_prio<<<nblocks, 256, 0, s>>>(m);
    }
    k_forest_init<<<nblocks, 256, 0, s>>>(m);
    for (uint32_t r = 1; r <= m.rounds; ++r) k_forest_round<<<nblocks, 256, 0, s>>>(m, r);
    B2_KERNEL_CHECK();
    return B2TEX_OK;
}

int enqueue_iteration(b2tex_ctx *c, const Mrf &m)
{
    cudaStream_t s = c->stream;
    const uint32_t n = m.ne - m.nb;
    if (!n) return B2TEX_OK;
    B2_TRY(enqueue_forest(c, m));
    const int grid = std::max(1, c->num_sms * 8);
    B2_CUDA(cudaMemsetAsync(m.cursor, 0, (m.rounds + 2) * sizeof(uint32_t), s));
    size_t sh = (m.rounds + 1) * sizeof(uint32_t);
    k_bucket_count<<<grid, 256, sh, s>>>(m);
    k_bucket_scan<<<1, 32, 0, s>>>(m);
    k_bucket_fill<<<grid, 256, 2 * sh, s>>>(m);
    for (int r = (int)m.rounds; r >= 0; --r) {
        switch (c->mrf_group) {
            case 4: launch_up<4>(m, r, grid, s); break;
            case 8: launch_up<8>(m, r, grid, s); break;
            case 16: launch_up<16>(m, r, grid, s); break;
            default: launch_up<32>(m, r, grid, s); break;
        }
    }
    for (uint32_t r = 0; r <= m.rounds; ++r) k_down<<<grid, 256, 0, s>>>(m, r);
    B2_CUDA(cudaMemsetAsync(m.energy, 0, sizeof(unsigned long long), s));
    k_energy<<<grid, 256, 0, s>>>(m);
    B2_KERNEL_CHECK();
    return B2TEX_OK;
}

int set_iter(b2tex_ctx *c, const Mrf &m, uint32_t t)
{
    B2_CUDA(cudaMemcpyAsync((void *)m.iter, &t, sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
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
    B2_TRY(c->mrf_order.alloc(F));
    B2_TRY(c->mrf_lvlptr.alloc(MAX_LEVELS + 8));
    B2_TRY(c->mrf_cursor.alloc(MAX_LEVELS + 32));
    B2_TRY(c->mrf_energy.alloc(4));
    if (!c->have_labels || c->labels.n != F) { B2_TRY(c->labels.alloc(F)); B2_TRY(c->labels.zero(c->stream)); }
    uint32_t nodes = c->face_end - c->face_begin;
    double rho = nodes ? (double)c->nnz / nodes : 0.0;
    c->mrf_group = rho >= 24 ? 32 : rho >= 12 ? 16 : rho >= 6 ? 8 : 4;
    if (c->mrf_graph_exec) { cudaGraphExecDestroy((cudaGraphExec_t)c->mrf_graph_exec); c->mrf_graph_exec = nullptr; }
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
    Mrf m = make_mrf(c);
    cudaStream_t s = c->stream;
    const int grid = std::max(1, c->num_sms * 8);
    if (m.ne > m.nb) {
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
    B2_CUDA(cudaMemsetAsync(m.energy, 0, sizeof(unsigned long long), s));
    if (m.ne > m.nb) k_energy<<<grid, 256, 0, s>>>(m);
    B2_KERNEL_CHECK();
    return read_energy(c, m, efix);
}

int mrf_iterate(b2tex_ctx *c, uint32_t t, int64_t *efix)
{
    if (!c->mrf_ready) { set_error("mrf_iterate before mrf_init"); return B2TEX_ERR_ARG; }
    Mrf m = make_mrf(c);
    cudaStream_t s = c->stream;
    B2_TRY(set_iter(c, m, t));
    ScopedTimer tm(c, "mrf_iteration", 14.0 * (double)c->nnz + 20.0 * (double)(c->face_end - c->face_begin));
    static const bool no_graph = getenv("B2TEX_NO_GRAPH") != nullptr;
    if (no_graph) {
        B2_TRY(enqueue_iteration(c, m));
    } else {
        if (!c->mrf_graph_exec) {
            cudaGraph_t graph;
            B2_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            int rc = enqueue_iteration(c, m);
            cudaError_t e = cudaStreamEndCapture(s, &graph);
            if (rc != B2TEX_OK) return rc;
            B2_CUDA(e);
            cudaGraphExec_t exec;
            B2_CUDA(cudaGraphInstantiate(&exec, graph, 0));
            cudaGraphDestroy(graph);
            c->mrf_graph_exec = exec;
        }
        B2_CUDA(cudaGraphLaunch((cudaGraphExec_t)c->mrf_graph_exec, s));
    }
    return read_energy(c, m, efix);
}

int mrf_sample_only(b2tex_ctx *c, const b2tex_mrf_params *p, uint32_t t, uint32_t *level_host)
{
    if (!c->mrf_ready) { int64_t e; B2_TRY(mrf_init(c, p, &e)); }
    c->mrf_params = *p;
    Mrf m = make_mrf(c);
    B2_TRY(set_iter(c, m, t));
    B2_TRY(enqueue_forest(c, m));
    B2_TRY(c->mrf_level.download(level_host, c->F, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return B2TEX_OK;
}

}  // namespace b2
