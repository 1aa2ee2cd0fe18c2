/* oracle/mrf.c -- TEST INFRASTRUCTURE (see oracle.h).
 *
 * MRF model: libs/tex/view_selection.cpp:26-82,120-131 (fully specified in-tree):
 *   node i = face i; labels = candidate view_id+1 ascending (:53-56), unseen faces {0} with
 *   unary 1.0 and no edges (:30,35,50-51,69-70); Potts(1.0) on edges between seen faces (:39,64).
 *   E(x) = sum_seen D_i(x_i) + #{(i,j): x_i != x_j} + #unseen.
 *
 * Solver: the reference calls mapMAP (dthuerck/mapmap_cpu@fa526e0, absent).  This file restates
 * the published core of mapMAP [UPSTREAM-RECALL, Thuerck et al. HPG'16 / Chen & Koltun CVPR'14]:
 * block coordinate descent where each block is a node subset that INDUCES A FOREST, solved
 * exactly by min-sum dynamic programming with every other neighbour's label held fixed
 * (use_acyclic + relax_acyclic_maximal, view_selection.cpp:106,110), seeded deterministically
 * (:114-115) and stopped by StopWhenReturnsDiminish(5, 0.01) (:84).  It is a stand-in for
 * mapMAP, not mapMAP: label-level parity with stock texrecon is NOT claimed; the CUDA path is
 * required to reproduce THIS solver bit for bit (same forest, same fp32 operation order).
 *
 * Forest sampling (level-synchronous, order independent, hence identical on CPU and GPU):
 *   prio_t(v) = mix32(v ^ seed_t) is a bijection -> strict total order per iteration t.
 *   round 0 : root candidates (mix32(prio ^ C) % root_div == 0) that beat adjacent candidates.
 *   round r : an undecided node with >=2 neighbours already in S is excluded for good;
 *             with exactly 1 it is a candidate and joins iff it beats every adjacent candidate.
 *   Joining nodes are pairwise non-adjacent leaves attached by exactly one edge => S stays an
 *   induced forest; parent(v) = the unique neighbour with a smaller level.
 * Logical partitions (multi-GPU emulation): neighbours in another partition are always treated
 * as fixed, and a node with remote neighbours is eligible only if it beats all of them.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* fork-join per forest level is latency bound: more than 16 threads only add wake-up cost */
static int g_mrf_threads = 1;
#define LVL_NONE 0xFFFFFFFFu
#define LVL_DEAD 0xFFFFFFFEu

static inline uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
static inline uint32_t iter_seed(uint32_t seed, uint32_t t) { return mix32(seed + 0x9E3779B9u * (t + 1u)); }
static inline uint32_t prio(uint32_t v, uint32_t seed_t) { return mix32(v ^ seed_t); }
static inline int root_cand(uint32_t v, uint32_t seed_t, uint32_t root_div)
{
    return mix32(prio(v, seed_t) ^ 0x68E31DA4u) % root_div == 0;
}
/* effective root spacing: tiny graphs still get roots; 0 = "single root" mode (the seen node
 * with the largest priority), which makes one sweep exact on tree-shaped inputs */
static inline uint32_t root_div_eff(uint32_t root_div, uint32_t F)
{
    if (root_div == 0) return 0;
    uint32_t cap = F / 8u; if (cap < 1u) cap = 1u;
    return root_div < cap ? root_div : cap;
}

typedef struct {
    uint32_t F;
    const uint32_t *adj_ptr, *adj_idx;
    const uint64_t *ptr;
    const uint16_t *view;
    const float *cost;
    uint32_t part_size;
} mrf_t;

static inline int seen(const mrf_t *m, uint32_t v) { return m->ptr[v + 1] > m->ptr[v]; }
static inline int same_part(const mrf_t *m, uint32_t a, uint32_t b) { return a / m->part_size == b / m->part_size; }

int64_t orc_mrf_energy_fixed(uint32_t F, const uint32_t *adj_ptr, const uint32_t *adj_idx,
                             const uint64_t *ptr, const uint16_t *view, const float *cost,
                             const uint32_t *labels)
{
    int64_t e = 0;
    for (uint32_t i = 0; i < F; ++i) {
        if (ptr[i + 1] == ptr[i]) { e += (int64_t)1 << 32; continue; }
        for (uint64_t k = ptr[i]; k < ptr[i + 1]; ++k)
            if ((uint32_t)view[k] + 1u == labels[i]) { e += (int64_t)((double)cost[k] * 4294967296.0); break; }
        for (uint32_t a = adj_ptr[i]; a < adj_ptr[i + 1]; ++a) {
            uint32_t j = adj_idx[a];
            if (j > i && ptr[j + 1] > ptr[j] && labels[i] != labels[j]) e += (int64_t)1 << 32;
        }
    }
    return e;
}

double orc_mrf_energy(uint32_t F, const uint32_t *adj_ptr, const uint32_t *adj_idx,
                      const uint64_t *ptr, const uint16_t *view, const float *cost,
                      const uint32_t *labels)
{
    double e = 0.0;
    for (uint32_t i = 0; i < F; ++i) {
        if (ptr[i + 1] == ptr[i]) { e += 1.0; continue; }
        int found = 0;
        for (uint64_t k = ptr[i]; k < ptr[i + 1]; ++k)
            if ((uint32_t)view[k] + 1u == labels[i]) { e += (double)cost[k]; found = 1; break; }
        if (!found) return INFINITY; /* label outside the node's label set */
        for (uint32_t a = adj_ptr[i]; a < adj_ptr[i + 1]; ++a) {
            uint32_t j = adj_idx[a];
            if (j > i && ptr[j + 1] > ptr[j] && labels[i] != labels[j]) e += 1.0;
        }
    }
    return e;
}

static void sample_forest(const mrf_t *m, const orc_mrf_params *pr, uint32_t t, uint32_t *level,
                          uint32_t *max_level_out)
{
    uint32_t F = m->F, seed_t = iter_seed(pr->seed, t);
    uint32_t rdiv = root_div_eff(pr->root_div, F);
    uint32_t best_prio = 0; int have_best = 0;
    if (rdiv == 0)
        for (uint32_t v = 0; v < F; ++v)
            if (seen(m, v) && (!have_best || prio(v, seed_t) > best_prio)) { best_prio = prio(v, seed_t); have_best = 1; }
    /* round 0: eligibility + roots */
    #pragma omp parallel for schedule(static) num_threads(g_mrf_threads) if (F > 8192)
    for (int64_t vv = 0; vv < (int64_t)F; ++vv) {
        uint32_t v = (uint32_t)vv;
        level[v] = LVL_NONE;
        if (!seen(m, v)) { level[v] = LVL_DEAD; continue; }
        uint32_t pv = prio(v, seed_t);
        int eligible = 1, is_root = rdiv ? root_cand(v, seed_t, rdiv) : (pv == best_prio);
        for (uint32_t a = m->adj_ptr[v]; a < m->adj_ptr[v + 1]; ++a) {
            uint32_t w = m->adj_idx[a];
            if (!seen(m, w)) continue;
            if (!same_part(m, v, w)) { if (prio(w, seed_t) > pv) eligible = 0; continue; }
            if (rdiv && is_root && root_cand(w, seed_t, rdiv) && prio(w, seed_t) > pv) is_root = 0;
        }
        if (!eligible) level[v] = LVL_DEAD;
        else if (is_root) level[v] = 0;
    }
    /* A root candidate that lost against a neighbour that is itself ineligible stays a
     * non-root: the rule only looks at candidate status, which keeps it one-pass. */
    uint32_t maxl = 0;
    for (uint32_t r = 1; r <= pr->rounds; ++r) {
        int joined = 0;
        #pragma omp parallel for schedule(static) reduction(|:joined) num_threads(g_mrf_threads) if (F > 8192)
        for (int64_t vv = 0; vv < (int64_t)F; ++vv) {
            uint32_t v = (uint32_t)vv;
            if (level[v] != LVL_NONE) continue;
            uint32_t c = 0;
            for (uint32_t a = m->adj_ptr[v]; a < m->adj_ptr[v + 1]; ++a) {
                uint32_t w = m->adj_idx[a];
                if (same_part(m, v, w) && level[w] < r) ++c;
            }
            if (c >= 2) { level[v] = LVL_DEAD; continue; }
            if (c != 1) continue;
            uint32_t pv = prio(v, seed_t);
            int win = 1;
            for (uint32_t a = m->adj_ptr[v]; a < m->adj_ptr[v + 1] && win; ++a) {
                uint32_t w = m->adj_idx[a];
                if (!same_part(m, v, w)) continue;
                uint32_t lw = level[w];
                if (!(lw == LVL_NONE || lw == r)) continue; /* decided before this round (or dead) */
                if (prio(w, seed_t) < pv) continue;
                uint32_t cw = 0;
                for (uint32_t b = m->adj_ptr[w]; b < m->adj_ptr[w + 1]; ++b) {
                    uint32_t x = m->adj_idx[b];
                    if (same_part(m, w, x) && level[x] < r) ++cw;
                }
                if (cw == 1) win = 0;
            }
            if (win) { level[v] = r; joined = 1; }
        }
        if (joined) maxl = r;
        /* no early exit: the CUDA path runs a fixed number of rounds; extra rounds are no-ops
         * only if nothing can join, which is not guaranteed, so keep going */
    }
    *max_level_out = maxl;
}

void orc_mrf_sample_forest(uint32_t F, const uint32_t *adj_ptr, const uint32_t *adj_idx,
                           const uint64_t *ptr, const orc_mrf_params *pr, uint32_t iteration,
                           uint32_t *level_out)
{
    mrf_t m = {F, adj_ptr, adj_idx, ptr, NULL, NULL, 0};
    uint32_t P = pr->num_parts ? pr->num_parts : 1;
    m.part_size = (F + P - 1) / P; if (!m.part_size) m.part_size = 1;
    uint32_t maxl;
    sample_forest(&m, pr, iteration, level_out, &maxl);
}

/* position of label `lab` in node w's sorted label list, or -1 */
static inline int64_t find_label(const mrf_t *m, uint32_t w, uint32_t lab)
{
    uint64_t lo = m->ptr[w], hi = m->ptr[w + 1];
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        uint32_t l = (uint32_t)m->view[mid] + 1u;
        if (l < lab) lo = mid + 1; else hi = mid;
    }
    if (lo < m->ptr[w + 1] && (uint32_t)m->view[lo] + 1u == lab) return (int64_t)lo;
    return -1;
}

int orc_view_selection(uint32_t F, const uint32_t *adj_ptr, const uint32_t *adj_idx,
                       const uint64_t *ptr, const uint16_t *view, const float *cost,
                       const orc_mrf_params *pr, int num_threads, uint32_t *labels,
                       double *trace, orc_mrf_info *info)
{
#ifdef _OPENMP
    g_mrf_threads = num_threads > 0 ? num_threads : omp_get_max_threads();
    if (g_mrf_threads > 16) g_mrf_threads = 16;
#else
    (void)num_threads;
#endif
    mrf_t m = {F, adj_ptr, adj_idx, ptr, view, cost, 0};
    uint32_t P = pr->num_parts ? pr->num_parts : 1;
    m.part_size = (F + P - 1) / P; if (!m.part_size) m.part_size = 1;
    uint64_t nnz = ptr[F];
    float *H = (float *)malloc(sizeof(float) * (nnz ? nnz : 1));
    float *hminp1 = (float *)malloc(sizeof(float) * (F ? F : 1));
    uint32_t *amin = (uint32_t *)malloc(sizeof(uint32_t) * (F ? F : 1));
    uint32_t *level = (uint32_t *)malloc(sizeof(uint32_t) * (F ? F : 1));
    uint32_t *order = (uint32_t *)malloc(sizeof(uint32_t) * (F ? F : 1));
    uint32_t *lvl_ptr = (uint32_t *)malloc(sizeof(uint32_t) * (pr->rounds + 2));
    int64_t *efix = (int64_t *)malloc(sizeof(int64_t) * (pr->max_iterations + 1));
    uint64_t unseen = 0;

    /* initial labeling: arg min of the unary (first minimum); unseen -> 0 */
    for (uint32_t i = 0; i < F; ++i) {
        if (!seen(&m, i)) { labels[i] = 0; ++unseen; continue; }
        uint64_t best = ptr[i];
        for (uint64_t k = ptr[i] + 1; k < ptr[i + 1]; ++k) if (cost[k] < cost[best]) best = k;
        labels[i] = (uint32_t)view[best] + 1u;
    }
    efix[0] = orc_mrf_energy_fixed(F, adj_ptr, adj_idx, ptr, view, cost, labels);
    if (trace) trace[0] = (double)efix[0] / 4294967296.0;
    info->energy_initial = orc_mrf_energy(F, adj_ptr, adj_idx, ptr, view, cost, labels);

    uint32_t t = 0;
    for (t = 1; t <= pr->max_iterations; ++t) {
        uint32_t maxl = 0;
        sample_forest(&m, pr, t, level, &maxl);
        /* bucket nodes by level (stable: ascending node id inside a level) */
        memset(lvl_ptr, 0, sizeof(uint32_t) * (pr->rounds + 2));
        for (uint32_t v = 0; v < F; ++v) if (level[v] <= pr->rounds) lvl_ptr[level[v] + 1]++;
        for (uint32_t r = 0; r <= pr->rounds; ++r) lvl_ptr[r + 1] += lvl_ptr[r];
        {
            uint32_t *pos = (uint32_t *)malloc(sizeof(uint32_t) * (pr->rounds + 2));
            memcpy(pos, lvl_ptr, sizeof(uint32_t) * (pr->rounds + 2));
            for (uint32_t v = 0; v < F; ++v) if (level[v] <= pr->rounds) order[pos[level[v]]++] = v;
            free(pos);
        }
        /* bottom-up min-sum messages */
        for (int64_t r = (int64_t)pr->rounds; r >= 0; --r) {
            #pragma omp parallel for schedule(dynamic, 256) num_threads(g_mrf_threads) if (lvl_ptr[r + 1] - lvl_ptr[r] > 2048)
            for (int64_t oi = lvl_ptr[r]; oi < (int64_t)lvl_ptr[r + 1]; ++oi) {
                uint32_t v = order[oi];
                float hmin = INFINITY;
                uint32_t hidx = 0;
                for (uint64_t k = ptr[v]; k < ptr[v + 1]; ++k) {
                    uint32_t lab = (uint32_t)view[k] + 1u;
                    float h = cost[k];
                    for (uint32_t a = adj_ptr[v]; a < adj_ptr[v + 1]; ++a) {
                        uint32_t w = adj_idx[a];
                        if (!seen(&m, w)) continue;
                        uint32_t lw = level[w];
                        if (same_part(&m, v, w) && lw <= pr->rounds) {
                            if (lw > (uint32_t)r) { /* child */
                                float msg = hminp1[w];
                                int64_t j = find_label(&m, w, lab);
                                if (j >= 0 && H[j] < msg) msg = H[j];
                                h = h + msg;
                            } /* else parent: skip */
                        } else {
                            h = h + (lab != labels[w] ? 1.0f : 0.0f);
                        }
                    }
                    H[k] = h;
                    if (h < hmin) { hmin = h; hidx = (uint32_t)(k - ptr[v]); }
                }
                hminp1[v] = hmin + 1.0f;
                amin[v] = hidx;
            }
        }
        /* top-down assignment */
        for (uint32_t r = 0; r <= pr->rounds; ++r) {
            #pragma omp parallel for schedule(static) num_threads(g_mrf_threads) if (lvl_ptr[r + 1] - lvl_ptr[r] > 8192)
            for (int64_t oi = lvl_ptr[r]; oi < (int64_t)lvl_ptr[r + 1]; ++oi) {
                uint32_t v = order[oi];
                uint32_t best = (uint32_t)view[ptr[v] + amin[v]] + 1u;
                if (r > 0) {
                    uint32_t xp = 0; int have = 0;
                    for (uint32_t a = adj_ptr[v]; a < adj_ptr[v + 1]; ++a) {
                        uint32_t w = adj_idx[a];
                        if (same_part(&m, v, w) && level[w] < r) { xp = labels[w]; have = 1; break; }
                    }
                    if (have) {
                        int64_t j = find_label(&m, v, xp);
                        if (j >= 0 && H[j] <= hminp1[v]) best = xp;
                    }
                }
                labels[v] = best;
            }
        }
        efix[t] = orc_mrf_energy_fixed(F, adj_ptr, adj_idx, ptr, view, cost, labels);
        if (trace) trace[t] = (double)efix[t] / 4294967296.0;
        if (t >= pr->window) { /* StopWhenReturnsDiminish, view_selection.cpp:84 */
            double e0 = (double)efix[t - pr->window], e1 = (double)efix[t];
            if (e0 <= 0.0 || (e0 - e1) / e0 < (double)pr->ratio) break;
        }
    }
    if (t > pr->max_iterations) t = pr->max_iterations;
    info->iterations = t;
    info->energy_final = orc_mrf_energy(F, adj_ptr, adj_idx, ptr, view, cost, labels);
    info->unseen = unseen;
    free(H); free(hminp1); free(amin); free(level); free(order); free(lvl_ptr); free(efix);
    return 0;
}

double orc_mrf_brute_force(uint32_t F, const uint32_t *adj_ptr, const uint32_t *adj_idx,
                           const uint64_t *ptr, const uint16_t *view, const float *cost,
                           uint32_t *labels_out)
{
    uint32_t *cur = (uint32_t *)calloc(F ? F : 1, sizeof(uint32_t));
    uint32_t *lab = (uint32_t *)calloc(F ? F : 1, sizeof(uint32_t));
    double best = INFINITY;
    for (;;) {
        for (uint32_t i = 0; i < F; ++i)
            lab[i] = ptr[i + 1] > ptr[i] ? (uint32_t)view[ptr[i] + cur[i]] + 1u : 0u;
        double e = orc_mrf_energy(F, adj_ptr, adj_idx, ptr, view, cost, lab);
        if (e < best) { best = e; memcpy(labels_out, lab, sizeof(uint32_t) * F); }
        uint32_t i = 0;
        for (; i < F; ++i) {
            uint64_t n = ptr[i + 1] - ptr[i];
            if (n == 0) continue;
            if (++cur[i] < n) break;
            cur[i] = 0;
        }
        if (i == F) break;
    }
    free(cur); free(lab);
    return best;
}
