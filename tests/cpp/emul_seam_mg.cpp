// tests/cpp/emul_seam_mg.cpp -- multi-rank fiber emulation (cuda_fiber.h, emul::launch_ranks) of csrc/seam_mg.cu: the
// seam system is assembled once with the emulated kernels of csrc/seam.cu, then `ranks` emulated devices run k_pcg_mg
// concurrently, exchanging p / partial sums / barrier epochs through each other's blocks exactly as peers do through
// cudaIpc-mapped memory.  seam_mg_kernels.inc: kernel part of seam_mg.cu, text unchanged except (a) the two PTX
// ld/st helpers are replaced by the plain C++ below and (b) the static __shared__ reduction scratch becomes per-block
// scratch of the emulator (several blocks of different ranks are alive at the same time).
#include "cuda_fiber.h"

#include <algorithm>
#include <vector>

#include "oracle.h"   // orc_view only
#include "seam_kernels.inc"

namespace b2 {
namespace {
inline void st_release_sys(uint32_t *p, uint32_t v) { *(volatile uint32_t *)p = v; }
inline uint32_t ld_acquire_sys(const uint32_t *p) { return *(const volatile uint32_t *)p; }
inline uint32_t ld_acquire_gpu(const uint32_t *p) { return *(const volatile uint32_t *)p; }
inline unsigned long long mg_timer_ns() { return 0; }
inline void st_volatile_v2(uint2 *p, uint32_t a, uint32_t b) { uint2 v; v.x = a; v.y = b; *p = v; }
inline uint2 ld_volatile_v2(const uint2 *p) { return *p; }
inline void st_volatile_v4(float4 *p, float a, float b, float c, uint32_t tag)
{
    float4 v; v.x = a; v.y = b; v.z = c; memcpy(&v.w, &tag, 4); *p = v;
}
inline uint4 ld_volatile_v4(const float4 *p) { uint4 v; memcpy(&v, p, 16); return v; }
}  // namespace
}  // namespace b2
#include "seam_mg_kernels.inc"

using namespace b2;

extern "C" {

void emul_set_schedule(unsigned long long seed) { emul::set_schedule(seed); }   // 0 = round robin

// fault injection: rank `r` never enters the kernel (a peer that died); the others must come back with barrier timeouts
// instead of hanging.  spin = polls before a waiter gives up (small here: the emulator is slow).
static int g_dead_rank = -1;
static unsigned long long g_spin_limit = 1000000000ull;
void emul_set_dead_rank(int r, unsigned long long spin) { g_dead_rank = r; g_spin_limit = spin ? spin : 1000000000ull; }

void emul_seam_mg_free(void *p) { free(p); }

// x_out: [ranks][R][3] -- the complete solution as every rank ends up with it; status_out: [ranks][16]
int emul_seam_mg(const float *verts, uint32_t Vn, const uint32_t *faces, uint32_t F, const uint32_t *vf_ptr, const uint32_t *vf_idx,
                 const uint32_t *vv_ptr, const uint32_t *vv_idx, const uint32_t *labels, const orc_view *views, uint32_t K,
                 uint32_t ranks, uint32_t grid, uint32_t *R_out, float **x_out, uint32_t *status_out)
{
    (void)F;
    if (ranks < 1 || ranks > (uint32_t)MG_MAX_RANKS) return -2;
    std::vector<ViewDev> vd(K);
    for (uint32_t v = 0; v < K; ++v) {
        ViewDev &d = vd[v];
        for (int i = 0; i < 3; ++i) { d.pos[i] = views[v].pos[i]; d.dir[i] = views[v].viewdir[i]; }
        for (int i = 0; i < 9; ++i) d.proj[i] = views[v].proj[i];
        for (int i = 0; i < 12; ++i) d.w2c[i] = views[v].w2c[i];
        d.w = views[v].width; d.h = views[v].height; d.rgb = views[v].rgb; d.grad = nullptr; d.valid4 = nullptr;
    }
    // ---- assembly (seam_run with solve = false), identical on every rank ----
    const uint32_t vb = (Vn + 127) / 128;
    std::vector<uint32_t> cnt((size_t)Vn + 1, 0u), row_ptr((size_t)Vn + 1, 0u);
    uint32_t limit_flags = 0;   // overflow of the fixed per-vertex / per-edge caps (none on these scenes)
    emul::launch_serial(vb, 128, [&] { k_vertex_labels<false>(Vn, vf_ptr, vf_idx, labels, cnt.data(), nullptr, nullptr, nullptr, &limit_flags); });
    for (uint32_t i = 0; i < Vn; ++i) row_ptr[i + 1] = row_ptr[i] + cnt[i];
    const uint32_t R = row_ptr[Vn];
    std::vector<uint32_t> row_label(R ? R : 1), row_vertex(R ? R : 1);
    emul::launch_serial(vb, 128, [&] { k_vertex_labels<true>(Vn, vf_ptr, vf_idx, labels, nullptr, row_ptr.data(), row_label.data(), row_vertex.data(), &limit_flags); });
    SeamMesh m{verts, faces, vf_ptr, vf_idx, vv_ptr, vv_idx, labels, row_ptr.data(), row_label.data(), vd.data()};
    std::fill(cnt.begin(), cnt.end(), 0u);
    emul::launch_serial(vb, 128, [&] { k_arows<false>(Vn, m, cnt.data(), nullptr, nullptr, nullptr, &limit_flags); });
    std::vector<uint32_t> arow_ptr((size_t)Vn + 1, 0u);
    for (uint32_t i = 0; i < Vn; ++i) arow_ptr[i + 1] = arow_ptr[i] + cnt[i];
    const uint32_t A = arow_ptr[Vn];
    std::vector<uint32_t> arow_rows(2 * (size_t)A + 2);
    std::vector<float> arow_b(3 * (size_t)A + 3);
    emul::launch_serial(vb, 128, [&] { k_arows<true>(Vn, m, nullptr, arow_ptr.data(), arow_rows.data(), arow_b.data(), &limit_flags); });
    std::vector<uint32_t> rcnt((size_t)R + 1, 0u), csr_ptr((size_t)R + 1, 0u);
    const uint32_t rb = (R + 127) / 128;
    if (R) emul::launch_serial(rb, 128, [&] { k_matrix<false>(R, m, row_vertex.data(), arow_ptr.data(), arow_rows.data(), arow_b.data(), rcnt.data(),
                                                              nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr); });
    for (uint32_t r = 0; r < R; ++r) csr_ptr[r + 1] = csr_ptr[r] + rcnt[r];
    const uint32_t nnzL = csr_ptr[R];
    std::vector<uint32_t> csr_col(nnzL ? nnzL : 1), csr_enc(nnzL ? nnzL : 1);
    std::vector<float> csr_val(nnzL ? nnzL : 1), dval(R ? R : 1), inv_diag(R ? R : 1), rhs3(3 * (size_t)(R ? R : 1));
    if (R) emul::launch_serial(rb, 128, [&] { k_matrix<true>(R, m, row_vertex.data(), arow_ptr.data(), arow_rows.data(), arow_b.data(), nullptr,
                                                             csr_ptr.data(), csr_col.data(), csr_val.data(), inv_diag.data(), rhs3.data(), csr_enc.data(),
                                                             dval.data()); });
    *R_out = R;
    float *x = (float *)malloc(sizeof(float) * 3 * (size_t)(R ? R : 1) * ranks);
    *x_out = x;
    if (!R) return 0;

    // ---- one "device" per rank: own peer block, own scratch; peer tables point at each other's blocks ----
    std::vector<std::vector<char> > blocks(ranks, std::vector<char>(mg_block_bytes(R), 0));
    std::vector<std::vector<float> > rr(ranks, std::vector<float>(3 * (size_t)R)), tt(ranks, std::vector<float>(3 * (size_t)R));
    std::vector<std::vector<double> > bp(ranks, std::vector<double>((size_t)grid * 8, 0.0));
    std::vector<std::vector<uint32_t> > st(ranks, std::vector<uint32_t>(32, 0u));
    std::vector<PcgMg> q(ranks);
    std::vector<std::vector<uint8_t> > dest(ranks, std::vector<uint8_t>(R, 0xFF));   // 0xFF outside the own rows: never read
    std::vector<std::vector<uint8_t> > imark(ranks, std::vector<uint8_t>(R, 0));
    std::vector<std::vector<uint32_t> > imp(ranks, std::vector<uint32_t>(R, 0u));
    std::vector<uint32_t> nimp(ranks, 0u);
    uint32_t halo_rows = 0;
    for (uint32_t k = 0; k < ranks; ++k) {
        PcgMg &p = q[k];
        p.R = R; p.r0 = (uint32_t)((uint64_t)R * k / ranks); p.r1 = (uint32_t)((uint64_t)R * (k + 1) / ranks);
        p.rank = k; p.nranks = ranks;
        p.csr_ptr = csr_ptr.data(); p.csr_enc = csr_enc.data(); p.diag_val = dval.data(); p.inv_diag = inv_diag.data(); p.rhs = rhs3.data();
        p.r = rr[k].data(); p.t = tt[k].data(); p.blockpart = bp[k].data(); p.status = st[k].data();
        for (uint32_t j = 0; j < (uint32_t)MG_MAX_RANKS; ++j) p.peer[j] = j < ranks ? (void *)blocks[j].data() : nullptr;
        p.timing = 0u; p.max_iters = 1000u; p.tol = 0.0001f; p.epoch0 = 0xFFFFFFF0u;   // start close to the wrap-around of the epoch counter
        p.spin_limit = g_spin_limit;
        if (p.r1 > p.r0)
            emul::launch_serial((p.r1 - p.r0 + 255) / 256, 256, [&] { k_pcg_mg_dest(R, p.r0, p.r1, k, ranks, csr_ptr.data(), csr_enc.data(), dest[k].data(), imark[k].data()); });
        emul::launch_serial((R + 255) / 256, 256, [&] { k_pcg_mg_imports(R, imark[k].data(), imp[k].data(), &nimp[k]); });
        p.dest = dest[k].data(); p.imp = imp[k].data(); p.n_imp = &nimp[k];
        for (uint32_t i = p.r0; i < p.r1; ++i) halo_rows += dest[k][i] != 0;
    }
    st[0][15] = halo_rows;   // reported through the status words of rank 0
    for (uint32_t k = 0; k < ranks; ++k)      // flags start at epoch0 ("everybody reached the epochs used so far")
        for (uint32_t j = 0; j < (uint32_t)MG_MAX_RANKS; ++j) mg_carve(blocks[k].data(), R).flag[j] = 0xFFFFFFF0u;
    if (!emul::launch_ranks(ranks, grid, MG_THREADS, [&](unsigned rank) { if ((int)rank != g_dead_rank) k_pcg_mg(q[rank]); })) return -1;
    for (uint32_t k = 0; k < ranks; ++k) {
        const float *xs = mg_carve(blocks[k].data(), R).x;
        for (uint32_t r = 0; r < R; ++r)
            for (int c = 0; c < 3; ++c) x[((size_t)k * R + r) * 3 + c] = xs[(size_t)c * R + r];
        for (int i = 0; i < 16; ++i) status_out[16 * k + i] = st[k][i];
    }
    return 0;
}

}  // extern "C"
