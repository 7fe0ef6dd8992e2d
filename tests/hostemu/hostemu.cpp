// Host emulation of the kernel bodies (TESTS ONLY -- never loaded by verifiable_mpc_b200/).
// Compiles csrc/kernels.cuh + csrc/pipeline.cuh with g++ (portable 64-bit arithmetic path of fe25519.cuh) and runs
// each functor in a plain loop, so the pipeline's indexing, digit recoding, bucket tree, Horner step, fold and
// fixed-base logic can be checked against oracle/ in the GPU-less build container.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "../../verifiable_mpc_b200/csrc/pipeline.cuh"

using namespace vmsm;

static uint32_t g_bs_chunks = 0;  // hostemu_set_block_sort

struct HostBE {
    void *alloc(size_t bytes) { return aligned_alloc(64, (bytes + 63) / 64 * 64 + 64); }
    void free(void *p) { ::free(p); }
    void zero(void *p, size_t bytes) { memset(p, 0, bytes); }
    template <class F>
    void launch(const F &f, uint32_t n) {
        for (uint32_t t = 0; t < n; t++) f(t);
    }
    template <class F>
    void launch_sort(const F &f, uint32_t n) {
        launch(f, n);
    }
    void scan_offsets(const uint32_t *counts, uint32_t *offsets, uint32_t *cursor, const MsmGeom &g) {
        for (uint32_t w = 0; w < g.S; w++) {  // one row per bucket set (S == W unless the bases are precomputed)
            uint32_t run = geom_set_start(g, w);
            for (uint32_t b = 0; b < g.NB; b++) {
                offsets[(size_t)w * g.NB + b] = cursor[(size_t)w * g.NB + b] = run;
                run += counts[(size_t)w * g.NB + b];
            }
        }
    }
    void scan_offsets_flat(const uint32_t *counts, uint32_t *offsets, uint32_t *cursor, const MsmGeom &g,
                           uint32_t *row_totals, uint32_t *seg_bucket, uint32_t *total, uint32_t L) {
        uint32_t run = 0;
        for (uint32_t w = 0; w < g.S; w++) {
            uint32_t start = run;
            for (uint32_t i = 0; i < g.NB; i++) {
                uint32_t b = w * g.NB + i, cnt = counts[b];
                offsets[b] = cursor[b] = run;
                for (uint32_t k = (run + L - 1) / L; (uint64_t)k * L < (uint64_t)run + cnt; k++) seg_bucket[k] = b;
                run += cnt;
            }
            row_totals[w] = run - start;
        }
        *total = run;
    }
    // block-privatised counting sort, restated as loops (CUDA: vmsm_bsort_* in vmsm.cu): KRecode is the shared kernel
    // body; per (set, chunk) histograms, exclusive prefix over the chunks, per-block cursors
    uint32_t bs_chunks = g_bs_chunks;  // 0 = atomic two-pass sort (default)
    std::vector<uint16_t> bs_dig;
    std::vector<uint32_t> bs_hist;
    uint32_t bsort_chunks(const MsmGeom &g) { return g.c <= 16 && g.n ? bs_chunks : 0; }
    int bsort_ensure(const MsmGeom &g, uint32_t C, int) {
        bs_dig.assign((size_t)g.W * ((g.n + 7u) & ~7u), 0xabcd);
        bs_hist.assign((size_t)g.S * C * g.NB, 0);
        return 0;
    }
    template <class F>
    void bsort_each(const MsmGeom &g, uint32_t C, F f) {
        const uint32_t stride = (g.n + 7u) & ~7u, per = (((g.n + C - 1) / C) + 7u) & ~7u;
        for (uint32_t s = 0; s < g.S; s++)
            for (uint32_t ch = 0; ch < C; ch++) {
                const uint32_t lo = ch * per;
                if (lo >= g.n) continue;
                const uint32_t hi = g.n - lo < per ? g.n : lo + per;
                for (uint32_t w = s, k = 0; w < g.W; w += g.S, k++)
                    for (uint32_t i = lo; i < hi; i++) {
                        uint32_t code = bs_dig[(size_t)w * stride + i];
                        if (code != 0xffffu) f(s, ch, k, i, code);
                    }
            }
    }
    void bsort_hist(const uint32_t *scalars, const MsmGeom &g, uint32_t C, int, uint32_t *counts) {
        KRecode k0 = {scalars, bs_dig.data(), (g.n + 7u) & ~7u, g};
        launch(k0, g.n);
        bsort_each(g, C, [&](uint32_t s, uint32_t ch, uint32_t, uint32_t, uint32_t code) {
            bs_hist[((size_t)s * C + ch) * g.NB + (code & 0x7fffu)]++;
        });
        for (uint32_t s = 0; s < g.S; s++)
            for (uint32_t b = 0; b < g.NB; b++) {
                uint32_t run = 0;
                for (uint32_t ch = 0; ch < C; ch++) {
                    uint32_t &v = bs_hist[((size_t)s * C + ch) * g.NB + b], t = v;
                    v = run;
                    run += t;
                }
                counts[(size_t)s * g.NB + b] = run;
            }
    }
    void bsort_scatter(const MsmGeom &g, uint32_t C, int, const uint32_t *offsets, uint32_t *idx) {
        for (uint32_t s = 0; s < g.S; s++)
            for (uint32_t ch = 0; ch < C; ch++)
                for (uint32_t b = 0; b < g.NB; b++) bs_hist[((size_t)s * C + ch) * g.NB + b] += offsets[(size_t)s * g.NB + b];
        bsort_each(g, C, [&](uint32_t s, uint32_t ch, uint32_t k, uint32_t i, uint32_t code) {
            uint32_t pos = bs_hist[((size_t)s * C + ch) * g.NB + (code & 0x7fffu)]++;
            idx[pos] = i | (k << g.lg) | ((code & 0x8000u) << 16);
        });
    }
    uint32_t resident_threads(bool) { return resident; }
    template <class F>
    uint32_t resident_threads_w() { return resident; }
    uint32_t resident = 48;  // small on purpose: several waves and straddling buckets even in tiny test cases
    bool order_buckets(const uint32_t *counts, uint32_t *order, uint32_t nb, uint32_t n) {
        if (n == 0) return false;
        std::iota(order, order + nb, 0u);
        std::stable_sort(order, order + nb, [&](uint32_t a, uint32_t b) { return counts[a] > counts[b]; });
        return true;
    }
    uint32_t overflow_warps() { return 3; }
    uint32_t combine_threads() { return 5; }
    void use_head(uint32_t) {}
    void head_done(uint32_t) {}
    void sort_begin(int) {}
    void sort_end(int) {}
    void acc_done(int) {}
    void after_final(ge_ext *, ge_aff *) {}
    void result_ready() {}
    void head_wait_tail(int) {}
    void tail_begin(int) {}
    void tail_end(int) {}
    void phase_begin() {}
    void phase_mark(int) {}
    void phase_end() {}
};

static uint32_t g_seg_mode = 1, g_seg_len = 0, g_bn_sets = 0;

extern "C" {

// accumulate-kernel selection for the following hostemu_msm* calls (MsmOptions::seg_mode / seg_len)
// counting-sort selection: 0 = two passes with atomics, C > 0 = block-privatised with C scalar chunks per bucket set
void hostemu_set_block_sort(uint32_t chunks) { g_bs_chunks = chunks; }

void hostemu_set_bn_sets(uint32_t sets) { g_bn_sets = sets; }  // MsmOptions::pre_sets_w

void hostemu_set_seg(uint32_t mode, uint32_t len) {
    g_seg_mode = mode;
    g_seg_len = len;
}

void hostemu_fe_op(int op, const uint32_t *a, const uint32_t *b, uint32_t *out) {
    HostBE be;
    fe x, y, r;
    memcpy(x.v, a, 32);
    memcpy(y.v, b, 32);
    KSelfTestFe k = {&x, &y, &r, op};
    be.launch(k, 1);
    memcpy(out, r.v, 32);
}

// returns the error word of the upload validation (0 = ok)
uint32_t hostemu_msm(const uint8_t *affine, const uint8_t *scalars, uint32_t n, uint32_t window_bits, int sort,
                     uint32_t log2r, uint8_t *out_affine, uint8_t *out_ext) {
    HostBE be;
    std::vector<ge_niels> niels(n ? n : 1);
    uint32_t err = 0;
    std::vector<ge_aff> aff(n ? n : 1);
    memcpy(aff.data(), affine, (size_t)n * 64);
    KAffToNiels k = {aff.data(), niels.data(), &err, 1u};
    be.launch(k, n);
    std::vector<uint32_t> sc((size_t)(n ? n : 1) * 8 + 8);
    memcpy(sc.data(), scalars, (size_t)n * 32);
    Workspace ws;
    MsmOptions opt;
    opt.window_bits = window_bits;
    opt.sort_buckets = sort != 0;
    opt.reduce_log2r = log2r;
    opt.seg_mode = g_seg_mode, opt.seg_len = g_seg_len;
    ge_ext oe;
    ge_aff oa;
    msm_run(be, ws, opt, 253, niels.data(), sc.data(), n, &oe, &oa);
    ws_release(be, ws);
    memcpy(out_affine, &oa, 64);
    if (out_ext) memcpy(out_ext, &oe, 128);
    return err;
}

// Pedersen form: n_main bases + n_extra extra bases (scalars contiguous), cf. vmsm_msm_ext
uint32_t hostemu_msm_ext(const uint8_t *affine, uint32_t n_main, const uint8_t *affine_extra, uint32_t n_extra,
                         const uint8_t *scalars, uint32_t window_bits, uint8_t *out_affine) {
    HostBE be;
    uint32_t err = 0;
    std::vector<ge_aff> aff(n_main ? n_main : 1), affx(n_extra ? n_extra : 1);
    std::vector<ge_niels> niels(n_main ? n_main : 1), nielsx(n_extra ? n_extra : 1);
    memcpy(aff.data(), affine, (size_t)n_main * 64);
    memcpy(affx.data(), affine_extra, (size_t)n_extra * 64);
    KAffToNiels k = {aff.data(), niels.data(), &err, 1u};
    be.launch(k, n_main);
    KAffToNiels kx = {affx.data(), nielsx.data(), &err, 1u};
    be.launch(kx, n_extra);
    uint32_t n = n_main + n_extra;
    std::vector<uint32_t> sc((size_t)(n ? n : 1) * 8 + 8);
    memcpy(sc.data(), scalars, (size_t)n * 32);
    Workspace ws;
    MsmOptions opt;
    opt.window_bits = window_bits;
    opt.seg_mode = g_seg_mode, opt.seg_len = g_seg_len;
    ge_ext oe;
    ge_aff oa;
    msm_run(be, ws, opt, 253, niels.data(), sc.data(), n, &oe, &oa, 0, nielsx.data(), n_extra);
    ws_release(be, ws);
    memcpy(out_affine, &oa, 64);
    return err;
}

// MSM over PRECOMPUTED bases (KPrecompute tables for `table_bits`), n_main main terms [off, off + n_main) of a vector
// of n_pts points plus n_extra extra terms with their own table, `sets` bucket sets (0 = auto); cf. run_msm_ps
int hostemu_msm_pre(const uint8_t *affine, uint32_t n_pts, uint32_t off, uint32_t n_main, const uint8_t *affine_extra,
                    uint32_t n_extra, const uint8_t *scalars, uint32_t table_bits, uint32_t sets, uint8_t *out_affine) {
    HostBE be;
    const uint32_t W = (253 + table_bits) / table_bits;
    std::vector<ge_aff> aff(n_pts ? n_pts : 1), affx(n_extra ? n_extra : 1);
    memcpy(aff.data(), affine, (size_t)n_pts * 64);
    memcpy(affx.data(), affine_extra, (size_t)n_extra * 64);
    std::vector<ge_niels> tbl((size_t)W * (n_pts ? n_pts : 1)), tblx((size_t)W * (n_extra ? n_extra : 1));
    KPrecompute kp = {aff.data(), tbl.data(), n_pts, table_bits, W};
    be.launch(kp, n_pts);
    KPrecompute kx = {affx.data(), tblx.data(), n_extra, table_bits, W};
    be.launch(kx, n_extra);
    uint32_t n = n_main + n_extra;
    std::vector<uint32_t> sc((size_t)(n ? n : 1) * 8 + 8);
    memcpy(sc.data(), scalars, (size_t)n * 32);
    Workspace ws;
    MsmOptions opt;
    opt.pre_sets = sets;
    opt.seg_mode = g_seg_mode, opt.seg_len = g_seg_len;
    PreTable pt = {n_pts, table_bits, W, n_extra ? tblx.data() : nullptr, n_extra};
    ge_ext oe;
    ge_aff oa;
    int rc = msm_run(be, ws, opt, 253, tbl.data() + off, sc.data(), n, &oe, &oa, 0, nullptr, n_extra, &pt);
    ws_release(be, ws);
    memcpy(out_affine, &oa, 64);
    return rc;
}

void hostemu_fold(const uint8_t *affine, uint32_t n, const uint8_t *c_le32, uint8_t *out_affine) {
    HostBE be;
    uint32_t half = n / 2, err = 0;
    std::vector<ge_aff> aff(n);
    std::vector<ge_niels> niels(n);
    std::vector<ge_ext> tmp(half);
    memcpy(aff.data(), affine, (size_t)n * 64);
    KAffToNiels k = {aff.data(), niels.data(), &err, 0u};
    be.launch(k, n);
    uint32_t cs[8];
    memcpy(cs, c_le32, 32);
    fold_run(be, aff.data(), niels.data(), tmp.data(), half, cs);
    memcpy(out_affine, aff.data(), (size_t)half * 64);
}

void hostemu_fixed_base(const uint8_t *scalars, uint64_t seed, uint32_t n, uint8_t *out_affine) {
    HostBE be;
    std::vector<ge_niels> tbl(512);
    build_fixed_base_table(tbl.data());
    std::vector<uint32_t> sc;
    if (scalars) {
        sc.resize((size_t)n * 8 + 8);
        memcpy(sc.data(), scalars, (size_t)n * 32);
    }
    std::vector<ge_ext> tmp(n);
    std::vector<ge_aff> aff(n);
    std::vector<ge_niels> niels(n);
    KFixedBase k = {tbl.data(), scalars ? sc.data() : nullptr, seed, tmp.data()};
    be.launch(k, n);
    KNormalize kn = {tmp.data(), aff.data(), niels.data()};
    be.launch(kn, n);
    memcpy(out_affine, aff.data(), (size_t)n * 64);
}

void hostemu_synth_scalars(uint64_t seed, uint32_t n, uint8_t *out) {
    HostBE be;
    std::vector<uint32_t> sc((size_t)n * 8 + 8);
    KSynthScalars k = {sc.data(), seed};
    be.launch(k, n);
    memcpy(out, sc.data(), (size_t)n * 32);
}

uint32_t hostemu_choose_window(uint64_t n) { return choose_window(n, 253); }

// transcript text of n points: slots + lens as KPointText writes them
void hostemu_point_text(const uint8_t *affine, uint32_t n, uint8_t *slots, uint32_t *lens) {
    HostBE be;
    std::vector<ge_aff> aff(n ? n : 1);
    memcpy(aff.data(), affine, (size_t)n * 64);
    KPointText k = {aff.data(), slots, lens, n};
    be.launch(k, n);
}

// ---- scalar vectors modulo the Ed25519 group order (sc25519.cuh), the launch sequences of vmsm.cu
void hostemu_scalars_fold(uint32_t *v, uint32_t half, const uint8_t *c_le32, int mode) {
    HostBE be;
    scl cs;
    memcpy(cs.v, c_le32, 32);
    KScalarAxpy k = {v, v + 8ull * half, scl_to_mont(cs), mode};
    be.launch(k, half);
}

void hostemu_scalars_axpy(uint32_t *dst, const uint32_t *src, uint32_t n, const uint8_t *c_le32, int mode) {
    HostBE be;
    scl cs;
    memcpy(cs.v, c_le32, 32);
    KScalarAxpy k = {dst, src, scl_to_mont(cs), mode};
    be.launch(k, n);
}

void hostemu_scalars_dot(const uint32_t *a, const uint32_t *b, uint32_t n, uint8_t *out_le32) {
    HostBE be;
    memset(out_le32, 0, 32);
    if (!n) return;
    const uint32_t kT1 = 1u << 16, kT2 = 256;  // stage sizes as vmsm.cu (dot_stage_sizes)
    uint64_t t = n / 16;
    if (t < 64) t = n < 64 ? n : 64;
    if (t > kT1) t = kT1;
    uint32_t t2 = 1;
    while ((uint64_t)t2 * t2 < t && t2 < kT2) t2 <<= 1;
    uint32_t T = (uint32_t)t, T2 = t2 < t ? t2 : (uint32_t)t;
    std::vector<uint32_t> scratch((kT1 + kT2 + 1) * 8 + 16);
    uint32_t *p1 = (uint32_t *)(((uintptr_t)scratch.data() + 15) & ~(uintptr_t)15), *p2 = p1 + kT1 * 8, *p3 = p2 + kT2 * 8;
    KScalarDotPartial k1 = {a, b, n, T, p1};
    be.launch(k1, T);
    KScalarSum k2 = {p1, T, T2, p2, 0};
    be.launch(k2, T2);
    KScalarSum k3 = {p2, T2, 1, p3, 1};
    be.launch(k3, 1);
    memcpy(out_le32, p3, 32);
}

void hostemu_scalar_text(const uint32_t *v, uint32_t n, int is_signed, uint8_t *slots, uint32_t *lens) {
    HostBE be;
    KScalarText k = {v, slots, lens, n, is_signed};
    be.launch(k, n);
}

// ---- BN256: field ops (plain in, plain out), MSM and fixed-base for G1 (g2 = 0) / G2 (g2 = 1)
void hostemu_fbn_op(int op, const uint32_t *a, const uint32_t *b, uint32_t *out) {
    fbn x, y, r;
    memcpy(x.v, a, 32);
    memcpy(y.v, b, 32);
    x = fbn_to_mont(x);
    y = fbn_to_mont(y);
    switch (op) {
        case 0: r = fbn_add(x, y); break;
        case 1: r = fbn_sub(x, y); break;
        case 2: r = fbn_mul(x, y); break;
        case 3: r = fbn_inv(x); break;
        default: r = fbn_neg(x); break;
    }
    r = fbn_from_mont(r);
    memcpy(out, r.v, 32);
}

}  // extern "C"

template <class F>
static uint32_t bn_msm(const uint8_t *wire, const uint8_t *scalars, uint32_t n, uint32_t window_bits, uint8_t *out_wire) {
    HostBE be;
    uint32_t err = 0;
    std::vector<waff<F>> w(n ? n : 1), base(n ? n : 1);
    memcpy(w.data(), wire, (size_t)n * sizeof(waff<F>));
    KUploadW<F> ku = {w.data(), base.data(), &err, 1u};
    be.launch(ku, n);
    std::vector<uint32_t> sc((size_t)(n ? n : 1) * 8 + 8);
    memcpy(sc.data(), scalars, (size_t)n * 32);
    Workspace ws;
    MsmOptions opt;
    opt.window_bits = window_bits;
    opt.seg_mode = g_seg_mode, opt.seg_len_w = g_seg_len;
    wjac<F> oj;
    waff<F> ow;
    msm_run_w<HostBE, F>(be, ws, opt, base.data(), sc.data(), n, &oj, &ow);
    ws_release(be, ws);
    memcpy(out_wire, &ow, sizeof(ow));
    return err;
}
// MSM over key tables (KPrecomputeW, window `table_bits`): main terms [off, off + n_main) of n_pts points + n_extra
// extra terms with their own table; host-side normalisation of the Jacobian result as vmsm.cu's fetch_slot does it
template <class F>
static int bn_msm_pre(const uint8_t *wire, uint32_t n_pts, uint32_t off, uint32_t n_main, const uint8_t *wire_extra,
                      uint32_t n_extra, const uint8_t *scalars, uint32_t table_bits, uint8_t *out_wire) {
    HostBE be;
    uint32_t err = 0;
    const uint32_t W = (256 + table_bits) / table_bits;
    std::vector<waff<F>> w(n_pts ? n_pts : 1), base(n_pts ? n_pts : 1), wx(n_extra ? n_extra : 1), basex(n_extra ? n_extra : 1);
    memcpy(w.data(), wire, (size_t)n_pts * sizeof(waff<F>));
    memcpy(wx.data(), wire_extra, (size_t)n_extra * sizeof(waff<F>));
    KUploadW<F> ku = {w.data(), base.data(), &err, 1u};
    be.launch(ku, n_pts);
    KUploadW<F> kx = {wx.data(), basex.data(), &err, 1u};
    be.launch(kx, n_extra);
    std::vector<waff<F>> tbl((size_t)W * (n_pts ? n_pts : 1)), tblx((size_t)W * (n_extra ? n_extra : 1));
    KPrecomputeW<F> kp = {base.data(), tbl.data(), n_pts, table_bits, W};
    be.launch(kp, n_pts);
    KPrecomputeW<F> kpx = {basex.data(), tblx.data(), n_extra, table_bits, W};
    be.launch(kpx, n_extra);
    uint32_t n = n_main + n_extra;
    std::vector<uint32_t> sc((size_t)(n ? n : 1) * 8 + 8);
    memcpy(sc.data(), scalars, (size_t)n * 32);
    Workspace ws;
    MsmOptions opt;
    opt.seg_mode = g_seg_mode, opt.seg_len_w = g_seg_len, opt.pre_sets_w = g_bn_sets;
    PreTable pt = {n_pts, table_bits, W, nullptr, n_extra};
    wjac<F> oj, hj;
    waff<F> ow;
    int rc = msm_run_w<HostBE, F>(be, ws, opt, tbl.data() + off, sc.data(), n, &oj, &ow, nullptr, n_extra, 0, &pt,
                                  n_extra ? tblx.data() : nullptr, &hj);
    ws_release(be, ws);
    ow = wa_to_wire(wj_to_aff(hj));
    memcpy(out_wire, &ow, sizeof(ow));
    return rc ? rc : (int)err;
}
extern "C" int hostemu_bn_msm_pre(int g2, const uint8_t *wire, uint32_t n_pts, uint32_t off, uint32_t n_main,
                                  const uint8_t *wire_extra, uint32_t n_extra, const uint8_t *scalars, uint32_t table_bits,
                                  uint8_t *out_wire) {
    return g2 ? bn_msm_pre<Fp2BN>(wire, n_pts, off, n_main, wire_extra, n_extra, scalars, table_bits, out_wire)
              : bn_msm_pre<FpBN>(wire, n_pts, off, n_main, wire_extra, n_extra, scalars, table_bits, out_wire);
}

extern "C" uint32_t hostemu_bn_msm(int g2, const uint8_t *wire, const uint8_t *scalars, uint32_t n, uint32_t window_bits,
                        uint8_t *out_wire) {
    return g2 ? bn_msm<Fp2BN>(wire, scalars, n, window_bits, out_wire) : bn_msm<FpBN>(wire, scalars, n, window_bits, out_wire);
}

template <class F>
static void bn_fixed_base(const uint8_t *scalars, uint64_t seed, uint32_t n, uint8_t *out_wire) {
    HostBE be;
    std::vector<waff<F>> tbl(520);
    build_fixed_base_table_w<F>(tbl.data());
    std::vector<uint32_t> sc;
    if (scalars) {
        sc.resize((size_t)n * 8 + 8);
        memcpy(sc.data(), scalars, (size_t)n * 32);
    }
    std::vector<wjac<F>> tmp(n);
    std::vector<waff<F>> wire(n), base(n);
    KFixedBaseW<F> k = {tbl.data(), scalars ? sc.data() : nullptr, seed, tmp.data()};
    be.launch(k, n);
    KNormalizeW<F> kn = {tmp.data(), wire.data(), base.data()};
    be.launch(kn, n);
    memcpy(out_wire, wire.data(), (size_t)n * sizeof(waff<F>));
}
extern "C" void hostemu_bn_fixed_base(int g2, const uint8_t *scalars, uint64_t seed, uint32_t n, uint8_t *out_wire) {
    if (g2) bn_fixed_base<Fp2BN>(scalars, seed, n, out_wire);
    else bn_fixed_base<FpBN>(scalars, seed, n, out_wire);
}
