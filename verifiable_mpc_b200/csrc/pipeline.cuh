// Backend-independent launch sequence of the MSM / fold pipeline.  `BE` is either the CUDA backend (vmsm.cu) or the
// host-emulation backend used by the CPU tests (tests/hostemu/hostemu.cpp); both run the kernel bodies of kernels.cuh.
#pragma once
#include <stddef.h>
#include "kernels.cuh"
#include "kernels_w.cuh"
#include "sc25519.cuh"

namespace vmsm {

enum Phase {
    PH_DIGITS = 0,
    PH_SCAN = 1,
    PH_SCATTER = 2,
    PH_ORDER = 3,
    PH_HANDOFF = 4,  // sort stream -> main stream: time the sorted lists wait for the previous MSM's accumulate kernel
    PH_ACCUMULATE = 5,
    PH_REDUCE = 6,
    PH_FINAL = 7,
    PH_COUNT = 8
};

// number of MSM tails (upper tree levels + Horner + inversion) that may be in flight at once, each on its own side
// stream with its own node buffers: the tails are latency-bound single-warp work, so several of them overlap freely
// (the eight MSMs of a Pinocchio proof, back-to-back commitments)
enum { kTailWays = 8, kTailWaysEd = 4, kTailWaysBN = 8 };

struct MsmOptions {
    uint32_t window_bits = 0;  // 0 = auto
    uint32_t reduce_log2r = 3;     // Ed25519 bucket tree: radix 8
    uint32_t reduce_log2r_w = 2;   // BN256: radix 4 (48 instead of 64 dependent additions down the 4096-bucket tree;
                                   // G2 2^14-term MSM 1.62 -> 1.54 ms, profiles/r01/bn256_msm_v5_windows.md)
    bool sort_buckets = true;
    uint32_t cap_factor = 8;   // a bucket's own thread sums at most max(64, cap_factor * n / NB) entries (kernels.cuh)
    uint32_t w_quad_acc = 1;       // BN256 accumulate with four lanes per bucket: 0 never, 1 by size, 2 always
    uint32_t quad_threshold = 16384;  // tree levels with at most this many output nodes run quad-cooperative; 0 = never
    uint32_t pre_sets = 0;     // MSMs over precomputed bases: number of bucket sets the windows share; 0 = auto
    uint32_t seg_len = 0;      // entries per thread of the balanced accumulate kernel; 0 = whole waves (seg_plan)
    uint32_t seg_mode = 1;     // accumulate kernel: 0 one thread per bucket, 1 by geometry (default), 2 segments
    uint32_t w_quad_fix = 1;   // BN256 segment fix-up with four lanes per bucket (0: one thread per bucket)
    uint32_t pre_sets_w = 0;   // BN256 MSMs over key tables: bucket sets shared by the windows; 0 = auto
    uint32_t seg_len_w = 0;    // BN256: entries per thread of the balanced accumulate kernel; 0 = auto
    // counting sort: 0 (default) two passes with global atomics, 1 block-privatised (shared-memory counters) where the
    // backend supports the geometry.  Measured NOT faster on B200 (profiles/r02/block_sort_experiment.md): shared-memory
    // atomics on random counters run at ~1-2 per clock per SM, and the largest carveout the sort blocks need slows the
    // accumulate kernel by 14 %; kept as a tested option and as the evidence for the atomic sort
    uint32_t block_sort = 0;
};

// Precomputed bases of an MSM call (KPrecompute): level w of `table` holds 2^(c*w) * P_i at table[w * stride + i].
// The MSM's `bases` argument then points at its first term inside level 0.
struct PreTable {
    uint32_t stride, c, W;
    const ge_niels *extra_table;  // table of the extra terms (same c, W), or null when the call has none
    uint32_t extra_stride;
};

struct Workspace {
    // outputs of the counting sort, double-buffered by MSM parity: the sort of MSM k+1 (atomics-bound) runs on its
    // own stream underneath the accumulate kernel of MSM k (integer-pipe-bound)
    uint32_t *counts_[2] = {nullptr, nullptr}, *offsets_[2] = {nullptr, nullptr}, *cursor_[2] = {nullptr, nullptr},
             *order_[2] = {nullptr, nullptr};  // W*NB each
    uint32_t *idx_[2] = {nullptr, nullptr};    // W*n
    int cur = 0;                               // parity in use by the MSM being issued

    // W*NB accumulator points (ge_ext for Ed25519, wjac<F> for BN256), sized in bytes; one set per tail way, because
    // the whole bucket tree of MSM k (leaf level included) runs on side stream k % kTailWays underneath the accumulate
    // kernels of the MSMs that follow
    void *buckets_[kTailWays] = {};
    // bucket-tree levels: [tail way][ping-pong]
    void *nodeS[kTailWays][2] = {}, *nodeT[kTailWays][2] = {};
    // long-bucket overflow (kernels.cuh: KOverflow / KCombine)
    // one set per tail way: the overflow / combine kernels of MSM k run on its side stream as well, while the
    // accumulate kernel of MSM k+1 is already filling the next set
    OverflowCtl *ctl_[kTailWays] = {};
    OverflowTask *tasks_[kTailWays] = {};
    LongBucket *longs_[kTailWays] = {};
    void *partials_[kTailWays] = {};
    // segmented accumulate (Ed25519): per-segment start bucket, entry total and row totals ride with the CSR parity;
    // the two partial sums per segment belong to the MSM's tail way (read by the fix-up kernels on its side stream)
    uint32_t *seg_bucket_[2] = {nullptr, nullptr}, *seg_total_[2] = {nullptr, nullptr}, *row_totals_[2] = {nullptr, nullptr};
    void *seg_partials_[kTailWays] = {};
    size_t cap_seg = 0, cap_seg_part = 0;  // segments; bytes of one way's partial sums
    size_t cap_buckets = 0, cap_idx = 0, cap_nodes = 0, cap_tasks = 0;  // element counts
    int ways = 0;  // tail ways that have buckets / node buffers (grows to the largest count requested)
    size_t elem_bytes = 0;  // size of one accumulator point the point buffers were allocated for
};

// Work model of SURVEY.md App. E (limb products), with the tree reduction's ~30 % overhead over a serial running sum.
// Windows whose top digit has few live bits are skipped: scalars are reduced below a ~2^(scalar_bits-1) group order,
// so the last window that holds bit (scalar_bits-2) sees r = (scalar_bits-1) mod c live bits and its 2^(r-1) buckets
// are 2^(c-r) times fuller than average (r = 0: a carry-only window, one bucket with n/2 entries).  The long-bucket
// path keeps such cases correct and bounded; the chooser simply avoids paying for them (measured: profiles/r01).
inline uint32_t choose_window(uint64_t n, uint32_t scalar_bits, bool avoid_skew = true, double madd = 504.0,
                              double add = 648.0 * 1.3) {
    if (n == 0) return 4;
    double best = 0;
    uint32_t best_c = 0;
    for (uint32_t c = 3; c <= 17; c++) {
        uint32_t r = (scalar_bits - 1) % c;
        if (avoid_skew && (r == 0 || c - r > 4)) continue;
        // below c = 11 every size measured (2^6 .. 2^16) is latency-bound and c = 11 was fastest or tied
        // (profiles/r01/sweep_tiny_windows.jsonl): few entries per bucket, short serial chains
        if (avoid_skew && c < 11) continue;
        // from ~2^15.5 terms the 2^10 buckets of c = 11 hold 64+ entries each and their serial chains bound the
        // launch; c = 15 (16 x fewer entries per bucket) is 14-36 % faster at 2^16 / 2^17 although the model, which
        // counts work, not chain length, still prefers 11 there (profiles/r01/sweep_mid_windows.jsonl)
        if (avoid_skew && n >= 49152 && c < 15) continue;
        double W = (double)((scalar_bits + 1 + c - 1) / c);
        double NB = (double)(1u << (c - 1));
        double cost = (double)n * W * madd + W * NB * 2.0 * add + (W - 1) * c * 464.0;
        if (!best_c || cost < best) {
            best = cost;
            best_c = c;
        }
    }
    return best_c ? best_c : 8;
}

inline MsmGeom make_geom(uint32_t n, uint32_t c, uint32_t scalar_bits, uint32_t sets = 0) {
    MsmGeom g;
    g.n = n;
    g.c = c;
    g.W = (scalar_bits + 1 + c - 1) / c;
    g.NB = 1u << (c - 1);
    g.S = sets && sets < g.W ? sets : g.W;
    g.lg = 0;
    while ((1ull << g.lg) < n) g.lg++;
    return g;
}

// Segment plan of the balanced accumulate kernel: whole waves of `resident` threads, at most 32 entries per thread.
struct SegPlan {
    uint32_t L, T;  // entries per segment, segments (threads) launched
};
// `fill_pct`: a launch that is a SINGLE wave is sized to this share of the resident threads instead of all of them.
// Measured on the table-based Ed25519 path (profiles/r02/segment_length.md): 2^16 terms 0.157 -> 0.146 ms with 0.58 of a
// wave, 2^17 terms 0.256 -> 0.234 ms with 0.69 -- fewer, longer segments leave fewer partial sums to fix up and leave
// room on the SMs for the sort of the next MSM and the tails of the previous ones, which run at the same time.
inline SegPlan seg_plan(uint64_t max_entries, uint32_t resident, uint32_t forced_L = 0, uint32_t min_L = 8,
                        uint32_t fill_pct = 100) {
    SegPlan p;
    if (max_entries == 0) max_entries = 1;
    uint64_t waves = (max_entries + 32ull * resident - 1) / (32ull * resident);
    uint64_t L = (max_entries + waves * resident - 1) / (waves * resident);
    if (waves == 1 && fill_pct < 100) {
        uint64_t threads = (uint64_t)resident * fill_pct / 100;
        if (threads == 0) threads = 1;
        L = (max_entries + threads - 1) / threads;
    }
    if (L < min_L) L = min_L;  // small MSMs: fewer threads with a handful of entries each
    if (forced_L) L = forced_L;
    p.L = (uint32_t)L;
    p.T = (uint32_t)((max_entries + L - 1) / L);
    return p;
}

template <class BE>
int ws_ensure_seg(BE &be, Workspace &ws, size_t segs, size_t elem_bytes = sizeof(ge_ext)) {
    if (segs > ws.cap_seg) {
        ws.cap_seg = 0;
        for (int k = 0; k < 2; k++) {
            be.free(ws.seg_bucket_[k]), be.free(ws.seg_total_[k]), be.free(ws.row_totals_[k]);
            ws.seg_bucket_[k] = (uint32_t *)be.alloc(segs * 4);
            ws.seg_total_[k] = (uint32_t *)be.alloc(16);
            ws.row_totals_[k] = (uint32_t *)be.alloc(256 * 4);  // one per bucket set: at most 127 (c = 2)
            if (!ws.seg_bucket_[k] || !ws.seg_total_[k] || !ws.row_totals_[k]) return -1;
        }
        ws.cap_seg = segs;
    }
    if (2 * segs * elem_bytes > ws.cap_seg_part) {  // after ws_ensure: ws.ways is final; capacity in bytes
        ws.cap_seg_part = 0;
        for (int k = 0; k < ws.ways; k++) {
            be.free(ws.seg_partials_[k]);
            ws.seg_partials_[k] = be.alloc(2 * segs * elem_bytes);
            if (!ws.seg_partials_[k]) return -1;
        }
        ws.cap_seg_part = 2 * segs * elem_bytes;
    }
    return 0;
}

template <class BE>
int ws_ensure(BE &be, Workspace &ws, const MsmGeom &g, uint32_t R, size_t elem_bytes = sizeof(ge_ext),
              uint32_t cap_floor = 64, uint32_t seg_floor = 256, int ways = kTailWaysEd) {
    size_t nb = (size_t)g.S * g.NB, ni = (size_t)g.W * g.n, nn = (size_t)g.S * ((g.NB + R - 1) / R);
    if (nn < g.S) nn = g.S;
    if (ways > ws.ways) {  // more MSM tails in flight than before: (re)allocate the per-way buffers
        ws.cap_buckets = ws.cap_nodes = ws.cap_tasks = ws.cap_seg_part = 0;
        ws.ways = ways;
    }
    if (elem_bytes > ws.elem_bytes) {  // a wider accumulator type than before: regrow the point buffers
        ws.cap_buckets = ws.cap_nodes = ws.cap_tasks = 0;
        ws.elem_bytes = elem_bytes;
    }
    elem_bytes = ws.elem_bytes;
    if (nb > ws.cap_buckets) {
        ws.cap_buckets = 0;
        for (int k = 0; k < ws.ways; k++) {
            be.free(ws.buckets_[k]);
            ws.buckets_[k] = be.alloc(nb * elem_bytes);
            if (!ws.buckets_[k]) return -1;
        }
        for (int k = 0; k < 2; k++) {
            be.free(ws.counts_[k]), be.free(ws.offsets_[k]), be.free(ws.cursor_[k]), be.free(ws.order_[k]);
            ws.counts_[k] = (uint32_t *)be.alloc(nb * 4);
            ws.offsets_[k] = (uint32_t *)be.alloc(nb * 4);
            ws.cursor_[k] = (uint32_t *)be.alloc(nb * 4);
            ws.order_[k] = (uint32_t *)be.alloc(nb * 4);
            if (!ws.counts_[k] || !ws.offsets_[k] || !ws.cursor_[k] || !ws.order_[k]) return -1;
        }
        ws.cap_buckets = nb;
    }
    if (ni > ws.cap_idx) {
        ws.cap_idx = 0;
        for (int k = 0; k < 2; k++) {
            be.free(ws.idx_[k]);
            ws.idx_[k] = (uint32_t *)be.alloc((ni ? ni : 1) * 4);
            if (!ws.idx_[k]) return -1;
        }
        ws.cap_idx = ni;
    }
    // >= sum over long buckets of their task counts: at most ni / cap_floor long buckets, each with one partly filled
    // segment, plus ni / seg_floor full segments
    size_t nt = ni / cap_floor + ni / seg_floor + 64;
    if (nt > ws.cap_tasks) {
        ws.cap_tasks = 0;
        for (int k = 0; k < ws.ways; k++) {
            be.free(ws.ctl_[k]), be.free(ws.tasks_[k]), be.free(ws.longs_[k]), be.free(ws.partials_[k]);
            ws.ctl_[k] = (OverflowCtl *)be.alloc(sizeof(OverflowCtl));
            ws.tasks_[k] = (OverflowTask *)be.alloc(nt * sizeof(OverflowTask));
            ws.longs_[k] = (LongBucket *)be.alloc(nt * sizeof(LongBucket));
            ws.partials_[k] = be.alloc(nt * elem_bytes);
            if (!ws.ctl_[k] || !ws.tasks_[k] || !ws.longs_[k] || !ws.partials_[k]) return -1;
        }
        ws.cap_tasks = nt;
    }
    if (nn > ws.cap_nodes) {
        for (int par = 0; par < ws.ways; par++)
            for (int k = 0; k < 2; k++) {
                be.free(ws.nodeS[par][k]), be.free(ws.nodeT[par][k]);
                ws.nodeS[par][k] = be.alloc(nn * elem_bytes);
                ws.nodeT[par][k] = be.alloc(nn * elem_bytes);
                if (!ws.nodeS[par][k] || !ws.nodeT[par][k]) {
                    ws.cap_nodes = 0;
                    return -1;
                }
            }
        ws.cap_nodes = nn;
    }
    return 0;
}

template <class BE>
void ws_release(BE &be, Workspace &ws) {
    for (int k = 0; k < kTailWays; k++) be.free(ws.buckets_[k]);
    for (int k = 0; k < 2; k++)
        be.free(ws.counts_[k]), be.free(ws.offsets_[k]), be.free(ws.cursor_[k]), be.free(ws.order_[k]), be.free(ws.idx_[k]);
    for (int k = 0; k < kTailWays; k++)
        be.free(ws.ctl_[k]), be.free(ws.tasks_[k]), be.free(ws.longs_[k]), be.free(ws.partials_[k]);
    for (int par = 0; par < kTailWays; par++)
        for (int k = 0; k < 2; k++) be.free(ws.nodeS[par][k]), be.free(ws.nodeT[par][k]);
    for (int k = 0; k < 2; k++) be.free(ws.seg_bucket_[k]), be.free(ws.seg_total_[k]), be.free(ws.row_totals_[k]);
    for (int k = 0; k < kTailWays; k++) be.free(ws.seg_partials_[k]);
    ws = Workspace();
}

// out = sum_i scalars[i] * bases[i].  All pointers are backend ("device") memory.  Returns 0 or -1 (allocation).
template <class BE>
int msm_run(BE &be, Workspace &ws, const MsmOptions &opt, uint32_t scalar_bits, const ge_niels *bases,
            const uint32_t *scalars, uint32_t n, ge_ext *out_ext, ge_aff *out_aff, uint32_t seq = 0,
            const ge_niels *extra = nullptr, uint32_t n_extra = 0, const PreTable *pre = nullptr,
            ge_ext *out_host_ext = nullptr) {
    // terms 0 .. n-n_extra-1 use `bases`, the last n_extra terms use `extra` (scalars are contiguous)
    const uint32_t n_main = n - n_extra;
    // precomputed bases fix the window (the table was built for it); the windows then share `S` bucket sets
    uint32_t c = pre ? pre->c : opt.window_bits ? opt.window_bits : choose_window(n, scalar_bits);
    // Two accumulate kernels (measured, profiles/r02/accumulate_segmented.md): one thread per BUCKET when there are many
    // more buckets than resident threads (the plain path: W bucket sets; large MSMs over tables: 8 sets), equal
    // SEGMENTS of the CSR array when the windows of a table-based MSM share one bucket set (few, long buckets).
    const uint32_t W_c = (scalar_bits + c) / c;
    // ... and on the plain path in the c = 11 regime (2^11 <= n < 49152: 24 x 1024 buckets of up to 48 entries, fewer
    // buckets than resident threads): 2^13 terms 0.096 -> 0.090 ms, 2^15 0.138 -> 0.119 ms, and the accumulate phase of a
    // lone MSM -- the cross-term commitments of a folding round -- 0.165 -> 0.102 ms (profiles/r02/segment_length.md)
    const bool seg = opt.seg_mode == 2 || (opt.seg_mode == 1 && ((pre && n < (1u << 19)) || (!pre && n >= 2048 && n < 49152)));
    const uint32_t sets = !pre ? 0u : opt.pre_sets ? opt.pre_sets : seg ? 1u : (W_c < 8 ? W_c : 8u);
    MsmGeom g = make_geom(n, c, scalar_bits, sets);
    if (pre && (pre->W != g.W || (n_extra && !pre->extra_table))) return -2;
    uint32_t R = 1u << opt.reduce_log2r;
    // small MSMs are bound by their tails (~0.45 ms of dependent doublings against a ~0.1 ms head): more of them in flight
    const int ways = n < (1u << 18) ? (int)kTailWays : (int)kTailWaysEd;
    if (ws_ensure(be, ws, g, R, sizeof(ge_ext), 64, 256, ways)) return -1;
    uint32_t nbuckets = g.S * g.NB;
    const int par = (int)(seq & 1);
    uint32_t *counts = ws.counts_[par], *offsets = ws.offsets_[par], *cursor = ws.cursor_[par], *idx = ws.idx_[par];
    uint32_t log2NB = g.c - 1;
    BaseRef br = {bases, extra, n_main, pre ? pre->stride : 0u, pre ? pre->extra_stride : 0u, g.lg, g.S, log2NB};
    if (pre) br.extra = pre->extra_table;
    // balanced accumulate: the CSR array (at most n * W entries) in equal segments, whole waves of resident threads
    SegPlan sp = {0, 0};
    uint32_t *seg_bucket = nullptr, *seg_total = nullptr;
    if (seg) {
        sp = seg_plan((uint64_t)n * g.W, be.resident_threads(pre != nullptr), opt.seg_len, 8, 65);
        if (ws_ensure_seg(be, ws, sp.T)) return -1;
        seg_bucket = ws.seg_bucket_[par], seg_total = ws.seg_total_[par];
    }

    // counting sort of (window, |digit|) -> CSR lists, on the sort stream (double-buffered by parity): block-privatised
    // (digits recoded once, shared-memory counters, kernels.cuh: KRecode) where the backend offers it for this geometry,
    // else two passes over the scalars with global atomics
    const uint32_t bs_chunks = opt.block_sort ? be.bsort_chunks(g) : 0u;
    if (bs_chunks && be.bsort_ensure(g, bs_chunks, par)) return -1;
    be.use_head(seq);
    be.sort_begin(par);
    be.phase_begin();
    if (bs_chunks) {
        be.bsort_hist(scalars, g, bs_chunks, par, counts);
    } else {
        be.zero(counts, (size_t)nbuckets * 4);
        if (n) {
            KDigitsHist k1 = {scalars, counts, g};
            be.launch_sort(k1, n);
        }
    }
    be.phase_mark(PH_DIGITS);
    if (seg) be.scan_offsets_flat(counts, offsets, cursor, g, ws.row_totals_[par], seg_bucket, seg_total, sp.L);
    else be.scan_offsets(counts, offsets, cursor, g);
    be.phase_mark(PH_SCAN);
    if (bs_chunks) {
        be.bsort_scatter(g, bs_chunks, par, offsets, idx);
    } else if (n) {
        KScatter k3 = {scalars, cursor, idx, g};
        be.launch_sort(k3, n);
    }
    be.phase_mark(PH_SCATTER);
    const uint32_t *order = nullptr;  // per-bucket kernel: buckets by decreasing population (segments are equal anyway)
    if (!seg && opt.sort_buckets && be.order_buckets(counts, ws.order_[par], nbuckets, n)) order = ws.order_[par];
    be.phase_mark(PH_ORDER);
    be.sort_end(par);
    be.phase_mark(PH_HANDOFF);
    // this MSM's buckets, bucket tree and Horner chain live in the buffers of tail way `tw`: wait until the MSM that
    // used them kTailWays issues ago has left them
    const int tw = (int)(seq % (uint32_t)ways);
    be.head_wait_tail(tw);
    void *const buckets = ws.buckets_[tw];
    const uint32_t wins_per_set = (g.W + g.S - 1) / g.S;
    be.zero(ws.ctl_[tw], sizeof(OverflowCtl));
    if (seg) {
        ge_ext *partials = (ge_ext *)ws.seg_partials_[tw];
        if (pre) {
            KAccumulateSegPre k5 = {br, offsets, counts, idx, seg_bucket, seg_total, (ge_ext *)buckets, partials, sp.L};
            be.launch(k5, sp.T);
        } else {
            KAccumulateSeg k5 = {br, offsets, counts, idx, seg_bucket, seg_total, (ge_ext *)buckets, partials, sp.L};
            be.launch(k5, sp.T);
        }
        be.phase_mark(PH_ACCUMULATE);
        // everything after the accumulate kernel belongs to this MSM's tail: partial sums of straddling buckets, bucket
        // tree, Horner -- on the CUDA backend on side stream `tw`, so the next MSM's accumulate kernel follows at once
        be.tail_begin(tw);
        const uint32_t long_span = 32;
        KSegFixup kf = {offsets, counts, partials, (ge_ext *)buckets, sp.L, long_span, ws.ctl_[tw], ws.longs_[tw]};
        be.launch(kf, nbuckets);
        be.acc_done(par);  // the CSR lists of this parity are free again (the fix-up was their last reader)
        if ((uint64_t)n * wins_per_set > (uint64_t)sp.L * long_span) {  // otherwise no bucket can span that many segments
            const uint32_t ow = be.overflow_warps();
            KSegLongFix kl = {ws.ctl_[tw], ws.longs_[tw], partials, (ge_ext *)buckets, ow};
            be.launch(kl, ow * 32);
        }
    } else {
        uint32_t cap = opt.cap_factor * (uint32_t)(((uint64_t)n * wins_per_set) >> (g.c - 1));
        if (cap < 64) cap = 64;
        if (pre) {
            KAccumulatePre k5 = {br, offsets, counts, idx, order, (ge_ext *)buckets, nbuckets, cap, ws.ctl_[tw],
                                 ws.tasks_[tw], ws.longs_[tw]};
            be.launch(k5, nbuckets);
        } else {
            KAccumulate k5 = {br, offsets, counts, idx, order, (ge_ext *)buckets, nbuckets, cap, ws.ctl_[tw], ws.tasks_[tw],
                              ws.longs_[tw]};
            be.launch(k5, nbuckets);
        }
        be.phase_mark(PH_ACCUMULATE);
        // everything after the accumulate kernel belongs to this MSM's tail: long-bucket overflow tasks, combine, bucket
        // tree, Horner -- on the CUDA backend on side stream `tw`, so the next MSM's accumulate kernel follows at once
        be.tail_begin(tw);
        if ((uint64_t)n * wins_per_set > cap) {  // otherwise no bucket can be long
            const uint32_t ow = be.overflow_warps();
            if (pre) {
                KOverflowPre ko = {br, idx, ws.ctl_[tw], ws.tasks_[tw], (ge_ext *)ws.partials_[tw], ow};
                be.launch(ko, ow * 32);
            } else {
                KOverflow ko = {br, idx, ws.ctl_[tw], ws.tasks_[tw], (ge_ext *)ws.partials_[tw], ow};
                be.launch(ko, ow * 32);
            }
            be.acc_done(par);  // the CSR lists of this parity are free again (the overflow tasks were their last reader)
            const uint32_t ct = be.combine_threads();
            KCombine kc = {ws.ctl_[tw], ws.longs_[tw], (const ge_ext *)ws.partials_[tw], (ge_ext *)buckets, ct};
            be.launch(kc, ct);
        } else {
            be.acc_done(par);
        }
    }
    // bucket tree (S,T radix-R levels, quad-cooperative once few nodes remain) + Horner over the windows
    const ge_ext *inS = (const ge_ext *)buckets, *inT = nullptr;
    uint32_t cnt = g.NB, log2s = 0;
    int pp = 0;
    do {
        uint32_t cnt_out = (cnt + R - 1) / R;
        uint32_t nodes = g.S * cnt_out;
        if (nodes <= opt.quad_threshold) {
            KReduceQ k6 = {inS, inT, (ge_ext *)ws.nodeS[tw][pp], (ge_ext *)ws.nodeT[tw][pp], cnt, cnt_out, R, log2s, nodes};
            be.launch(k6, (4 * nodes + 31) & ~31u);
        } else {
            KReduce k6 = {inS, inT, (ge_ext *)ws.nodeS[tw][pp], (ge_ext *)ws.nodeT[tw][pp], cnt, cnt_out, R, log2s};
            be.launch(k6, nodes);
        }
        inS = (const ge_ext *)ws.nodeS[tw][pp];
        inT = (const ge_ext *)ws.nodeT[tw][pp];
        pp ^= 1;
        cnt = cnt_out;
        log2s += opt.reduce_log2r;
    } while (cnt > 1);
    be.phase_mark(PH_REDUCE);
    {
        // one root per bucket set; over precomputed bases the sets carry no weight: zero doublings between them
        const uint32_t dbl = pre ? 0u : g.c;
        if (opt.quad_threshold) {
            KFinalQ k7 = {inS, inT, out_ext, out_aff, g.S, dbl, out_host_ext};
            be.launch(k7, 32);
        } else {
            KFinal k7 = {inS, inT, out_ext, out_aff, g.S, dbl, out_host_ext};
            be.launch(k7, 1);
        }
    }
    be.phase_mark(PH_FINAL);
    be.after_final(out_ext, out_aff);  // multi-GPU: push this partial to the owner's mailbox / gather on the owner
    be.result_ready();
    be.tail_end(tw);
    be.head_done(seq);  // everything this MSM put on its head stream (the tail too when side streams are off)
    be.phase_end();
    return 0;
}

// ---------------------------------------------------------------------------------------------- BN256 (G1 / G2)
// Same sequence as msm_run for the Weierstrass groups; everything on the main stream (these MSMs are 2^14-sized).
template <class BE, class F>
int msm_run_w(BE &be, Workspace &ws, const MsmOptions &opt, const waff<F> *bases, const uint32_t *scalars, uint32_t n,
              wjac<F> *out_jac, waff<F> *out_wire, const waff<F> *extra = nullptr, uint32_t n_extra = 0,
              uint32_t seq = 0, const PreTable *pre = nullptr, const waff<F> *extra_table = nullptr,
              wjac<F> *out_host_jac = nullptr) {
    const uint32_t scalar_bits = 256, n_main = n - n_extra;
    // Jacobian costs in field multiplications (mixed addition 7M+4S, addition 11M+5S); the long-bucket path makes any
    // window safe, so skewed top windows are allowed here (these MSMs are small and latency matters more)
    uint32_t c = opt.window_bits ? opt.window_bits : choose_window(n, scalar_bits, false, 11.0, 16.0 * 0.6);
    if (c > 16) c = 16;
    // 2^9 .. 2^16 terms are latency-bound on this curve (few, expensive additions per thread): c = 13 gives 20 x 4096
    // short bucket chains and a bucket tree of exactly six full radix-4 levels (4096 = 4^6); measured fastest or tied at 2^10, 2^12,
    // 2^14 and 2^16 on G1 and G2 (profiles/r01/bn256_msm_v5_windows.md), 35 % faster than the work-minimising c = 11
    if (!opt.window_bits && n >= 512 && n <= (1u << 16)) c = 13;
    // plain path from 2^11 terms: equal segments of the sorted entries instead of one thread per bucket, and then the
    // chain length no longer depends on the window, so up to 2^14 terms the work-minimising c = 11 wins (1024 buckets
    // per window: a bucket tree a quarter the size).  Measured, G1 / G2 ms per MSM (profiles/r02/bn256_shared_sets_segments.md):
    // 2^12 0.252 -> 0.219 / 0.742 -> 0.615, 2^14 0.329 -> 0.279 / 0.825 -> 0.758, 2^16 (c = 13) 0.828 -> 0.502 / 2.63 -> 1.49
    const bool seg_plain = !pre && n >= 2048 && opt.seg_mode == 1;
    if (seg_plain && !opt.window_bits && n < (1u << 15)) c = 11;
    if (pre) c = pre->c;  // tables fix the window
    // Over key tables the bucket sets carry no weight (no Horner chain), so the W windows may share S < W sets: the
    // bucket tree, which costs as much as the accumulation at these sizes (2.3 full additions per bucket against n * W
    // / (W * NB) = 4 mixed additions per bucket at 2^14 terms), shrinks by W / S.  Fewer, fuller buckets are then summed
    // by equal SEGMENTS of the sorted entries (KAccumulateSegW), as on the Ed25519 path.
    const uint32_t W_c = (scalar_bits + c) / c;
    const bool seg = (n >= 256 && (opt.seg_mode == 2 || (opt.seg_mode == 1 && pre))) || seg_plain;
    const uint32_t sets = !pre ? 0u : opt.pre_sets_w ? opt.pre_sets_w : seg ? 2u : W_c;
    MsmGeom g = make_geom(n, c, scalar_bits, sets);
    if (pre && (pre->W != g.W || (n_extra && !extra_table))) return -2;
    BaseRefW<F> br = {bases, pre ? extra_table : extra, n_main, pre ? pre->stride : 0u, pre ? pre->extra_stride : 0u, g.c - 1,
                      g.lg, g.S};
    uint32_t R = 1u << opt.reduce_log2r_w;
    // long-bucket granularity for this curve: additions are 3-9x dearer than on Ed25519 and the MSMs are small, so a
    // bucket's own thread takes at most max(16, ...) entries and overflow segments may be as short as one warp pass
    const uint32_t kCapFloor = 16, kSegFloor = 32;
    if (ws_ensure(be, ws, g, R, sizeof(wjac<F>), kCapFloor, kSegFloor, kTailWaysBN)) return -1;
    uint32_t nbuckets = g.S * g.NB;
    const uint32_t wins_per_set = (g.W + g.S - 1) / g.S;
    const int par = (int)(seq & 1);
    uint32_t *counts = ws.counts_[par], *offsets = ws.offsets_[par], *cursor = ws.cursor_[par], *idx = ws.idx_[par];
    SegPlan sp = {0, 0};
    uint32_t *seg_bucket = nullptr, *seg_total = nullptr;
    if (seg) {
        // at least 12 entries per thread: at 2^14 terms that is half a wave of threads, which measured faster than a
        // full wave of 8-entry segments (fewer partial sums to fix up; several MSMs are in flight anyway) and than 16
        // (profiles/r02/bn256_shared_sets_segments.md)
        sp = seg_plan((uint64_t)n * g.W, be.template resident_threads_w<F>(), opt.seg_len_w, 12);
        if (ws_ensure_seg(be, ws, sp.T, sizeof(wjac<F>))) return -1;
        seg_bucket = ws.seg_bucket_[par], seg_total = ws.seg_total_[par];
    }
    // over key tables consecutive MSMs alternate between two head streams, as on the Ed25519 path (a launch is half a
    // wave: two of them overlap; 2^12 terms G1 0.118 -> 0.100 ms, G2 0.286 -> 0.230 ms); the plain path measured slower
    // that way (2^14 terms G1 0.267 -> 0.303 ms) and keeps its accumulate kernels on one stream
    if (pre) be.use_head(seq);
    be.sort_begin(par);
    be.phase_begin();
    be.zero(counts, (size_t)nbuckets * 4);
    if (n) {
        KDigitsHist k1 = {scalars, counts, g};
        be.launch_sort(k1, n);
    }
    be.phase_mark(PH_DIGITS);
    if (seg) be.scan_offsets_flat(counts, offsets, cursor, g, ws.row_totals_[par], seg_bucket, seg_total, sp.L);
    else be.scan_offsets(counts, offsets, cursor, g);
    be.phase_mark(PH_SCAN);
    if (n) {
        KScatter k3 = {scalars, cursor, idx, g};
        be.launch_sort(k3, n);
    }
    be.phase_mark(PH_SCATTER);
    const uint32_t *order = nullptr;
    if (!seg && opt.sort_buckets && be.order_buckets(counts, ws.order_[par], nbuckets, n)) order = ws.order_[par];
    be.phase_mark(PH_ORDER);
    be.sort_end(par);
    be.phase_mark(PH_HANDOFF);
    // eight tails in flight: a lone tail on this curve takes 1.5 ms (G1) to 3.6 ms (G2), the head 0.3-1 ms
    const int tw = (int)(seq % kTailWaysBN);
    be.head_wait_tail(tw);
    void *const buckets = ws.buckets_[tw];
    be.zero(ws.ctl_[tw], sizeof(OverflowCtl));
    if (seg) {
        wjac<F> *partials = (wjac<F> *)ws.seg_partials_[tw];
        KAccumulateSegW<F> k5 = {br, offsets, counts, idx, seg_bucket, seg_total, (wjac<F> *)buckets, partials, sp.L};
        be.launch(k5, sp.T);
        be.phase_mark(PH_ACCUMULATE);
        be.tail_begin(tw);
        const uint32_t long_span = 32;
        // four lanes per bucket over key tables (few, full buckets spanning ~4 segments: 2^14 terms G2 0.383 -> 0.357 ms,
        // 2^16 terms G1 0.467 -> 0.424 ms); the plain path's many short buckets measured 2-3 % slower that way
        if (opt.w_quad_fix && pre) {
            KSegFixupWQ<F> kf = {offsets, counts, partials, (wjac<F> *)buckets, sp.L, long_span, ws.ctl_[tw], ws.longs_[tw]};
            be.launch(kf, 4 * nbuckets);
        } else {
            KSegFixupW<F> kf = {offsets, counts, partials, (wjac<F> *)buckets, sp.L, long_span, ws.ctl_[tw], ws.longs_[tw]};
            be.launch(kf, nbuckets);
        }
        be.acc_done(par);
        if ((uint64_t)n * wins_per_set > (uint64_t)sp.L * long_span) {
            const uint32_t ow = be.overflow_warps();
            KSegLongFixW<F> kl = {ws.ctl_[tw], ws.longs_[tw], partials, (wjac<F> *)buckets, ow};
            be.launch(kl, ow * 32);
        }
    } else {
        uint32_t cap = opt.cap_factor * (uint32_t)(((uint64_t)n * wins_per_set) >> (g.c - 1));
        if (cap < kCapFloor) cap = kCapFloor;
        // four lanes per bucket where measured faster (2^15: -8 %, 2^16: -14 %; slower at <= 2^14, where the tails set
        // the pace, and from 2^17, where the kernel is throughput-bound): profiles/r01/bn256_msm_v5_windows.md
        if (opt.w_quad_acc == 2 || (opt.w_quad_acc == 1 && n > (1u << 14) && n <= (1u << 16))) {
            KAccumulateWQ<F> k5 = {br, offsets, counts, idx, order, (wjac<F> *)buckets, nbuckets, cap, ws.ctl_[tw],
                                   ws.tasks_[tw], ws.longs_[tw], kSegFloor};
            be.launch(k5, 4 * nbuckets);
        } else {
            KAccumulateW<F> k5 = {br, offsets, counts, idx, order, (wjac<F> *)buckets, nbuckets, cap, ws.ctl_[tw],
                                  ws.tasks_[tw], ws.longs_[tw], kSegFloor};
            be.launch(k5, nbuckets);
        }
        be.phase_mark(PH_ACCUMULATE);
        // overflow tasks, combine, bucket tree, Horner and inversion on side stream `tw`, underneath the heads of the
        // following MSMs (the eight MSMs of a Pinocchio proof are independent)
        be.tail_begin(tw);
        if ((uint64_t)n * wins_per_set > cap) {
            const uint32_t ow = be.overflow_warps();
            KOverflowW<F> ko = {br, idx, ws.ctl_[tw], ws.tasks_[tw], (wjac<F> *)ws.partials_[tw], ow};
            be.launch(ko, ow * 32);
            be.acc_done(par);
            const uint32_t ct = be.combine_threads();
            KCombineW<F> kc = {ws.ctl_[tw], ws.longs_[tw], (const wjac<F> *)ws.partials_[tw], (wjac<F> *)buckets, ct};
            be.launch(kc, ct);
        } else {
            be.acc_done(par);
        }
    }
    const wjac<F> *inS = (const wjac<F> *)buckets, *inT = nullptr;
    uint32_t cnt = g.NB, log2s = 0;
    int pp = 0;
    do {
        uint32_t cnt_out = (cnt + R - 1) / R;
        // four lanes per node: these MSMs are small, the tree is latency-bound at every level
        KReduceWQ<F> k6 = {inS, inT, (wjac<F> *)ws.nodeS[tw][pp], (wjac<F> *)ws.nodeT[tw][pp], cnt, cnt_out, R, log2s};
        be.launch(k6, 4 * g.S * cnt_out);
        inS = (const wjac<F> *)ws.nodeS[tw][pp];
        inT = (const wjac<F> *)ws.nodeT[tw][pp];
        pp ^= 1;
        cnt = cnt_out;
        log2s += opt.reduce_log2r_w;
    } while (cnt > 1);
    be.phase_mark(PH_REDUCE);
    KFinalWQ<F> k7 = {inS, inT, out_jac, out_wire, g.S, pre ? 0u : g.c, out_host_jac};
    be.launch(k7, 32);
    be.phase_mark(PH_FINAL);
    be.result_ready();
    be.tail_end(tw);
    if (pre) be.head_done(seq);
    be.phase_end();
    return 0;
}

// Host-side construction of the BN256 fixed-base tables tbl[w][j-1] = j * 16^w * G, w < 65, j = 1..8 (Montgomery affine)
template <class F>
inline void build_fixed_base_table_w(waff<F> *tbl /* 520 entries, host memory */) {
    waff<F> G = {F::gen_x(), F::gen_y()};
    wjac<F> base = wa_to_jac(G);
    for (int w = 0; w < 65; w++) {
        wjac<F> m = base;
        for (int j = 1; j <= 8; j++) {
            tbl[w * 8 + (j - 1)] = wj_to_aff(m);
            if (j < 8) m = wj_add(m, base);
        }
        for (int k = 0; k < 4; k++) base = wj_dbl(base);
    }
}

// Non-adjacent form of a 256-bit little-endian scalar as two bit masks; returns the top non-zero digit index or -1.
inline int32_t naf_masks(const uint32_t s_in[8], uint32_t nz[9], uint32_t ng[9]) {
    uint32_t s[10];
    for (int i = 0; i < 8; i++) s[i] = s_in[i];
    s[8] = s[9] = 0;
    for (int i = 0; i < 9; i++) nz[i] = ng[i] = 0;
    int32_t top = -1;
    for (int i = 0; i < 288; i++) {
        bool any = false;
        for (int k = 0; k < 10; k++) any |= s[k] != 0;
        if (!any) break;
        if (s[0] & 1u) {
            if ((s[0] & 3u) == 3u) {  // digit -1: s += 1
                ng[i >> 5] |= 1u << (i & 31);
                uint64_t cy = 1;
                for (int k = 0; k < 10 && cy; k++) {
                    cy += s[k];
                    s[k] = (uint32_t)cy;
                    cy >>= 32;
                }
            } else {  // digit +1: s -= 1
                s[0] &= ~1u;
            }
            nz[i >> 5] |= 1u << (i & 31);
            top = i;
        }
        for (int k = 0; k < 9; k++) s[k] = (s[k] >> 1) | (s[k + 1] << 31);
        s[9] >>= 1;
    }
    return top;
}

// In place fold of a point vector held as (aff, niels): P[j] = c*P[j] + P[half+j].  `tmp` holds `half` extended points.
template <class BE>
void fold_run(BE &be, ge_aff *aff, ge_niels *niels, ge_ext *tmp, uint32_t half, const uint32_t c[8],
              uint32_t quad_max_half = 1u << 13) {
    if (half <= quad_max_half) {  // latency-bound: too few elements to fill the machine with one thread each
        KFoldQ kq;
        kq.aff = aff;
        kq.niels = niels;
        kq.half = half;
        kq.top = naf_masks(c, kq.nz, kq.ng);
        be.launch(kq, (4 * half + 31) & ~31u);
        return;
    }
    KFold kf;
    kf.niels = niels;
    kf.out = tmp;
    kf.half = half;
    kf.top = naf_masks(c, kf.nz, kf.ng);
    be.launch(kf, half);
    KNormalize kn = {tmp, aff, niels};
    be.launch(kn, half);
}

// Host-side construction of the fixed-base table tbl[w][j-1] = j * 16^w * B (niels), w < 64, j = 1..8.
inline void build_fixed_base_table(ge_niels *tbl /* 512 entries, host memory */) {
    ge_aff B;
    const uint32_t bx[8] = {0x8f25d51au, 0xc9562d60u, 0x9525a7b2u, 0x692cc760u, 0xfdd6dc5cu, 0xc0a4e231u, 0xcd6e53feu, 0x216936d3u};
    const uint32_t by[8] = {0x66666658u, 0x66666666u, 0x66666666u, 0x66666666u, 0x66666666u, 0x66666666u, 0x66666666u, 0x66666666u};
    for (int i = 0; i < 8; i++) B.x.v[i] = bx[i], B.y.v[i] = by[i];
    ge_ext base = ge_aff_to_ext(B);
    for (int w = 0; w < 64; w++) {
        ge_ext m = base;
        for (int j = 1; j <= 8; j++) {
            tbl[w * 8 + (j - 1)] = ge_aff_to_niels(ge_ext_to_aff(m));
            if (j < 8) m = ge_add(m, base);
        }
        for (int k = 0; k < 4; k++) base = ge_dbl(base);
    }
}

}  // namespace vmsm
