// libvmsm.so -- CUDA backend and C ABI (include/vmsm.h) of the MSM / generator-fold engine.  sm_100a only.
// No PyTorch, no NCCL, no CPU fallback: every entry point needs a live CUDA context.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/vmsm.h"
#include "pipeline.cuh"

using namespace vmsm;

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int32_t fail(int32_t code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(VMSM_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));      \
    } while (0)

// ------------------------------------------------------------------------------------------------ kernels
template <class F>
struct min_blocks {
    template <class G>
    static constexpr int get(decltype(G::kMinBlocks) *) { return G::kMinBlocks; }
    template <class G>
    static constexpr int get(...) { return 1; }
    static constexpr int value = get<F>(nullptr);
};

// dynamic shared memory a kernel asks for WITHOUT using it: an occupancy limiter (kDynSmem in the functor)
template <class F>
struct dyn_smem {
    template <class G>
    static constexpr int get(decltype(G::kDynSmem) *) { return G::kDynSmem; }
    template <class G>
    static constexpr int get(...) { return 0; }
    static constexpr int value = get<F>(nullptr);
};

template <class F>
__global__ void __launch_bounds__(F::kBlock, min_blocks<F>::value) vmsm_kernel(const F f, uint32_t n) {
    uint32_t tid = blockIdx.x * (uint32_t)F::kBlock + threadIdx.x;
    if (tid < n) f(tid);
}

// grid-stride variant for the per-scalar kernels of the counting sort when they run underneath an accumulate kernel.
// Deliberately plain: batching the scatter's atomics / prefetching the next scalar made the sort itself ~30 % faster
// but slowed the accumulate kernel it shares the memory system with by more (profiles/r01/accumulate_variants.md).
template <class F>
__global__ void __launch_bounds__(F::kBlock) vmsm_kernel_strided(const F f, uint32_t n) {
    const uint32_t stride = gridDim.x * (uint32_t)F::kBlock;
    for (uint32_t tid = blockIdx.x * (uint32_t)F::kBlock + threadIdx.x; tid < n; tid += stride) {
        f(tid);
        if (tid + stride < tid) break;
    }
}

// Exclusive scan of the bucket populations, one block per (window, tile of 1024 buckets).  Each block first sums
// the tiles before it in the same window (coalesced re-read of at most NB counters from L2), then scans its own
// tile: one launch, W * NB / 1024 blocks, no inter-block dependency.  Offsets are absolute positions in idx.
#define SCAN_TILE 1024
__device__ __forceinline__ uint32_t block_scan_1024(uint32_t v, uint32_t *warp_sums, uint32_t *total) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t ws = warp_sums[lane], wi = ws;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= (uint32_t)d) wi += t;
        }
        warp_sums[lane] = wi - ws;
        if (lane == 31) *total = wi;
    }
    __syncthreads();
    return warp_sums[wid] + incl - v;  // exclusive prefix of v within the block
}

__global__ void __launch_bounds__(SCAN_TILE) vmsm_scan_offsets(const uint32_t *__restrict__ counts,
                                                               uint32_t *__restrict__ offsets,
                                                               uint32_t *__restrict__ cursor, MsmGeom g,
                                                               uint32_t tiles_per_window) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t total;
    const uint32_t w = blockIdx.x / tiles_per_window, tile = blockIdx.x - w * tiles_per_window, t = threadIdx.x;
    const uint32_t *cw = counts + (size_t)w * g.NB;
    uint32_t before = 0;
    for (uint32_t i = t; i < tile * SCAN_TILE; i += SCAN_TILE) before += cw[i];
    block_scan_1024(before, warp_sums, &total);
    uint32_t prefix = total;
    __syncthreads();
    uint32_t i = tile * SCAN_TILE + t;
    uint32_t cnt = i < g.NB ? cw[i] : 0u;
    uint32_t excl = block_scan_1024(cnt, warp_sums, &total);
    if (i < g.NB) {
        uint32_t off = geom_set_start(g, w) + prefix + excl;
        offsets[(size_t)w * g.NB + i] = off;
        cursor[(size_t)w * g.NB + i] = off;
    }
}

// Contiguous variant for the segmented accumulate kernel (Ed25519 path): the rows (bucket sets) follow each other
// without gaps, so the CSR array is one run of E = sum(counts) entries; the scan also records, for every segment of L
// entries, the bucket that holds its first entry, and E itself.
__global__ void __launch_bounds__(1024) vmsm_row_totals(const uint32_t *__restrict__ counts, uint32_t NB,
                                                        uint32_t *__restrict__ row_totals) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t total;
    const uint32_t *cw = counts + (size_t)blockIdx.x * NB;
    uint32_t s = 0;
    for (uint32_t i = threadIdx.x; i < NB; i += 1024) s += cw[i];
    block_scan_1024(s, warp_sums, &total);
    if (threadIdx.x == 0) row_totals[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_TILE) vmsm_scan_offsets_flat(const uint32_t *__restrict__ counts,
                                                                    uint32_t *__restrict__ offsets,
                                                                    uint32_t *__restrict__ cursor, MsmGeom g,
                                                                    uint32_t tiles_per_row,
                                                                    const uint32_t *__restrict__ row_totals,
                                                                    uint32_t *__restrict__ seg_bucket,
                                                                    uint32_t *__restrict__ total_out, uint32_t L) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t total;
    const uint32_t w = blockIdx.x / tiles_per_row, tile = blockIdx.x - w * tiles_per_row, t = threadIdx.x;
    const uint32_t *cw = counts + (size_t)w * g.NB;
    uint32_t row_base = 0;
    for (uint32_t r = 0; r < w; r++) row_base += row_totals[r];
    if (blockIdx.x == 0 && t == 0) {
        uint32_t e = 0;
        for (uint32_t r = 0; r < g.S; r++) e += row_totals[r];
        *total_out = e;
    }
    uint32_t before = 0;
    for (uint32_t i = t; i < tile * SCAN_TILE; i += SCAN_TILE) before += cw[i];
    block_scan_1024(before, warp_sums, &total);
    uint32_t prefix = total;
    __syncthreads();
    uint32_t i = tile * SCAN_TILE + t;
    uint32_t cnt = i < g.NB ? cw[i] : 0u;
    uint32_t excl = block_scan_1024(cnt, warp_sums, &total);
    if (i < g.NB) {
        uint32_t off = row_base + prefix + excl;
        uint32_t b = w * g.NB + i;
        offsets[b] = off;
        cursor[b] = off;
        // segments whose first entry lies in this bucket (usually none or one; a skewed bucket owns many)
        for (uint32_t k = (off + L - 1) / L; (uint64_t)k * L < (uint64_t)off + cnt; k++) seg_bucket[k] = b;
    }
}

// ---- block-privatised counting sort (kernels.cuh: KRecode).  Block (set s, chunk ch) owns the codes of the windows
// of bucket set s for scalars [ch * per, (ch + 1) * per) and the NB counters of the set in shared memory:
//   vmsm_bsort_hist     per-block histogram -> blockhist[s][ch][b]
//   vmsm_bsort_colscan  blockhist[s][.][b] -> exclusive prefix over the chunks, counts[s][b] = bucket population
//   (vmsm_scan_offsets / _flat: counts -> CSR offsets, as for the atomic sort)
//   vmsm_bsort_scatter  cursor[b] = offsets[s][b] + blockhist[s][ch][b] in shared memory; entries written to idx
// No global atomics; the codes (2 B per digit, window-major) are read with 16-byte loads.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, BLOCK <= 256 ? 8 : 1)
    vmsm_bsort_hist(const uint16_t *__restrict__ dig, uint32_t stride, MsmGeom g, uint32_t C, uint32_t per,
                    uint32_t *__restrict__ blockhist) {
    extern __shared__ uint32_t bs_smem[];
    const uint32_t s = blockIdx.x / C, ch = blockIdx.x - s * C;
    for (uint32_t i = threadIdx.x; i < g.NB; i += BLOCK) bs_smem[i] = 0;
    __syncthreads();
    const uint32_t lo = ch * per, hi = g.n - lo < per ? g.n : lo + per;
    if (lo < g.n) {
        const uint32_t vecs = (hi - lo + 7) >> 3, nw = (g.W - s + g.S - 1) / g.S;
        for (uint32_t t = threadIdx.x; t < nw * vecs; t += BLOCK) {
            const uint32_t wi = t / vecs, v = t - wi * vecs, i = lo + 8 * v;
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(dig + (size_t)(s + wi * g.S) * stride + i));
            const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t code = (wd[j >> 1] >> (16 * (j & 1))) & 0xffffu;
                if (code != 0xffffu && i + j < hi) atomicAdd(&bs_smem[code & 0x7fffu], 1u);
            }
        }
    }
    __syncthreads();
    uint32_t *out = blockhist + (size_t)blockIdx.x * g.NB;
    for (uint32_t i = threadIdx.x; i < g.NB; i += BLOCK) out[i] = bs_smem[i];
}

__global__ void __launch_bounds__(256) vmsm_bsort_colscan(uint32_t *__restrict__ blockhist, uint32_t C, uint32_t NB,
                                                          uint32_t nbuckets, uint32_t *__restrict__ counts) {
    const uint32_t t = blockIdx.x * 256u + threadIdx.x;
    if (t >= nbuckets) return;
    const uint32_t s = t / NB, b = t - s * NB;
    uint32_t *p = blockhist + (size_t)s * C * NB + b;
    uint32_t run = 0;
    for (uint32_t ch = 0; ch < C; ch++, p += NB) {
        const uint32_t v = *p;
        *p = run;
        run += v;
    }
    counts[t] = run;
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, BLOCK <= 256 ? 8 : 1)
    vmsm_bsort_scatter(const uint16_t *__restrict__ dig, uint32_t stride, MsmGeom g, uint32_t C, uint32_t per,
                       const uint32_t *__restrict__ blockhist, const uint32_t *__restrict__ offsets,
                       uint32_t *__restrict__ idx) {
    extern __shared__ uint32_t bs_smem[];
    const uint32_t s = blockIdx.x / C, ch = blockIdx.x - s * C;
    const uint32_t lo = ch * per, hi = g.n - lo < per ? g.n : lo + per;
    if (lo >= g.n) return;
    const uint32_t *bh = blockhist + (size_t)blockIdx.x * g.NB, *off = offsets + (size_t)s * g.NB;
    for (uint32_t i = threadIdx.x; i < g.NB; i += BLOCK) bs_smem[i] = off[i] + bh[i];
    __syncthreads();
    const uint32_t vecs = (hi - lo + 7) >> 3, nw = (g.W - s + g.S - 1) / g.S;
    for (uint32_t t = threadIdx.x; t < nw * vecs; t += BLOCK) {
        const uint32_t wi = t / vecs, v = t - wi * vecs, i = lo + 8 * v;
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(dig + (size_t)(s + wi * g.S) * stride + i));
        const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
        const uint32_t kbits = wi << g.lg;  // windows sharing a bucket set: which of them (MsmGeom)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t code = (wd[j >> 1] >> (16 * (j & 1))) & 0xffffu;
            if (code != 0xffffu && i + j < hi) {
                const uint32_t pos = atomicAdd(&bs_smem[code & 0x7fffu], 1u);
                idx[pos] = (i + j) | kbits | ((code & 0x8000u) << 16);
            }
        }
    }
}

// Exclusive prefix sums of n 32-bit lengths into 64-bit offsets (transcript text compaction): per-block sums,
// one block scanning the block sums, then a per-block scan with the block's base.  Returns the total in totals[nblk].
__global__ void __launch_bounds__(1024) vmsm_lens_block_sums(const uint32_t *__restrict__ lens, uint32_t n,
                                                             uint64_t *__restrict__ sums) {
    __shared__ uint32_t ws[32];
    __shared__ uint32_t total;
    uint32_t i = blockIdx.x * 1024u + threadIdx.x;
    block_scan_1024(i < n ? lens[i] : 0u, ws, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
__global__ void vmsm_lens_scan_sums(uint64_t *sums, uint32_t nblk) {
    if (threadIdx.x || blockIdx.x) return;
    uint64_t run = 0;
    for (uint32_t b = 0; b < nblk; b++) {
        uint64_t v = sums[b];
        sums[b] = run;
        run += v;
    }
    sums[nblk] = run;
}
__global__ void __launch_bounds__(1024) vmsm_lens_offsets(const uint32_t *__restrict__ lens, uint32_t n,
                                                          const uint64_t *__restrict__ sums,
                                                          uint64_t *__restrict__ offsets) {
    __shared__ uint32_t ws[32];
    __shared__ uint32_t total;
    uint32_t i = blockIdx.x * 1024u + threadIdx.x;
    uint32_t excl = block_scan_1024(i < n ? lens[i] : 0u, ws, &total);
    if (i < n) offsets[i] = sums[blockIdx.x] + excl;
}

// Counting sort of bucket ids by population, largest first (keys clamped to ORDER_BINS-1).
#define ORDER_BINS 1024
#define ORDER_TILE 2048  // buckets per block (256 threads x 8)
__global__ void __launch_bounds__(256) vmsm_order_hist(const uint32_t *__restrict__ counts, uint32_t nb,
                                                       uint32_t *__restrict__ bins) {
    __shared__ uint32_t h[ORDER_BINS];
    for (uint32_t i = threadIdx.x; i < ORDER_BINS; i += 256) h[i] = 0;
    __syncthreads();
    uint32_t base = blockIdx.x * ORDER_TILE;
    for (uint32_t k = threadIdx.x; k < ORDER_TILE; k += 256) {
        uint32_t b = base + k;
        if (b < nb) {
            uint32_t c = counts[b];
            atomicAdd(&h[c < ORDER_BINS ? c : ORDER_BINS - 1], 1u);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < ORDER_BINS; i += 256)
        if (h[i]) atomicAdd(&bins[i], h[i]);
}
// bins[k] <- number of buckets with key > k (start of key k in descending order)
__global__ void __launch_bounds__(ORDER_BINS) vmsm_order_scan(uint32_t *bins) {
    __shared__ uint32_t s[ORDER_BINS];
    uint32_t t = threadIdx.x;
    uint32_t r = ORDER_BINS - 1 - t;  // reversed index: position t in descending order holds key r
    uint32_t v = bins[r];
    s[t] = v;
    __syncthreads();
    for (uint32_t d = 1; d < ORDER_BINS; d <<= 1) {
        uint32_t add = t >= d ? s[t - d] : 0;
        __syncthreads();
        s[t] += add;
        __syncthreads();
    }
    bins[r] = s[t] - v;
}
__global__ void __launch_bounds__(256) vmsm_order_scatter(const uint32_t *__restrict__ counts, uint32_t nb,
                                                          uint32_t *__restrict__ bins, uint32_t *__restrict__ order) {
    __shared__ uint32_t h[ORDER_BINS];
    __shared__ uint32_t start[ORDER_BINS];
    for (uint32_t i = threadIdx.x; i < ORDER_BINS; i += 256) h[i] = 0;
    __syncthreads();
    uint32_t base = blockIdx.x * ORDER_TILE;
    uint32_t key[ORDER_TILE / 256], rank[ORDER_TILE / 256];
#pragma unroll
    for (uint32_t j = 0; j < ORDER_TILE / 256; j++) {
        uint32_t b = base + j * 256 + threadIdx.x;
        key[j] = 0xffffffffu;
        if (b < nb) {
            uint32_t c = counts[b];
            key[j] = c < ORDER_BINS ? c : ORDER_BINS - 1;
            rank[j] = atomicAdd(&h[key[j]], 1u);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < ORDER_BINS; i += 256)
        if (h[i]) start[i] = atomicAdd(&bins[i], h[i]);
    __syncthreads();
#pragma unroll
    for (uint32_t j = 0; j < ORDER_TILE / 256; j++) {
        uint32_t b = base + j * 256 + threadIdx.x;
        if (key[j] != 0xffffffffu) order[start[key[j]] + rank[j]] = b;
    }
}

// Integer-pipe peak for limb products: carry-chained 32x32+64 multiply-accumulates exactly as a multi-precision
// multiplication issues them (IMAD.WIDE.U32 / IMAD.WIDE.U32.X with predicate carry), two independent chains of four
// per thread and iteration.  Every IMAD.WIDE form measured on B200 issues at 32 lanes/clk/SM, half the 32-bit IMAD
// rate (tools/microbench.cu, profiles/r01/s3_microbench_valid.jsonl); this form reaches that limit.
#define MB_ITERS 2048
__global__ void __launch_bounds__(512) vmsm_imad_peak(uint32_t *out, uint32_t a, uint32_t b) {
    uint32_t r[2][9];
#pragma unroll
    for (int k = 0; k < 2; k++)
#pragma unroll
        for (int i = 0; i < 9; i++) r[k][i] = threadIdx.x + i + k;
    uint32_t x0 = a + threadIdx.x, x1 = x0 * 3, x2 = x0 * 5, x3 = x0 * 7, y = b;
    for (int it = 0; it < MB_ITERS; it++) {
#pragma unroll
        for (int k = 0; k < 2; k++)
            fe_mad4(r[k][0], r[k][1], r[k][2], r[k][3], r[k][4], r[k][5], r[k][6], r[k][7], r[k][8], x0, x1, x2, x3, y);
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 2; k++)
#pragma unroll
        for (int i = 0; i < 9; i++) s ^= r[k][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ------------------------------------------------------------------------------------------------ context
namespace {

struct PointSet {
    int32_t curve;
    uint64_t n;
    ge_aff *aff;      // Ed25519: canonical affine (wire form)
    ge_niels *niels;  // Ed25519: (y+x, y-x, 2dxy)
    void *w_wire;     // BN256: plain canonical affine (wire form), waff<F>[n]; identity = all zero
    void *w_base;     // BN256: Montgomery affine, waff<F>[n]
    // optional table of 2^(pre_c * w) * P_i, w < pre_W, pre_n points per level (vmsm_points_precompute)
    void *pre = nullptr;
    uint32_t pre_c = 0, pre_W = 0;
    uint64_t pre_n = 0;
};

inline size_t wire_bytes(int32_t curve) { return curve == VMSM_CURVE_BN256_G2 ? 128 : 64; }
struct ScalarSet {
    uint64_t n;
    uint32_t *data;
};
struct EventSet {
    cudaEvent_t ev[PH_COUNT + 1];
};

constexpr uint32_t kSlots = 64;
constexpr uint32_t kLaRing = 8;  // asynchronous small linear combinations in flight
constexpr uint32_t kDotT1 = 1u << 16, kDotT2 = 256;  // most partial-sum threads of vmsm_scalars_dot
inline void dot_stage_sizes(uint64_t n, uint32_t *T, uint32_t *T2) {
    uint64_t t = n / 16;
    if (t < 64) t = n < 64 ? n : 64;
    if (t > kDotT1) t = kDotT1;
    uint32_t t2 = 1;
    while ((uint64_t)t2 * t2 < t && t2 < kDotT2) t2 <<= 1;
    *T = (uint32_t)t;
    *T2 = t2 < t ? t2 : (uint32_t)t;
}

struct Ctx {
    std::recursive_mutex mu;  // serialises the entry points that use this context (GET_CTX)
    int device = 0;
    cudaStream_t stream = nullptr;  // main stream: everything except MSM tails
    cudaStream_t tails[kTailWays] = {};  // side streams: upper bucket-tree levels + Horner of the previous MSMs
    cudaStream_t sort = nullptr;    // side stream: counting sort (digits, scan, scatter, order) of the NEXT MSM
    // two HEAD streams taken in turn by consecutive MSMs for their accumulate kernel, so the ramp-down of one overlaps
    // the ramp-up of the next (below ~2^18 terms a launch is one or two waves of threads that finish together); all
    // other work of the context stays ordered on `stream`
    cudaStream_t heads[2] = {nullptr, nullptr};
    cudaEvent_t ev_issue = nullptr, ev_head_done[2] = {nullptr, nullptr};
    bool head_pending[2] = {false, false};
    bool dual_head = true;
    cudaEvent_t ev_sorted[2] = {nullptr, nullptr}, ev_acc_done[2] = {nullptr, nullptr}, ev_sort_in = nullptr;
    bool acc_pending[2] = {false, false};
    bool async_sort = true;
    uint32_t sort_blocks = 0;
    uint32_t sms = 148;
    bool sort_blocks_auto = true;  // until VMSM_OPT_SORT_BLOCKS names a count
    uint32_t fold_quad_max = 1u << 13;  // measured crossover (profiles/r01/fold_kernel_quad.md)
    cudaEvent_t scalars_ready = nullptr;  // set by the entry point when the scalars of the next MSM are still in flight
    cudaEvent_t ev_head = nullptr, ev_tail[kTailWays] = {};
    bool tail_pending[kTailWays] = {};
    bool async_tail = true;
    uint32_t msm_seq = 0;
    // asynchronous end-to-end path (vmsm_msm_async): H2D of the scalars on a copy stream into one of two staging
    // buffers, overlapping the previous MSM; results land in host-mapped pinned memory, one event per result slot
    cudaStream_t copy = nullptr;
    uint32_t *astage[2] = {nullptr, nullptr};
    size_t astage_cap[2] = {0, 0};
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr}, ev_dot[2] = {nullptr, nullptr};
    bool astage_used[2] = {false, false};
    // vmsm_lincomb_async: a ring of private staging sets (64 points + 64 scalars + error word each)
    ge_aff *la_aff = nullptr;
    ge_niels *la_niels = nullptr;
    uint32_t *la_scalars = nullptr, *la_err = nullptr;
    cudaEvent_t ev_la_sorted[kLaRing] = {};
    bool la_used[kLaRing] = {};
    uint32_t la_seq = 0;
    // freed point / scalar vectors are kept for reuse: cudaFree synchronises the whole device and was measured at up to
    // 0.18 s per call inside a proof (a prover frees its private copy of g_hat at the end of every proof)
    std::multimap<size_t, void *> pool;
    std::unordered_map<void *, size_t> pool_size;
    size_t pool_bytes = 0;
    // device-resident scalar vectors (vmsm_scalars_fold / _dot): written on the main stream, read by the copy stream
    cudaEvent_t ev_sc_written = nullptr;
    bool sc_dirty = false;
    uint32_t *dot_scratch = nullptr;  // kDotT1 + kDotT2 + 1 partial sums
    uint32_t async_seq = 0;
    cudaEvent_t ev_slot[64] = {nullptr};
    uint32_t cur_slot = 0;
    ge_aff *res_aff_host = nullptr;  // pinned + mapped, kSlots entries: written by the final kernel itself
    // host-side normalisation (default): the final kernel stores the EXTENDED result here and fetch_slot inverts Z on
    // the CPU -- the device inversion is a 0.19 ms dependent chain at the end of every MSM, a CPU core needs ~15 us
    ge_ext *res_xyz_host = nullptr;  // pinned + mapped, kSlots entries
    bool host_norm = true;
    bool slot_host_norm[64] = {false};
    // BN256 results: Jacobian on the device, plain affine in mapped pinned memory; which curve a slot last held
    uint8_t *res_w_dev = nullptr;   // kSlots x 192 B
    uint8_t *res_w_host = nullptr;  // kSlots x 128 B, mapped
    uint8_t *res_wj_host = nullptr;  // kSlots x 192 B, mapped: Jacobian results for host-side normalisation
    int32_t slot_curve[64] = {0};
    void *fbw_table[2] = {nullptr, nullptr};  // fixed-base tables for G1 / G2 (520 entries), built on first use
    void *small_w_wire = nullptr, *small_w_base = nullptr;  // lincomb scratch, 64 x 128 B each
    // transcript text scratch (vmsm_points_text): grow-only device buffers and a pinned host buffer
    uint8_t *txt_slots = nullptr, *txt_text = nullptr, *txt_host = nullptr;
    uint8_t *txt_host_sc = nullptr;  // scalar text has its own pinned buffer: a point view and a scalar view coexist
    uint32_t *txt_lens = nullptr;
    uint64_t *txt_offsets = nullptr, *txt_sums = nullptr;
    size_t txt_cap = 0;
    uint32_t *res_status_host = nullptr;  // pinned + mapped, kSlots words: 0 ok, 1 = a multi-GPU partial timed out
    // multi-GPU mailbox (kernels.cuh: KPushPartial / KGatherPartials)
    MailSlot *mailbox = nullptr;  // [kSlots][mb_world]; owner: local allocation, others: peer mapping
    bool mb_owner = false, mb_ipc = false;
    uint32_t mb_world = 0, mb_rank = 0;
    uint32_t shard_seq = 0;  // != 0 while a sharded MSM is being issued
    uint32_t shard_next = 0;  // VMSM_OPT_SHARD_SEQ: applies to the next MSM call of any flavour, then clears
    MsmOptions opt;
    // block-privatised counting sort: digit codes (W x stride x 2 B) and per-block histograms (S x C x NB x 4 B),
    // double-buffered by MSM parity like the CSR lists
    uint16_t *bs_dig[2] = {nullptr, nullptr};
    uint32_t *bs_hist[2] = {nullptr, nullptr};
    size_t bs_dig_cap[2] = {0, 0}, bs_hist_cap[2] = {0, 0};
    uint32_t bs_min_terms = 1u << 15;  // smaller MSMs are latency-bound: three short kernels beat five
    bool bs_attr_set = false;
    uint32_t seg_resident_w[2] = {0, 0};  // same for the BN256 kernels (G1 / G2)
    uint32_t seg_resident[2] = {0, 0};  // resident threads of the segmented accumulate kernels (plain / tables), cached
    uint64_t pre_min_terms = 256;  // MSMs shorter than this ignore a precomputed table (VMSM_OPT_PRE_MIN_TERMS)
    bool phase_timing = false;
    bool check_points = true;
    Workspace ws;
    uint32_t *order_bins = nullptr;
    uint32_t *err_word = nullptr;   // device
    ge_niels *fb_table = nullptr;   // device 64 x 8
    ge_ext *res_ext = nullptr;      // kSlots
    ge_aff *res_aff = nullptr;      // kSlots
    uint32_t *stage_scalars = nullptr;  // device staging for vmsm_msm / vmsm_lincomb
    size_t stage_cap = 0;
    ge_ext *tmp_ext = nullptr;  // fold / fixed-base scratch
    size_t tmp_cap = 0;
    ge_aff *small_aff = nullptr;  // lincomb scratch (64 points)
    ge_niels *small_niels = nullptr;
    uint8_t *pin = nullptr;  // pinned host bounce buffer (results, error word)
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    std::map<uint64_t, PointSet> points;
    std::map<uint64_t, ScalarSet> scalars;
    uint64_t next_id = 1;
    uint64_t launches = 0;
    // phase timing
    std::vector<EventSet> ev_pool;
    size_t ev_used = 0;
    int ev_cur = -1;
    double phase_ms[PH_COUNT] = {0};
    uint64_t phase_calls = 0;
};

std::mutex g_mu;
std::set<Ctx *> g_ctxs;

Ctx *get_ctx(uint64_t h) {
    std::lock_guard<std::mutex> lk(g_mu);
    Ctx *c = reinterpret_cast<Ctx *>(h);
    return g_ctxs.count(c) ? c : nullptr;
}

struct CudaBE {
    Ctx *c;
    cudaError_t err = cudaSuccess;
    cudaStream_t cur = nullptr;  // stream the next launch goes to (main unless inside a tail)
    cudaStream_t head = nullptr;  // stream of this MSM's accumulate kernel: one of c->heads, or the main stream
    explicit CudaBE(Ctx *ctx) : c(ctx), cur(ctx->stream), head(ctx->stream) {}
    // consecutive MSMs alternate between the two head streams; each starts after everything issued on the main stream
    void use_head(uint32_t seq) {
        if (!c->dual_head) return;
        const int h = (int)(seq & 1);
        head = c->heads[h];
        note(cudaEventRecord(c->ev_issue, c->stream));
        note(cudaStreamWaitEvent(head, c->ev_issue, 0));
        c->head_pending[h] = true;
        cur = head;
    }
    void head_done(uint32_t seq) {
        if (c->dual_head) note(cudaEventRecord(c->ev_head_done[seq & 1], head));
    }
    uint32_t overflow_warps() { return 148u * 8u * 4u; }  // 8 blocks of 4 warps per SM, grid-stride over the tasks
    uint32_t combine_threads() { return 148u * 128u; }
    // the node buffers of parity `par` may still be read by the tail of the MSM two calls ago
    bool thin_sort = false;
    void head_wait_tail(int par) {
        if (c->tail_pending[par]) note(cudaStreamWaitEvent(head, c->ev_tail[par], 0));
    }
    // counting sort on its own stream: ordered after whatever produced the scalars on the main stream (H2D, synth)
    // and after the accumulate kernel that last read this parity's CSR lists
    void sort_begin(int par) {
        if (!c->async_sort) return;
        // order the sort after the copy that produces its scalars -- NOT after the main stream as a whole, or it
        // would wait for the previous MSM's accumulate kernel, which is exactly what it is meant to run under
        if (c->scalars_ready) note(cudaStreamWaitEvent(c->sort, c->scalars_ready, 0));
        c->scalars_ready = nullptr;
        if (c->acc_pending[par]) note(cudaStreamWaitEvent(c->sort, c->ev_acc_done[par], 0));
        // is the previous MSM still accumulating?  Then this sort runs underneath it and should stay thin; a lone MSM
        // (or the first of a burst) gets the whole machine
        thin_sort = c->acc_pending[par ^ 1] && cudaEventQuery(c->ev_acc_done[par ^ 1]) == cudaErrorNotReady;
        cur = c->sort;
    }
    void sort_end(int par) {
        if (!c->async_sort) return;
        note(cudaEventRecord(c->ev_sorted[par], c->sort));
        note(cudaStreamWaitEvent(head, c->ev_sorted[par], 0));
        cur = head;
    }
    void acc_done(int par) {  // recorded whatever the sort mode: vmsm_fold / vmsm_points_free order themselves on it
        note(cudaEventRecord(c->ev_acc_done[par], cur));  // on the stream of the last reader of the CSR lists
        c->acc_pending[par] = true;
    }
    void after_final(ge_ext *out_ext, ge_aff *out_aff) {
        if (!c->shard_seq || !c->mailbox) return;
        MailSlot *box = c->mailbox + (size_t)c->cur_slot * c->mb_world;
        KPushPartial kp = {out_ext, box + c->mb_rank, c->shard_seq};
        launch(kp, 32);
        if (c->mb_owner) {
            KGatherPartials kg = {box, c->mb_world, c->shard_seq, out_ext, out_aff, c->res_status_host + c->cur_slot,
                                  c->slot_host_norm[c->cur_slot] ? c->res_xyz_host + c->cur_slot : nullptr};
            launch(kg, 32);
        }
    }
    void result_ready() { note(cudaEventRecord(c->ev_slot[c->cur_slot], cur)); }
    void tail_begin(int way) {
        if (!c->async_tail) return;
        note(cudaEventRecord(c->ev_head, head));
        note(cudaStreamWaitEvent(c->tails[way], c->ev_head, 0));
        cur = c->tails[way];
    }
    void tail_end(int way) {
        if (!c->async_tail) return;
        note(cudaEventRecord(c->ev_tail[way], c->tails[way]));
        c->tail_pending[way] = true;
        cur = head;
    }
    void note(cudaError_t e) {
        if (err == cudaSuccess && e != cudaSuccess) err = e;
    }
    void *alloc(size_t bytes) {
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
        if (e != cudaSuccess) {
            note(e);
            return nullptr;
        }
        return p;
    }
    void free(void *p) {
        if (p) cudaFree(p);
    }
    void zero(void *p, size_t bytes) { note(cudaMemsetAsync(p, 0, bytes, cur)); }
    template <class F>
    void launch(const F &f, uint32_t n) {
        if (!n) return;
        uint32_t grid = (n + F::kBlock - 1) / F::kBlock;
        vmsm_kernel<F><<<grid, F::kBlock, dyn_smem<F>::value, cur>>>(f, n);
        c->launches++;
        note(cudaGetLastError());
    }
    // per-scalar kernels of the counting sort: on the sort stream they run as a thin grid-stride slice of every SM
    template <class F>
    void launch_sort(const F &f, uint32_t n) {
        uint32_t grid = (n + F::kBlock - 1) / F::kBlock;
        // thin grid: four blocks per SM below 2^18 terms, two up to 2^20, one from 2^21 (round 2,
        // profiles/r02/sort_blocks_sweep.jsonl: with the accumulate kernels on two head streams and single-wave launches
        // at 65 % of the resident threads, one block per SM -- the round-1 optimum at every size -- is 1-5 % slower from
        // 2^16 to 2^20 and still 3-6 % faster at 2^21 / 2^22)
        const uint32_t sb = c->sort_blocks_auto ? (n < (1u << 18) ? 4u : n < (1u << 21) ? 2u : 1u) * c->sms : c->sort_blocks;
        if (!c->async_sort || !thin_sort || !sb || grid <= sb) return launch(f, n);
        vmsm_kernel_strided<F><<<sb, F::kBlock, 0, cur>>>(f, n);
        c->launches++;
        note(cudaGetLastError());
    }
    void scan_offsets(const uint32_t *counts, uint32_t *offsets, uint32_t *cursor, const MsmGeom &g) {
        uint32_t tiles = (g.NB + SCAN_TILE - 1) / SCAN_TILE;
        vmsm_scan_offsets<<<g.S * tiles, SCAN_TILE, 0, cur>>>(counts, offsets, cursor, g, tiles);  // one row per bucket set
        c->launches++;
        note(cudaGetLastError());
    }
    void scan_offsets_flat(const uint32_t *counts, uint32_t *offsets, uint32_t *cursor, const MsmGeom &g,
                           uint32_t *row_totals, uint32_t *seg_bucket, uint32_t *total, uint32_t L) {
        uint32_t tiles = (g.NB + SCAN_TILE - 1) / SCAN_TILE;
        vmsm_row_totals<<<g.S, 1024, 0, cur>>>(counts, g.NB, row_totals);
        vmsm_scan_offsets_flat<<<g.S * tiles, SCAN_TILE, 0, cur>>>(counts, offsets, cursor, g, tiles, row_totals, seg_bucket,
                                                                   total, L);
        c->launches += 2;
        note(cudaGetLastError());
    }
    // ---- block-privatised counting sort (vmsm_bsort_*): number of scalar chunks, or 0 = not for this geometry
    uint32_t bsort_chunks(const MsmGeom &g) {
        if (g.n < c->bs_min_terms || g.c > 16 || g.NB * 4u > 160u * 1024u || g.S > 148u) return 0;
        uint32_t C = 148u / g.S;  // about one block per SM
        return C ? C : 1u;
    }
    static uint32_t bsort_stride(const MsmGeom &g) { return (g.n + 7u) & ~7u; }
    static uint32_t bsort_per(const MsmGeom &g, uint32_t C) { return (((g.n + C - 1) / C) + 7u) & ~7u; }
    int bsort_ensure(const MsmGeom &g, uint32_t C, int par) {
        if (!c->bs_attr_set) {
            const int big = 160 * 1024;
            note(cudaFuncSetAttribute(vmsm_bsort_hist<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
            note(cudaFuncSetAttribute(vmsm_bsort_hist<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
            note(cudaFuncSetAttribute(vmsm_bsort_scatter<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
            note(cudaFuncSetAttribute(vmsm_bsort_scatter<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
            c->bs_attr_set = true;
        }
        const size_t need_dig = (size_t)g.W * bsort_stride(g) * 2, need_hist = (size_t)g.S * C * g.NB * 4;
        if (need_dig > c->bs_dig_cap[par]) {
            if (c->bs_dig[par]) cudaFree(c->bs_dig[par]);
            c->bs_dig_cap[par] = 0;
            c->bs_dig[par] = (uint16_t *)alloc(need_dig);
            if (!c->bs_dig[par]) return -1;
            c->bs_dig_cap[par] = need_dig;
        }
        if (need_hist > c->bs_hist_cap[par]) {
            if (c->bs_hist[par]) cudaFree(c->bs_hist[par]);
            c->bs_hist_cap[par] = 0;
            c->bs_hist[par] = (uint32_t *)alloc(need_hist);
            if (!c->bs_hist[par]) return -1;
            c->bs_hist_cap[par] = need_hist;
        }
        return 0;
    }
    void bsort_hist(const uint32_t *scalars, const MsmGeom &g, uint32_t C, int par, uint32_t *counts) {
        const uint32_t stride = bsort_stride(g), per = bsort_per(g, C), nb = g.S * g.NB;
        KRecode k0 = {scalars, c->bs_dig[par], stride, g};
        launch_sort(k0, g.n);
        // underneath an accumulate kernel: 256 threads at <= 32 registers, what four resident accumulate blocks leave
        // free on an SM; alone: 1024 threads per block
        if (thin_sort) vmsm_bsort_hist<256><<<g.S * C, 256, g.NB * 4, cur>>>(c->bs_dig[par], stride, g, C, per, c->bs_hist[par]);
        else vmsm_bsort_hist<1024><<<g.S * C, 1024, g.NB * 4, cur>>>(c->bs_dig[par], stride, g, C, per, c->bs_hist[par]);
        vmsm_bsort_colscan<<<(nb + 255) / 256, 256, 0, cur>>>(c->bs_hist[par], C, g.NB, nb, counts);
        c->launches += 2;
        note(cudaGetLastError());
    }
    void bsort_scatter(const MsmGeom &g, uint32_t C, int par, const uint32_t *offsets, uint32_t *idx) {
        const uint32_t stride = bsort_stride(g), per = bsort_per(g, C);
        if (thin_sort)
            vmsm_bsort_scatter<256><<<g.S * C, 256, g.NB * 4, cur>>>(c->bs_dig[par], stride, g, C, per, c->bs_hist[par], offsets, idx);
        else
            vmsm_bsort_scatter<1024><<<g.S * C, 1024, g.NB * 4, cur>>>(c->bs_dig[par], stride, g, C, per, c->bs_hist[par], offsets, idx);
        c->launches++;
        note(cudaGetLastError());
    }
    // threads of the segmented accumulate kernel one wave holds: SMs x resident blocks x block size
    uint32_t resident_threads(bool pre) {
        uint32_t &r = c->seg_resident[pre ? 1 : 0];
        if (!r) {
            int blocks = 0, sms = 0;
            if (pre) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, vmsm_kernel<KAccumulateSegPre>, KAccumulateSegPre::kBlock, KAccumulateSegPre::kDynSmem);
            else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, vmsm_kernel<KAccumulateSeg>, KAccumulateSeg::kBlock, KAccumulateSeg::kDynSmem);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
            r = (uint32_t)(blocks > 0 ? blocks : 4) * (uint32_t)(sms > 0 ? sms : 148) * KAccumulateSeg::kBlock;
        }
        return r;
    }
    template <class F>
    uint32_t resident_threads_w() {
        uint32_t &r = c->seg_resident_w[sizeof(typename F::T) > 32 ? 1 : 0];
        if (!r) {
            int blocks = 0, sms = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, vmsm_kernel<KAccumulateSegW<F>>, KAccumulateSegW<F>::kBlock, 0);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
            r = (uint32_t)(blocks > 0 ? blocks : 2) * (uint32_t)(sms > 0 ? sms : 148) * KAccumulateSegW<F>::kBlock;
        }
        return r;
    }
    bool order_buckets(const uint32_t *counts, uint32_t *order, uint32_t nb, uint32_t n) {
        if (nb < 8192 || n == 0) return false;  // too few buckets for ordering to matter
        note(cudaMemsetAsync(c->order_bins, 0, ORDER_BINS * 4, cur));
        uint32_t grid = (nb + ORDER_TILE - 1) / ORDER_TILE;
        vmsm_order_hist<<<grid, 256, 0, cur>>>(counts, nb, c->order_bins);
        vmsm_order_scan<<<1, ORDER_BINS, 0, cur>>>(c->order_bins);
        vmsm_order_scatter<<<grid, 256, 0, cur>>>(counts, nb, c->order_bins, order);
        c->launches += 3;
        note(cudaGetLastError());
        return true;
    }
    // phase timing
    void phase_begin() {
        c->ev_cur = -1;
        if (!c->phase_timing) return;
        if (c->ev_used == c->ev_pool.size()) {
            EventSet es;
            for (auto &e : es.ev) note(cudaEventCreate(&e));
            c->ev_pool.push_back(es);
        }
        c->ev_cur = (int)c->ev_used++;
        note(cudaEventRecord(c->ev_pool[c->ev_cur].ev[0], cur));
    }
    void phase_mark(int ph) {
        if (c->ev_cur >= 0) note(cudaEventRecord(c->ev_pool[c->ev_cur].ev[ph + 1], cur));
    }
    void phase_end() {}
};

// The last readers of an MSM's bases (and CSR lists) are its accumulate kernel on the main stream and, for long
// buckets, its overflow kernel on a tail stream; `ev_acc_done` is recorded right after whichever comes last.  Anything
// that rewrites bases in place (vmsm_fold) or hands their memory to the pool (vmsm_points_free) goes through here.
cudaError_t wait_bases_released(Ctx *c, bool host_side) {
    for (int k = 0; k < 2; k++)
        if (c->acc_pending[k]) {
            cudaError_t e = host_side ? cudaEventSynchronize(c->ev_acc_done[k]) : cudaStreamWaitEvent(c->stream, c->ev_acc_done[k], 0);
            if (e != cudaSuccess) return e;
        }
    return cudaSuccess;
}

// make the main stream wait for every MSM tail issued so far (tails are ordered on the side stream)
cudaError_t join_tail(Ctx *c) {
    for (int h = 0; h < 2; h++)
        if (c->head_pending[h]) {
            cudaError_t e = cudaStreamWaitEvent(c->stream, c->ev_head_done[h], 0);
            if (e != cudaSuccess) return e;
        }
    for (int w = 0; w < kTailWays; w++)
        if (c->tail_pending[w]) {
            cudaError_t e = cudaStreamWaitEvent(c->stream, c->ev_tail[w], 0);
            if (e != cudaSuccess) return e;
        }
    return cudaSuccess;
}

int32_t harvest_phases(Ctx *c) {
    if (!c->ev_used) return VMSM_OK;
    CU(join_tail(c));
    CU(cudaStreamSynchronize(c->stream));
    for (size_t k = 0; k < c->ev_used; k++) {
        for (int p = 0; p < PH_COUNT; p++) {
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, c->ev_pool[k].ev[p], c->ev_pool[k].ev[p + 1]));
            c->phase_ms[p] += ms;
        }
        c->phase_calls++;
    }
    c->ev_used = 0;
    return VMSM_OK;
}

int32_t ensure_stage(Ctx *c, size_t n_scalars) {
    if (n_scalars <= c->stage_cap) return VMSM_OK;
    if (c->stage_scalars) cudaFree(c->stage_scalars);
    c->stage_scalars = nullptr;
    c->stage_cap = 0;
    CU(cudaMalloc(&c->stage_scalars, n_scalars * 32));
    c->stage_cap = n_scalars;
    return VMSM_OK;
}
// ---- cached device allocations for point / scalar vectors (context of the calling entry point)
static thread_local Ctx *tl_ctx = nullptr;
constexpr size_t kPoolCapBytes = 16ull << 30;

static size_t pool_round(size_t bytes) {
    if (bytes < 512) bytes = 512;
    if (bytes <= (1u << 20)) {
        size_t p = 512;
        while (p < bytes) p <<= 1;
        return p;
    }
    return (bytes + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);
}

template <class T>
static cudaError_t pool_alloc(T **p, size_t bytes) {
    Ctx *c = tl_ctx;
    bytes = pool_round(bytes);
    auto it = c->pool.lower_bound(bytes);
    if (it != c->pool.end() && it->first <= bytes + bytes / 4) {
        *p = (T *)it->second;
        c->pool_bytes -= it->first;
        c->pool.erase(it);
        return cudaSuccess;
    }
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e != cudaSuccess && !c->pool.empty()) {  // out of memory with blocks parked in the pool: give them back, retry
        for (auto &kv : c->pool) cudaFree(kv.second), c->pool_size.erase(kv.second);
        c->pool.clear();
        c->pool_bytes = 0;
        cudaGetLastError();
        e = cudaMalloc(&q, bytes);
    }
    if (e != cudaSuccess) return e;
    c->pool_size[q] = bytes;
    *p = (T *)q;
    return cudaSuccess;
}

// the caller has already made sure no stream still uses the block
static void pool_free(void *p) {
    if (!p) return;
    Ctx *c = tl_ctx;
    auto it = c->pool_size.find(p);
    if (it == c->pool_size.end() || c->pool_bytes + it->second > kPoolCapBytes) {
        if (it != c->pool_size.end()) c->pool_size.erase(it);
        cudaFree(p);
        return;
    }
    c->pool.insert({it->second, p});
    c->pool_bytes += it->second;
}

int32_t ensure_tmp(Ctx *c, size_t n_pts) {
    if (n_pts <= c->tmp_cap) return VMSM_OK;
    if (c->tmp_ext) cudaFree(c->tmp_ext);
    c->tmp_ext = nullptr;
    c->tmp_cap = 0;
    CU(cudaMalloc(&c->tmp_ext, n_pts * sizeof(ge_ext)));
    c->tmp_cap = n_pts;
    return VMSM_OK;
}

int32_t run_msm(Ctx *c, const ge_niels *bases, const uint32_t *scalars, uint64_t n, uint32_t slot,
                const ge_niels *extra = nullptr, uint32_t n_extra = 0, const PreTable *pre = nullptr) {
    if (n > (1ull << 26)) return fail(VMSM_ERR_UNSUPPORTED, "n = %llu exceeds 2^26 terms per MSM call", (unsigned long long)n);
    CudaBE be(c);
    c->cur_slot = slot;
    c->slot_curve[slot] = VMSM_CURVE_ED25519;
    if (c->shard_next && slot != kSlots - 1) {
        if (!c->mailbox) return fail(VMSM_ERR_INVALID, "VMSM_OPT_SHARD_SEQ set but the context has no mailbox");
        c->shard_seq = c->shard_next;
        c->shard_next = 0;
    }
    // results leave the device in extended coordinates and the fetching call normalises on the CPU; a sharded MSM too:
    // every rank's final kernel skips the inversion (only the extended partial is pushed), and the owner's gather kernel
    // overwrites the owner's host slot with the extended sum (same stream, after its own final kernel)
    const bool hn = c->host_norm;
    c->slot_host_norm[slot] = hn;
    int rc = msm_run(be, c->ws, c->opt, 253, bases, scalars, (uint32_t)n, c->res_ext + slot, c->res_aff_host + slot,
                     c->msm_seq++, extra, n_extra, pre, hn ? c->res_xyz_host + slot : nullptr);
    c->shard_seq = 0;
    if (rc == -2) return fail(VMSM_ERR_INVALID, "precomputed table does not match the MSM geometry");
    if (rc) return fail(VMSM_ERR_NOMEM, "workspace allocation failed: %s", cudaGetErrorString(be.err));
    if (be.err != cudaSuccess) return fail(VMSM_ERR_CUDA, "msm launch: %s", cudaGetErrorString(be.err));
    return VMSM_OK;
}

// table of 2^(cb * w) * P_i for one point vector (vmsm_points_precompute; built lazily for the extra terms of an MSM)
int32_t precompute_ps(Ctx *c, PointSet &ps, uint32_t cb) {
    if (ps.curve != VMSM_CURVE_ED25519) return fail(VMSM_ERR_UNSUPPORTED, "precomputed tables exist for the Ed25519 path only");
    if (ps.pre && ps.pre_c == cb) return VMSM_OK;
    if (ps.n == 0) return VMSM_OK;
    CU(wait_bases_released(c, true));
    if (ps.pre) pool_free(ps.pre);
    ps.pre = nullptr;
    ps.pre_c = ps.pre_W = 0;
    const uint32_t W = (253 + cb) / cb;
    ge_niels *tbl = nullptr;
    cudaError_t e = pool_alloc(&tbl, (size_t)W * ps.n * sizeof(ge_niels));
    if (e != cudaSuccess)
        return fail(VMSM_ERR_NOMEM, "table of %u levels x %llu points: %s", W, (unsigned long long)ps.n, cudaGetErrorString(e));
    CudaBE be(c);
    KPrecompute k = {ps.aff, tbl, (uint32_t)ps.n, cb, W};
    be.launch(k, (uint32_t)ps.n);
    if (be.err != cudaSuccess) {
        pool_free(tbl);
        return fail(VMSM_ERR_CUDA, "precompute: %s", cudaGetErrorString(be.err));
    }
    ps.pre = tbl;
    ps.pre_c = cb;
    ps.pre_W = W;
    ps.pre_n = ps.n;
    return VMSM_OK;
}

// Ed25519 MSM over terms [off, off + n_total - n_extra) of `ps` plus n_extra terms of `eps`: through the precomputed
// tables when both vectors have one for the same window (vmsm_points_precompute), else the plain windowed path.
int32_t run_msm_ps(Ctx *c, const PointSet &ps, uint64_t off, const uint32_t *scalars, uint64_t n_total, uint32_t slot,
                   PointSet *eps = nullptr, uint64_t eoff = 0, uint32_t n_extra = 0) {
    bool pre_ok = ps.pre && n_total >= c->pre_min_terms;
    if (pre_ok && n_extra && !(eps->pre && eps->pre_c == ps.pre_c)) {
        // the blinding bases (h, k: one-element vectors cached by the caller) get their table on first use
        if (eps->n <= 64 && eps != &ps) {
            int32_t rc = precompute_ps(c, *eps, ps.pre_c);
            if (rc) return rc;
        } else {
            pre_ok = false;
        }
    }
    if (!pre_ok)
        return run_msm(c, ps.niels + off, scalars, n_total, slot, eps ? eps->niels + eoff : nullptr, n_extra);
    PreTable pt = {(uint32_t)ps.pre_n, ps.pre_c, ps.pre_W, n_extra ? (const ge_niels *)eps->pre + eoff : nullptr,
                   n_extra ? (uint32_t)eps->pre_n : 0u};
    return run_msm(c, (const ge_niels *)ps.pre + off, scalars, n_total, slot, nullptr, n_extra, &pt);
}


// ------------------------------------------------------------------------------------------------ BN256 paths
template <class F>
int32_t w_new_pointset(int32_t curve, uint64_t n, PointSet *ps) {
    ps->curve = curve;
    ps->n = n;
    ps->aff = nullptr;
    ps->niels = nullptr;
    ps->w_wire = ps->w_base = nullptr;
    CU(pool_alloc(&ps->w_wire, (n ? n : 1) * sizeof(waff<F>)));
    cudaError_t e = pool_alloc(&ps->w_base, (n ? n : 1) * sizeof(waff<F>));
    if (e != cudaSuccess) {
        pool_free(ps->w_wire);
        return fail(VMSM_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
    }
    return VMSM_OK;
}

template <class F>
int32_t w_upload(Ctx *c, int32_t curve, const uint8_t *wire, uint64_t n, PointSet *ps) {
    int32_t rc = w_new_pointset<F>(curve, n, ps);
    if (rc || !n) return rc;
    CudaBE be(c);
    be.note(cudaMemcpyAsync(ps->w_wire, wire, n * sizeof(waff<F>), cudaMemcpyHostToDevice, c->stream));
    be.zero(c->err_word, 4);
    KUploadW<F> k = {(const waff<F> *)ps->w_wire, (waff<F> *)ps->w_base, c->err_word, c->check_points ? 1u : 0u};
    be.launch(k, (uint32_t)n);
    be.note(cudaMemcpyAsync(c->pin, c->err_word, 4, cudaMemcpyDeviceToHost, c->stream));
    be.note(cudaStreamSynchronize(c->stream));
    uint32_t ew = *reinterpret_cast<uint32_t *>(c->pin);
    if (be.err != cudaSuccess || ew) {
        pool_free(ps->w_wire), pool_free(ps->w_base);
        if (be.err != cudaSuccess) return fail(VMSM_ERR_CUDA, "upload: %s", cudaGetErrorString(be.err));
        return fail(VMSM_ERR_POINT, "invalid point in upload (%s%s)", (ew & 1) ? "non-canonical coordinate " : "",
                    (ew & 2) ? "not on curve" : "");
    }
    return VMSM_OK;
}

template <class F>
int32_t w_table(Ctx *c, int which) {
    if (c->fbw_table[which]) return VMSM_OK;
    std::vector<waff<F>> tbl(520);
    build_fixed_base_table_w<F>(tbl.data());
    CU(cudaMalloc(&c->fbw_table[which], 520 * sizeof(waff<F>)));
    CU(cudaMemcpy(c->fbw_table[which], tbl.data(), 520 * sizeof(waff<F>), cudaMemcpyHostToDevice));
    return VMSM_OK;
}

template <class F>
int32_t w_fixed_base(Ctx *c, int32_t curve, const uint8_t *scalars, uint64_t seed, uint64_t n, PointSet *ps) {
    int32_t rc = w_table<F>(c, curve == VMSM_CURVE_BN256_G2 ? 1 : 0);
    if (rc) return rc;
    rc = w_new_pointset<F>(curve, n, ps);
    if (rc || !n) return rc;
    const uint32_t *dsc = nullptr;
    if (scalars) {
        rc = ensure_stage(c, n);
        if (!rc) {
            cudaError_t e = cudaMemcpyAsync(c->stage_scalars, scalars, n * 32, cudaMemcpyHostToDevice, c->stream);
            if (e != cudaSuccess) rc = fail(VMSM_ERR_CUDA, "H2D: %s", cudaGetErrorString(e));
        }
        dsc = c->stage_scalars;
    }
    if (!rc) rc = ensure_tmp(c, (n * sizeof(wjac<F>) + sizeof(ge_ext) - 1) / sizeof(ge_ext));
    if (rc) {
        pool_free(ps->w_wire), pool_free(ps->w_base);
        return rc;
    }
    CudaBE be(c);
    KFixedBaseW<F> k = {(const waff<F> *)c->fbw_table[curve == VMSM_CURVE_BN256_G2 ? 1 : 0], dsc, seed, (wjac<F> *)c->tmp_ext};
    be.launch(k, (uint32_t)n);
    KNormalizeW<F> kn = {(const wjac<F> *)c->tmp_ext, (waff<F> *)ps->w_wire, (waff<F> *)ps->w_base};
    be.launch(kn, (uint32_t)n);
    be.note(cudaStreamSynchronize(c->stream));
    if (be.err != cudaSuccess) {
        pool_free(ps->w_wire), pool_free(ps->w_base);
        return fail(VMSM_ERR_CUDA, "fixed_base: %s", cudaGetErrorString(be.err));
    }
    return VMSM_OK;
}

template <class F>
int32_t w_precompute_ps(Ctx *c, PointSet &ps, uint32_t cb) {
    if (ps.pre && ps.pre_c == cb) return VMSM_OK;
    if (ps.n == 0) return VMSM_OK;
    CU(wait_bases_released(c, true));
    if (ps.pre) pool_free(ps.pre);
    ps.pre = nullptr;
    ps.pre_c = ps.pre_W = 0;
    const uint32_t W = (256 + cb) / cb;
    waff<F> *tbl = nullptr;
    cudaError_t e = pool_alloc(&tbl, (size_t)W * ps.n * sizeof(waff<F>));
    if (e != cudaSuccess)
        return fail(VMSM_ERR_NOMEM, "table of %u levels x %llu points: %s", W, (unsigned long long)ps.n, cudaGetErrorString(e));
    CudaBE be(c);
    KPrecomputeW<F> k = {(const waff<F> *)ps.w_base, tbl, (uint32_t)ps.n, cb, W};
    be.launch(k, (uint32_t)ps.n);
    if (be.err != cudaSuccess) {
        pool_free(tbl);
        return fail(VMSM_ERR_CUDA, "precompute: %s", cudaGetErrorString(be.err));
    }
    ps.pre = tbl;
    ps.pre_c = cb;
    ps.pre_W = W;
    ps.pre_n = ps.n;
    return VMSM_OK;
}

template <class F>
int32_t w_run_msm(Ctx *c, const PointSet &ps, uint64_t off, const uint32_t *scalars, uint64_t n_total, uint32_t slot,
                  PointSet *extra, uint64_t extra_off, uint32_t n_extra) {
    if (n_total > (1ull << 24)) return fail(VMSM_ERR_UNSUPPORTED, "BN256 MSMs are limited to 2^24 terms");
    // through the key tables (vmsm_points_precompute) when the vector has one; a few extra terms get theirs on first use
    bool pre_ok = ps.pre && n_total >= c->pre_min_terms && !c->opt.window_bits;
    if (pre_ok && n_extra && !(extra->pre && extra->pre_c == ps.pre_c)) {
        if (extra->n <= 64 && extra != &ps) {
            int32_t rcp = w_precompute_ps<F>(c, *extra, ps.pre_c);
            if (rcp) return rcp;
        } else {
            pre_ok = false;
        }
    }
    CudaBE be(c);
    c->cur_slot = slot;
    c->slot_curve[slot] = ps.curve;
    c->slot_host_norm[slot] = c->host_norm;
    MsmOptions opt = c->opt;
    wjac<F> *host_jac = c->host_norm ? (wjac<F> *)(c->res_wj_host + 192 * slot) : nullptr;
    int rc;
    if (pre_ok) {
        PreTable pt = {(uint32_t)ps.pre_n, ps.pre_c, ps.pre_W, nullptr, n_extra ? (uint32_t)extra->pre_n : 0u};
        rc = msm_run_w<CudaBE, F>(be, c->ws, opt, (const waff<F> *)ps.pre + off, scalars, (uint32_t)n_total,
                                  (wjac<F> *)(c->res_w_dev + 192 * slot), (waff<F> *)(c->res_w_host + 128 * slot), nullptr,
                                  n_extra, c->msm_seq++, &pt, n_extra ? (const waff<F> *)extra->pre + extra_off : nullptr,
                                  host_jac);
    } else {
        rc = msm_run_w<CudaBE, F>(be, c->ws, opt, (const waff<F> *)ps.w_base + off, scalars, (uint32_t)n_total,
                                  (wjac<F> *)(c->res_w_dev + 192 * slot), (waff<F> *)(c->res_w_host + 128 * slot),
                                  extra ? (const waff<F> *)extra->w_base + extra_off : nullptr, n_extra, c->msm_seq++,
                                  nullptr, nullptr, host_jac);
    }
    if (rc == -2) return fail(VMSM_ERR_INVALID, "precomputed table does not match the MSM geometry");
    if (rc) return fail(VMSM_ERR_NOMEM, "workspace allocation failed: %s", cudaGetErrorString(be.err));
    if (be.err != cudaSuccess) return fail(VMSM_ERR_CUDA, "msm launch: %s", cudaGetErrorString(be.err));
    return VMSM_OK;
}

int32_t w_run_msm_any(Ctx *c, const PointSet &ps, uint64_t off, const uint32_t *scalars, uint64_t n_total, uint32_t slot,
                      PointSet *extra = nullptr, uint64_t extra_off = 0, uint32_t n_extra = 0) {
    if (ps.curve == VMSM_CURVE_BN256_G1) return w_run_msm<FpBN>(c, ps, off, scalars, n_total, slot, extra, extra_off, n_extra);
    return w_run_msm<Fp2BN>(c, ps, off, scalars, n_total, slot, extra, extra_off, n_extra);
}

// fetch a finished result slot into `out` (64 B Ed25519 / BN256 G1, 128 B BN256 G2)
int32_t fetch_slot(Ctx *c, uint32_t slot, uint8_t *out) {
    CU(cudaEventSynchronize(c->ev_slot[slot]));
    if (uint32_t st = c->res_status_host[slot]) {
        c->res_status_host[slot] = 0;
        if (st == 2) return fail(VMSM_ERR_POINT, "invalid point in the input of the asynchronous call for slot %u", slot);
        return fail(VMSM_ERR_TIMEOUT, "a multi-GPU partial for slot %u did not arrive", slot);
    }
    if (c->slot_curve[slot] == VMSM_CURVE_ED25519) {
        if (c->slot_host_norm[slot]) {
            ge_aff a = ge_ext_to_aff(c->res_xyz_host[slot]);  // portable host path of fe25519.cuh: one inversion on the CPU
            memcpy(out, &a, sizeof(ge_aff));
        } else {
            memcpy(out, c->res_aff_host + slot, sizeof(ge_aff));
        }
    } else if (c->slot_host_norm[slot]) {  // Jacobian from the device, one inversion in Fp / Fp2 on the CPU
        if (c->slot_curve[slot] == VMSM_CURVE_BN256_G1) {
            waff<FpBN> w = wa_to_wire(wj_to_aff(*(const wjac<FpBN> *)(c->res_wj_host + 192 * slot)));
            memcpy(out, &w, sizeof(w));
        } else {
            waff<Fp2BN> w = wa_to_wire(wj_to_aff(*(const wjac<Fp2BN> *)(c->res_wj_host + 192 * slot)));
            memcpy(out, &w, sizeof(w));
        }
    } else {
        memcpy(out, c->res_w_host + 128 * slot, wire_bytes(c->slot_curve[slot]));
    }
    return VMSM_OK;
}

}  // namespace

// Entry points on one context are serialised by a per-context lock: host threads may share a context (the prover
// hashes its first pre-image on a worker thread; Python finalizers free handles from whatever thread the collector
// runs on) and ctypes drops the GIL around every call.
#define GET_CTX(h)                                                        \
    Ctx *c = get_ctx(h);                                                  \
    if (!c) return fail(VMSM_ERR_INVALID, "invalid context handle");      \
    std::lock_guard<std::recursive_mutex> ctx_lock(c->mu);                \
    tl_ctx = c;                                                           \
    CU(cudaSetDevice(c->device))

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" {

int32_t vmsm_version(void) { return 100; }
const char *vmsm_last_error(void) { return g_err; }

int32_t vmsm_device_count(int32_t *count) {
    if (!count) return fail(VMSM_ERR_INVALID, "null argument");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(VMSM_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return VMSM_OK;
}

int32_t vmsm_ctx_create(int32_t device, uint64_t *ctx) {
    if (!ctx) return fail(VMSM_ERR_INVALID, "null argument");
    *ctx = 0;
    int n = 0;
    CU(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(VMSM_ERR_INVALID, "device %d out of range (have %d)", device, n);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(VMSM_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    Ctx *c = new Ctx();
    c->device = device;
    c->sms = (uint32_t)prop.multiProcessorCount;
    c->sort_blocks = 2 * c->sms;  // see CudaBE::launch_sort
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int h = 0; h < 2; h++) {
        CU(cudaStreamCreateWithFlags(&c->heads[h], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->ev_head_done[h], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&c->ev_issue, cudaEventDisableTiming));
    {
        int lo = 0, hi = 0;  // hi = numerically smallest = greatest priority
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        for (int w = 0; w < kTailWays; w++) CU(cudaStreamCreateWithPriority(&c->tails[w], cudaStreamNonBlocking, hi));
    }
    {
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(cudaStreamCreateWithPriority(&c->sort, cudaStreamNonBlocking, hi));
    }
    CU(cudaEventCreateWithFlags(&c->ev_sort_in, cudaEventDisableTiming));
    for (int k = 0; k < 2; k++) {
        CU(cudaEventCreateWithFlags(&c->ev_sorted[k], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_acc_done[k], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&c->ev_head, cudaEventDisableTiming));
    for (int w = 0; w < kTailWays; w++) CU(cudaEventCreateWithFlags(&c->ev_tail[w], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_sc_written, cudaEventDisableTiming));
    CU(cudaMalloc(&c->dot_scratch, (kDotT1 + kDotT2 + 3) * 32));  // partial sums, the result, two results kept on the device
    CU(cudaMalloc(&c->la_aff, kLaRing * 64 * sizeof(ge_aff)));
    CU(cudaMalloc(&c->la_niels, kLaRing * 64 * sizeof(ge_niels)));
    CU(cudaMalloc(&c->la_scalars, kLaRing * 64 * 32));
    CU(cudaMalloc(&c->la_err, kLaRing * 4));
    for (uint32_t k = 0; k < kLaRing; k++) CU(cudaEventCreateWithFlags(&c->ev_la_sorted[k], cudaEventDisableTiming));
    CU(cudaMalloc(&c->order_bins, ORDER_BINS * 4));
    CU(cudaMalloc(&c->err_word, 16));
    CU(cudaMalloc(&c->fb_table, 512 * sizeof(ge_niels)));
    CU(cudaMalloc(&c->res_ext, kSlots * sizeof(ge_ext)));
    CU(cudaMalloc(&c->res_aff, kSlots * sizeof(ge_aff)));
    CU(cudaMalloc(&c->small_aff, 64 * sizeof(ge_aff)));
    CU(cudaMalloc(&c->small_niels, 64 * sizeof(ge_niels)));
    CU(cudaHostAlloc(&c->pin, 4096, cudaHostAllocDefault));
    CU(cudaHostAlloc(&c->res_aff_host, kSlots * sizeof(ge_aff), cudaHostAllocMapped));
    CU(cudaHostAlloc(&c->res_xyz_host, kSlots * sizeof(ge_ext), cudaHostAllocMapped));
    CU(cudaHostAlloc(&c->res_wj_host, kSlots * 192, cudaHostAllocMapped));
    CU(cudaHostAlloc(&c->res_status_host, kSlots * sizeof(uint32_t), cudaHostAllocMapped));
    CU(cudaMalloc(&c->res_w_dev, kSlots * 192));
    CU(cudaHostAlloc(&c->res_w_host, kSlots * 128, cudaHostAllocMapped));
    CU(cudaMalloc(&c->small_w_wire, 64 * 128));
    CU(cudaMalloc(&c->small_w_base, 64 * 128));
    memset(c->res_status_host, 0, kSlots * sizeof(uint32_t));
    CU(cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
        CU(cudaEventCreateWithFlags(&c->ev_copied[k], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_consumed[k], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_dot[k], cudaEventDisableTiming));
    }
    for (uint32_t k = 0; k < kSlots; k++) CU(cudaEventCreateWithFlags(&c->ev_slot[k], cudaEventDisableTiming));
    CU(cudaEventCreate(&c->t0));
    CU(cudaEventCreate(&c->t1));
    {
        std::vector<ge_niels> tbl(512);
        build_fixed_base_table(tbl.data());
        CU(cudaMemcpy(c->fb_table, tbl.data(), 512 * sizeof(ge_niels), cudaMemcpyHostToDevice));
    }
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_ctxs.insert(c);
    }
    *ctx = reinterpret_cast<uint64_t>(c);
    return VMSM_OK;
}

int32_t vmsm_ctx_destroy(uint64_t ctx) {
    Ctx *c = get_ctx(ctx);
    if (!c) return fail(VMSM_ERR_INVALID, "invalid context handle");
    {  // unpublish first (no new call can find the handle), then let a call still inside the context drain
        std::lock_guard<std::mutex> lk(g_mu);
        g_ctxs.erase(c);
    }
    c->mu.lock();
    c->mu.unlock();
    tl_ctx = c;
    CU(cudaSetDevice(c->device));
    cudaStreamSynchronize(c->copy);
    cudaStreamSynchronize(c->sort);
    for (int w = 0; w < kTailWays; w++) cudaStreamSynchronize(c->tails[w]);
    for (int h = 0; h < 2; h++) cudaStreamSynchronize(c->heads[h]);
    cudaStreamSynchronize(c->stream);
    for (int k = 0; k < 2; k++)
        cudaFree(c->astage[k]), cudaEventDestroy(c->ev_copied[k]), cudaEventDestroy(c->ev_consumed[k]), cudaEventDestroy(c->ev_dot[k]);
    for (uint32_t k = 0; k < kSlots; k++) cudaEventDestroy(c->ev_slot[k]);
    cudaFreeHost(c->res_aff_host);
    cudaFreeHost(c->res_xyz_host);
    cudaFreeHost(c->res_wj_host);
    cudaFreeHost(c->res_status_host);
    cudaEventDestroy(c->ev_sc_written);
    cudaFree(c->dot_scratch);
    cudaFree(c->la_aff), cudaFree(c->la_niels), cudaFree(c->la_scalars), cudaFree(c->la_err);
    for (uint32_t k = 0; k < kLaRing; k++) cudaEventDestroy(c->ev_la_sorted[k]);
    cudaFree(c->txt_slots), cudaFree(c->txt_text), cudaFree(c->txt_lens), cudaFree(c->txt_offsets), cudaFree(c->txt_sums);
    if (c->txt_host) cudaFreeHost(c->txt_host);
    if (c->txt_host_sc) cudaFreeHost(c->txt_host_sc);
    cudaFree(c->res_w_dev), cudaFreeHost(c->res_w_host), cudaFree(c->fbw_table[0]), cudaFree(c->fbw_table[1]);
    cudaFree(c->small_w_wire), cudaFree(c->small_w_base);
    if (c->mailbox) {
        if (c->mb_owner) cudaFree(c->mailbox);
        else if (c->mb_ipc) cudaIpcCloseMemHandle(c->mailbox);
    }
    cudaStreamDestroy(c->copy);
    for (auto &kv : c->points)
        cudaFree(kv.second.aff), cudaFree(kv.second.niels), cudaFree(kv.second.w_wire), cudaFree(kv.second.w_base),
            cudaFree(kv.second.pre);
    for (auto &kv : c->scalars) cudaFree(kv.second.data);
    for (auto &kv : c->pool) cudaFree(kv.second);
    CudaBE be(c);
    ws_release(be, c->ws);
    for (int k = 0; k < 2; k++) cudaFree(c->bs_dig[k]), cudaFree(c->bs_hist[k]);
    cudaFree(c->order_bins), cudaFree(c->err_word), cudaFree(c->fb_table), cudaFree(c->res_ext), cudaFree(c->res_aff);
    cudaFree(c->stage_scalars), cudaFree(c->tmp_ext), cudaFree(c->small_aff), cudaFree(c->small_niels);
    cudaFreeHost(c->pin);
    for (auto &es : c->ev_pool)
        for (auto &e : es.ev) cudaEventDestroy(e);
    cudaEventDestroy(c->t0), cudaEventDestroy(c->t1);
    cudaEventDestroy(c->ev_head);
    for (int w = 0; w < kTailWays; w++) cudaEventDestroy(c->ev_tail[w]), cudaStreamDestroy(c->tails[w]);
    cudaStreamDestroy(c->sort);
    cudaEventDestroy(c->ev_sort_in);
    for (int k = 0; k < 2; k++) cudaEventDestroy(c->ev_sorted[k]), cudaEventDestroy(c->ev_acc_done[k]);
    for (int h = 0; h < 2; h++) cudaStreamDestroy(c->heads[h]), cudaEventDestroy(c->ev_head_done[h]);
    cudaEventDestroy(c->ev_issue);
    cudaStreamDestroy(c->stream);
    delete c;
    return VMSM_OK;
}

int32_t vmsm_ctx_set_option(uint64_t ctx, int32_t key, int64_t value) {
    GET_CTX(ctx);
    switch (key) {
        case VMSM_OPT_WINDOW_BITS:
            if (value != 0 && (value < 2 || value > 18)) return fail(VMSM_ERR_INVALID, "window bits must be 0 or in [2, 18]");
            c->opt.window_bits = (uint32_t)value;
            return VMSM_OK;
        case VMSM_OPT_PHASE_TIMING:
            c->phase_timing = value != 0;
            return VMSM_OK;
        case VMSM_OPT_SORT_BUCKETS:
            c->opt.sort_buckets = value != 0;
            return VMSM_OK;
        case VMSM_OPT_CHECK_POINTS:
            c->check_points = value != 0;
            return VMSM_OK;
        case VMSM_OPT_REDUCE_RADIX:
            if (value < 1 || value > 6) return fail(VMSM_ERR_INVALID, "reduce radix log2 must be in [1, 6]");
            c->opt.reduce_log2r = c->opt.reduce_log2r_w = (uint32_t)value;
            return VMSM_OK;
        case VMSM_OPT_SHARD_SEQ:
            if (value < 0 || value > 0xffffffffll) return fail(VMSM_ERR_INVALID, "shard seq out of range");
            c->shard_next = (uint32_t)value;
            return VMSM_OK;
        case VMSM_OPT_CAP_FACTOR:
            if (value < 1 || value > 65536) return fail(VMSM_ERR_INVALID, "cap factor out of range");
            c->opt.cap_factor = (uint32_t)value;
            return VMSM_OK;
        case VMSM_OPT_ASYNC_TAIL:
            c->async_tail = value != 0;
            return VMSM_OK;
        case VMSM_OPT_ASYNC_SORT:
            c->async_sort = value != 0;
            return VMSM_OK;
        case VMSM_OPT_PRE_SETS:
            if (value < 0 || value > 32) return fail(VMSM_ERR_INVALID, "bucket sets out of range");
            c->opt.pre_sets = (uint32_t)value;
            return VMSM_OK;
        case VMSM_OPT_DUAL_HEAD:
            CU(join_tail(c));
            CU(cudaStreamSynchronize(c->stream));
            c->dual_head = value != 0;
            return VMSM_OK;
        case VMSM_OPT_HOST_NORMALIZE:
            c->host_norm = value != 0;
            return VMSM_OK;
        case VMSM_OPT_SEG_MODE:
            if (value < 0 || value > 2) return fail(VMSM_ERR_INVALID, "segment mode out of range");
            c->opt.seg_mode = (uint32_t)value;
            return VMSM_OK;
        case VMSM_OPT_SEG_LEN:
            if (value < 0 || value > 4096) return fail(VMSM_ERR_INVALID, "segment length out of range");
            c->opt.seg_len = (uint32_t)value;
            return VMSM_OK;
        case VMSM_OPT_PRE_MIN_TERMS:
            if (value < 0) return fail(VMSM_ERR_INVALID, "negative term count");
            c->pre_min_terms = (uint64_t)value;
            return VMSM_OK;
        case VMSM_OPT_BN_QUAD_ACC:
            if (value < 0 || value > 2) return fail(VMSM_ERR_INVALID, "BN quad accumulate mode out of range");
            c->opt.w_quad_acc = (uint32_t)value;
            return VMSM_OK;
        case VMSM_OPT_FOLD_QUAD_MAX:
            if (value < 0 || value > (1 << 26)) return fail(VMSM_ERR_INVALID, "fold quad threshold out of range");
            c->fold_quad_max = (uint32_t)value;
            return VMSM_OK;
        case VMSM_OPT_SORT_BLOCKS:
            if (value < -1 || value > (1 << 20)) return fail(VMSM_ERR_INVALID, "sort blocks out of range");
            c->sort_blocks_auto = value < 0;
            if (value >= 0) c->sort_blocks = (uint32_t)value;
            return VMSM_OK;
        case VMSM_OPT_QUAD_THRESHOLD:
            if (value < 0 || value > (1 << 24)) return fail(VMSM_ERR_INVALID, "quad threshold out of range");
            c->opt.quad_threshold = (uint32_t)value;
            return VMSM_OK;
        case VMSM_OPT_BN_PRE_SETS:
            if (value < 0 || value > 64) return fail(VMSM_ERR_INVALID, "bucket sets out of range");
            c->opt.pre_sets_w = (uint32_t)value;
            return VMSM_OK;
        case VMSM_OPT_BN_QUAD_FIX:
            c->opt.w_quad_fix = value != 0;
            return VMSM_OK;
        case VMSM_OPT_BN_SEG_LEN:
            if (value < 0 || value > 4096) return fail(VMSM_ERR_INVALID, "segment length out of range");
            c->opt.seg_len_w = (uint32_t)value;
            return VMSM_OK;
        case VMSM_OPT_BLOCK_SORT:
            c->opt.block_sort = value != 0;
            return VMSM_OK;
        case VMSM_OPT_BLOCK_SORT_MIN:
            if (value < 0 || value > (1ll << 26)) return fail(VMSM_ERR_INVALID, "block sort threshold out of range");
            c->bs_min_terms = (uint32_t)value;
            return VMSM_OK;
        case VMSM_OPT_ACC_CARVEOUT: {
            if (value < -1 || value > 100) return fail(VMSM_ERR_INVALID, "carveout must be -1 (driver default) or a percentage");
            const int pct = (int)value;
            CU(cudaFuncSetAttribute(vmsm_kernel<KAccumulate>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
            CU(cudaFuncSetAttribute(vmsm_kernel<KAccumulatePre>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
            CU(cudaFuncSetAttribute(vmsm_kernel<KAccumulateSeg>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
            CU(cudaFuncSetAttribute(vmsm_kernel<KAccumulateSegPre>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
            return VMSM_OK;
        }
    }
    return fail(VMSM_ERR_INVALID, "unknown option %d", key);
}

int32_t vmsm_sync(uint64_t ctx) {
    GET_CTX(ctx);
    CU(join_tail(c));
    CU(cudaStreamSynchronize(c->stream));
    return VMSM_OK;
}

int32_t vmsm_timer_start(uint64_t ctx) {
    GET_CTX(ctx);
    CU(cudaEventRecord(c->t0, c->stream));
    return VMSM_OK;
}
int32_t vmsm_timer_stop(uint64_t ctx, float *ms) {
    GET_CTX(ctx);
    if (!ms) return fail(VMSM_ERR_INVALID, "null argument");
    CU(join_tail(c));
    CU(cudaEventRecord(c->t1, c->stream));
    CU(cudaEventSynchronize(c->t1));
    CU(cudaEventElapsedTime(ms, c->t0, c->t1));
    return VMSM_OK;
}

int32_t vmsm_phase_times(uint64_t ctx, double *ms_out, uint64_t *calls) {
    GET_CTX(ctx);
    if (!ms_out || !calls) return fail(VMSM_ERR_INVALID, "null argument");
    int32_t rc = harvest_phases(c);
    if (rc) return rc;
    for (int p = 0; p < PH_COUNT; p++) ms_out[p] = c->phase_ms[p], c->phase_ms[p] = 0;
    *calls = c->phase_calls;
    c->phase_calls = 0;
    return VMSM_OK;
}

int32_t vmsm_launch_count(uint64_t ctx, uint64_t *launches) {
    GET_CTX(ctx);
    if (!launches) return fail(VMSM_ERR_INVALID, "null argument");
    *launches = c->launches;
    return VMSM_OK;
}

// ---- points
static int32_t new_pointset(Ctx *c, int32_t curve, uint64_t n, PointSet *ps) {
    ps->curve = curve;
    ps->n = n;
    ps->aff = nullptr;
    ps->niels = nullptr;
    ps->w_wire = ps->w_base = nullptr;
    CU(pool_alloc(&ps->aff, (n ? n : 1) * sizeof(ge_aff)));
    cudaError_t e = pool_alloc(&ps->niels, (n ? n : 1) * sizeof(ge_niels));
    if (e != cudaSuccess) {
        pool_free(ps->aff);
        return fail(VMSM_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
    }
    return VMSM_OK;
}

int32_t vmsm_points_upload(uint64_t ctx, int32_t curve, const uint8_t *affine, uint64_t n, uint64_t *pts) {
    GET_CTX(ctx);
    if (!pts || (n && !affine)) return fail(VMSM_ERR_INVALID, "null argument");
    if (curve < VMSM_CURVE_ED25519 || curve > VMSM_CURVE_BN256_G2) return fail(VMSM_ERR_UNSUPPORTED, "unknown curve %d", curve);
    if (n > (1ull << 28)) return fail(VMSM_ERR_UNSUPPORTED, "too many points");
    PointSet ps;
    if (curve != VMSM_CURVE_ED25519) {
        int32_t rcw = curve == VMSM_CURVE_BN256_G1 ? w_upload<FpBN>(c, curve, affine, n, &ps)
                                                   : w_upload<Fp2BN>(c, curve, affine, n, &ps);
        if (rcw) return rcw;
        uint64_t idw = c->next_id++;
        c->points[idw] = ps;
        *pts = idw;
        return VMSM_OK;
    }
    int32_t rc = new_pointset(c, curve, n, &ps);
    if (rc) return rc;
    CudaBE be(c);
    if (n) {
        be.note(cudaMemcpyAsync(ps.aff, affine, n * sizeof(ge_aff), cudaMemcpyHostToDevice, c->stream));
        be.zero(c->err_word, 4);
        KAffToNiels k = {ps.aff, ps.niels, c->err_word, c->check_points ? 1u : 0u};
        be.launch(k, (uint32_t)n);
        be.note(cudaMemcpyAsync(c->pin, c->err_word, 4, cudaMemcpyDeviceToHost, c->stream));
        be.note(cudaStreamSynchronize(c->stream));
        uint32_t ew = *reinterpret_cast<uint32_t *>(c->pin);
        if (be.err != cudaSuccess || ew) {
            pool_free(ps.aff), pool_free(ps.niels);
            if (be.err != cudaSuccess) return fail(VMSM_ERR_CUDA, "upload: %s", cudaGetErrorString(be.err));
            return fail(VMSM_ERR_POINT, "invalid point in upload (%s%s)", (ew & 1) ? "non-canonical coordinate " : "",
                        (ew & 2) ? "not on curve" : "");
        }
    }
    uint64_t id = c->next_id++;
    c->points[id] = ps;
    *pts = id;
    return VMSM_OK;
}

int32_t vmsm_points_fixed_base(uint64_t ctx, int32_t curve, const uint8_t *scalars, uint64_t seed, uint64_t n,
                               uint64_t *pts) {
    GET_CTX(ctx);
    if (!pts) return fail(VMSM_ERR_INVALID, "null argument");
    if (curve < VMSM_CURVE_ED25519 || curve > VMSM_CURVE_BN256_G2) return fail(VMSM_ERR_UNSUPPORTED, "unknown curve %d", curve);
    if (n > (1ull << 28)) return fail(VMSM_ERR_UNSUPPORTED, "too many points");
    PointSet ps;
    if (curve != VMSM_CURVE_ED25519) {
        int32_t rcw = curve == VMSM_CURVE_BN256_G1 ? w_fixed_base<FpBN>(c, curve, scalars, seed, n, &ps)
                                                   : w_fixed_base<Fp2BN>(c, curve, scalars, seed, n, &ps);
        if (rcw) return rcw;
        uint64_t idw = c->next_id++;
        c->points[idw] = ps;
        *pts = idw;
        return VMSM_OK;
    }
    int32_t rc = new_pointset(c, curve, n, &ps);
    if (rc) return rc;
    if (n) {
        const uint32_t *dsc = nullptr;
        if (scalars) {
            rc = ensure_stage(c, n);
            if (!rc) {
                cudaError_t e = cudaMemcpyAsync(c->stage_scalars, scalars, n * 32, cudaMemcpyHostToDevice, c->stream);
                if (e != cudaSuccess) rc = fail(VMSM_ERR_CUDA, "H2D: %s", cudaGetErrorString(e));
            }
            dsc = c->stage_scalars;
        }
        if (!rc) rc = ensure_tmp(c, n);
        if (rc) {
            pool_free(ps.aff), pool_free(ps.niels);
            return rc;
        }
        CudaBE be(c);
        KFixedBase k = {c->fb_table, dsc, seed, c->tmp_ext};
        be.launch(k, (uint32_t)n);
        KNormalize kn = {c->tmp_ext, ps.aff, ps.niels};
        be.launch(kn, (uint32_t)n);
        be.note(cudaStreamSynchronize(c->stream));
        if (be.err != cudaSuccess) {
            pool_free(ps.aff), pool_free(ps.niels);
            return fail(VMSM_ERR_CUDA, "fixed_base: %s", cudaGetErrorString(be.err));
        }
    }
    uint64_t id = c->next_id++;
    c->points[id] = ps;
    *pts = id;
    return VMSM_OK;
}

int32_t vmsm_points_precompute(uint64_t ctx, uint64_t pts, uint32_t window_bits) {
    GET_CTX(ctx);
    auto it = c->points.find(pts);
    if (it == c->points.end()) return fail(VMSM_ERR_INVALID, "invalid points handle");
    if (window_bits != 0 && (window_bits < 8 || window_bits > 16))
        return fail(VMSM_ERR_INVALID, "table window must be 0 (auto) or in [8, 16]");
    if (it->second.curve == VMSM_CURVE_BN256_G1) return w_precompute_ps<FpBN>(c, it->second, window_bits ? window_bits : 13u);
    if (it->second.curve == VMSM_CURVE_BN256_G2) return w_precompute_ps<Fp2BN>(c, it->second, window_bits ? window_bits : 13u);
    return precompute_ps(c, it->second, window_bits ? window_bits : (it->second.n >= (1u << 13) ? 16u : 13u));
}

int32_t vmsm_points_download(uint64_t ctx, uint64_t pts, uint64_t off, uint64_t n, uint8_t *affine_out) {
    GET_CTX(ctx);
    auto it = c->points.find(pts);
    if (it == c->points.end()) return fail(VMSM_ERR_INVALID, "invalid points handle");
    if (off > it->second.n || n > it->second.n - off) return fail(VMSM_ERR_INVALID, "range out of bounds");
    if (n && !affine_out) return fail(VMSM_ERR_INVALID, "null argument");
    if (n && it->second.curve == VMSM_CURVE_ED25519)
        CU(cudaMemcpyAsync(affine_out, it->second.aff + off, n * sizeof(ge_aff), cudaMemcpyDeviceToHost, c->stream));
    else if (n) {
        size_t wb = wire_bytes(it->second.curve);
        CU(cudaMemcpyAsync(affine_out, (const uint8_t *)it->second.w_wire + off * wb, n * wb, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return VMSM_OK;
}

// grow-only scratch of the text pipeline, sized for n slots of VMSM_TEXT_SLOT bytes
static int32_t text_ensure(Ctx *c, uint64_t n) {
    if (n <= c->txt_cap) return VMSM_OK;
    cudaFree(c->txt_slots), cudaFree(c->txt_text), cudaFree(c->txt_lens), cudaFree(c->txt_offsets), cudaFree(c->txt_sums);
    if (c->txt_host) cudaFreeHost(c->txt_host);
    if (c->txt_host_sc) cudaFreeHost(c->txt_host_sc);
    c->txt_slots = c->txt_text = c->txt_host = c->txt_host_sc = nullptr;
    c->txt_lens = nullptr;
    c->txt_offsets = c->txt_sums = nullptr;
    c->txt_cap = 0;
    size_t cap = n + n / 4 + 1024;
    CU(cudaMalloc(&c->txt_slots, cap * VMSM_TEXT_SLOT));
    CU(cudaMalloc(&c->txt_text, cap * VMSM_TEXT_SLOT));
    CU(cudaMalloc(&c->txt_lens, cap * 4));
    CU(cudaMalloc(&c->txt_offsets, cap * 8));
    CU(cudaMalloc(&c->txt_sums, (cap / 1024 + 2) * 8));
    CU(cudaHostAlloc(&c->txt_host, cap * VMSM_TEXT_SLOT, cudaHostAllocDefault));
    CU(cudaHostAlloc(&c->txt_host_sc, cap * VMSM_SCALAR_TEXT_SLOT, cudaHostAllocDefault));
    c->txt_cap = cap;
    return VMSM_OK;
}

// slots + lens (already launched on the main stream) -> ", "-joined text in c->txt_host; *len = its length
static int32_t text_finish(Ctx *c, CudaBE &be, uint64_t n, uint32_t slot_bytes, uint64_t *len, uint8_t *host = nullptr) {
    if (!host) host = c->txt_host;
    uint32_t nblk = (uint32_t)((n + 1023) / 1024);
    vmsm_lens_block_sums<<<nblk, 1024, 0, c->stream>>>(c->txt_lens, (uint32_t)n, c->txt_sums);
    vmsm_lens_scan_sums<<<1, 32, 0, c->stream>>>(c->txt_sums, nblk);
    vmsm_lens_offsets<<<nblk, 1024, 0, c->stream>>>(c->txt_lens, (uint32_t)n, c->txt_sums, c->txt_offsets);
    c->launches += 3;
    KTextCompact kc = {c->txt_slots, c->txt_lens, c->txt_offsets, c->txt_text, slot_bytes};
    be.launch(kc, (uint32_t)n);
    be.note(cudaMemcpyAsync(c->pin, c->txt_sums + nblk, 8, cudaMemcpyDeviceToHost, c->stream));
    be.note(cudaStreamSynchronize(c->stream));
    if (be.err != cudaSuccess) return fail(VMSM_ERR_CUDA, "text: %s", cudaGetErrorString(be.err));
    uint64_t total = *reinterpret_cast<uint64_t *>(c->pin);
    CU(cudaMemcpyAsync(host, c->txt_text, total, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *len = total;
    return VMSM_OK;
}

// canonical bytes of a device range straight into the context's page-locked buffer (binary transcript mode)
static int32_t bytes_to_pinned(Ctx *c, const void *dev, size_t nbytes, const uint8_t **ptr) {
    int32_t rc = text_ensure(c, nbytes / VMSM_TEXT_SLOT + 1);
    if (rc) return rc;
    if (nbytes) CU(cudaMemcpyAsync(c->txt_host, dev, nbytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *ptr = c->txt_host;
    return VMSM_OK;
}

static int32_t points_text_impl(Ctx *c, uint64_t pts, uint64_t off, uint64_t n, uint64_t *len) {
    auto it = c->points.find(pts);
    if (it == c->points.end()) return fail(VMSM_ERR_INVALID, "invalid points handle");
    if (it->second.curve != VMSM_CURVE_ED25519) return fail(VMSM_ERR_UNSUPPORTED, "points_text: Ed25519 only");
    if (off > it->second.n || n > it->second.n - off) return fail(VMSM_ERR_INVALID, "range out of bounds");
    *len = 0;
    if (!n) return VMSM_OK;
    if (n > (1ull << 26)) return fail(VMSM_ERR_UNSUPPORTED, "too many points");
    int32_t rc = text_ensure(c, n);
    if (rc) return rc;
    CudaBE be(c);
    KPointText kt = {it->second.aff + off, c->txt_slots, c->txt_lens, (uint32_t)n};
    be.launch(kt, (uint32_t)n);
    return text_finish(c, be, n, VMSM_TEXT_SLOT, len);
}

int32_t vmsm_points_download_ptr(uint64_t ctx, uint64_t pts, uint64_t off, uint64_t n, const uint8_t **affine) {
    GET_CTX(ctx);
    if (!affine) return fail(VMSM_ERR_INVALID, "null argument");
    auto it = c->points.find(pts);
    if (it == c->points.end()) return fail(VMSM_ERR_INVALID, "invalid points handle");
    if (off > it->second.n || n > it->second.n - off) return fail(VMSM_ERR_INVALID, "range out of bounds");
    if (n > (1ull << 26)) return fail(VMSM_ERR_UNSUPPORTED, "too many points");
    if (it->second.curve == VMSM_CURVE_ED25519) return bytes_to_pinned(c, it->second.aff + off, n * sizeof(ge_aff), affine);
    size_t wb = wire_bytes(it->second.curve);
    return bytes_to_pinned(c, (const uint8_t *)it->second.w_wire + off * wb, n * wb, affine);
}

int32_t vmsm_scalars_download_ptr(uint64_t ctx, uint64_t sc, uint64_t off, uint64_t n, const uint8_t **le32) {
    GET_CTX(ctx);
    if (!le32) return fail(VMSM_ERR_INVALID, "null argument");
    auto it = c->scalars.find(sc);
    if (it == c->scalars.end()) return fail(VMSM_ERR_INVALID, "invalid scalars handle");
    if (off > it->second.n || n > it->second.n - off) return fail(VMSM_ERR_INVALID, "range out of bounds");
    if (n > (1ull << 28)) return fail(VMSM_ERR_UNSUPPORTED, "too many scalars");
    return bytes_to_pinned(c, it->second.data + off * 8, n * 32, le32);
}

int32_t vmsm_points_text(uint64_t ctx, uint64_t pts, uint64_t off, uint64_t n, uint8_t *out, uint64_t cap,
                         uint64_t *len) {
    GET_CTX(ctx);
    if (!len || (n && !out)) return fail(VMSM_ERR_INVALID, "null argument");
    int32_t rc = points_text_impl(c, pts, off, n, len);
    if (rc) return rc;
    if (*len > cap) return fail(VMSM_ERR_INVALID, "text buffer too small: need %llu bytes", (unsigned long long)*len);
    if (*len) memcpy(out, c->txt_host, *len);
    return VMSM_OK;
}

int32_t vmsm_points_text_ptr(uint64_t ctx, uint64_t pts, uint64_t off, uint64_t n, const uint8_t **text,
                             uint64_t *len) {
    GET_CTX(ctx);
    if (!len || !text) return fail(VMSM_ERR_INVALID, "null argument");
    int32_t rc = points_text_impl(c, pts, off, n, len);
    if (rc) return rc;
    *text = c->txt_host;
    return VMSM_OK;
}

int32_t vmsm_points_count(uint64_t ctx, uint64_t pts, uint64_t *n) {
    GET_CTX(ctx);
    auto it = c->points.find(pts);
    if (it == c->points.end() || !n) return fail(VMSM_ERR_INVALID, "invalid points handle");
    *n = it->second.n;
    return VMSM_OK;
}

int32_t vmsm_points_free(uint64_t ctx, uint64_t pts) {
    GET_CTX(ctx);
    auto it = c->points.find(pts);
    if (it == c->points.end()) return fail(VMSM_ERR_INVALID, "invalid points handle");
    CU(cudaStreamSynchronize(c->stream));
    CU(wait_bases_released(c, true));  // overflow kernels of MSMs still in flight read the bases on tail streams
    pool_free(it->second.aff), pool_free(it->second.niels), pool_free(it->second.w_wire), pool_free(it->second.w_base);
    pool_free(it->second.pre);
    c->points.erase(it);
    return VMSM_OK;
}

// ---- scalars
int32_t vmsm_scalars_upload(uint64_t ctx, const uint8_t *le32, uint64_t n, uint64_t *sc) {
    GET_CTX(ctx);
    if (!sc || (n && !le32)) return fail(VMSM_ERR_INVALID, "null argument");
    ScalarSet ss{n, nullptr};
    CU(pool_alloc(&ss.data, (n ? n : 1) * 32));
    if (n) {
        cudaError_t e = cudaMemcpyAsync(ss.data, le32, n * 32, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) {
            pool_free(ss.data);
            return fail(VMSM_ERR_CUDA, "H2D: %s", cudaGetErrorString(e));
        }
    }
    uint64_t id = c->next_id++;
    c->scalars[id] = ss;
    *sc = id;
    return VMSM_OK;
}

int32_t vmsm_scalars_synth(uint64_t ctx, int32_t curve, uint64_t seed, uint64_t n, uint64_t *sc) {
    GET_CTX(ctx);
    if (!sc) return fail(VMSM_ERR_INVALID, "null argument");
    if (curve < VMSM_CURVE_ED25519 || curve > VMSM_CURVE_BN256_G2) return fail(VMSM_ERR_UNSUPPORTED, "unknown curve %d", curve);
    if (n > (1ull << 28)) return fail(VMSM_ERR_UNSUPPORTED, "too many scalars");
    ScalarSet ss{n, nullptr};
    CU(pool_alloc(&ss.data, (n ? n : 1) * 32));
    CudaBE be(c);
    if (curve == VMSM_CURVE_ED25519) {
        KSynthScalars k = {ss.data, seed};
        be.launch(k, (uint32_t)n);
    } else {
        KSynthScalarsBN k = {ss.data, seed};
        be.launch(k, (uint32_t)n);
    }
    be.note(cudaStreamSynchronize(c->stream));
    if (be.err != cudaSuccess) {
        pool_free(ss.data);
        return fail(VMSM_ERR_CUDA, "synth: %s", cudaGetErrorString(be.err));
    }
    uint64_t id = c->next_id++;
    c->scalars[id] = ss;
    *sc = id;
    return VMSM_OK;
}

int32_t vmsm_scalars_download(uint64_t ctx, uint64_t sc, uint64_t off, uint64_t n, uint8_t *le32_out) {
    GET_CTX(ctx);
    auto it = c->scalars.find(sc);
    if (it == c->scalars.end()) return fail(VMSM_ERR_INVALID, "invalid scalars handle");
    if (off > it->second.n || n > it->second.n - off) return fail(VMSM_ERR_INVALID, "range out of bounds");
    if (n && !le32_out) return fail(VMSM_ERR_INVALID, "null argument");
    if (n) CU(cudaMemcpyAsync(le32_out, it->second.data + off * 8, n * 32, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VMSM_OK;
}

int32_t vmsm_scalars_free(uint64_t ctx, uint64_t sc) {
    GET_CTX(ctx);
    auto it = c->scalars.find(sc);
    if (it == c->scalars.end()) return fail(VMSM_ERR_INVALID, "invalid scalars handle");
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->copy));  // msm_dev_ext stages device scalars on the copy stream
    CU(cudaStreamSynchronize(c->sort));
    pool_free(it->second.data);
    c->scalars.erase(it);
    return VMSM_OK;
}

// ---- MSM
int32_t vmsm_msm(uint64_t ctx, uint64_t pts, uint64_t off, uint64_t n, const uint8_t *scalars_le32,
                 uint8_t *out_affine) {
    GET_CTX(ctx);
    auto it = c->points.find(pts);
    if (it == c->points.end()) return fail(VMSM_ERR_INVALID, "invalid points handle");
    if (off > it->second.n || n > it->second.n - off) return fail(VMSM_ERR_INVALID, "Not enough generators.");
    if (!out_affine || (n && !scalars_le32)) return fail(VMSM_ERR_INVALID, "null argument");
    int32_t rc = ensure_stage(c, n ? n : 1);
    if (rc) return rc;
    if (n) CU(cudaMemcpyAsync(c->stage_scalars, scalars_le32, n * 32, cudaMemcpyHostToDevice, c->stream));
    CU(cudaEventRecord(c->ev_sort_in, c->stream));
    c->scalars_ready = c->ev_sort_in;
    if (it->second.curve == VMSM_CURVE_ED25519) rc = run_msm_ps(c, it->second, off, c->stage_scalars, n, kSlots - 1);
    else rc = w_run_msm_any(c, it->second, off, c->stage_scalars, n, kSlots - 1);
    if (rc) return rc;
    return fetch_slot(c, kSlots - 1, out_affine);  // the final kernel wrote the point into mapped pinned memory
}

int32_t vmsm_msm_ext(uint64_t ctx, uint64_t pts, uint64_t off, uint64_t n, uint64_t extra_pts, uint64_t extra_off,
                     uint64_t n_extra, const uint8_t *scalars_le32, uint8_t *out_affine) {
    GET_CTX(ctx);
    auto it = c->points.find(pts);
    auto ie = c->points.find(extra_pts);
    if (it == c->points.end() || ie == c->points.end()) return fail(VMSM_ERR_INVALID, "invalid points handle");
    if (off > it->second.n || n > it->second.n - off) return fail(VMSM_ERR_INVALID, "Not enough generators.");
    if (extra_off > ie->second.n || n_extra > ie->second.n - extra_off)
        return fail(VMSM_ERR_INVALID, "extra range out of bounds");
    uint64_t tot = n + n_extra;
    if (!out_affine || (tot && !scalars_le32)) return fail(VMSM_ERR_INVALID, "null argument");
    int32_t rc = ensure_stage(c, tot ? tot : 1);
    if (rc) return rc;
    if (tot) CU(cudaMemcpyAsync(c->stage_scalars, scalars_le32, tot * 32, cudaMemcpyHostToDevice, c->stream));
    CU(cudaEventRecord(c->ev_sort_in, c->stream));
    c->scalars_ready = c->ev_sort_in;
    if (it->second.curve != ie->second.curve) return fail(VMSM_ERR_INVALID, "point vectors are on different curves");
    if (it->second.curve == VMSM_CURVE_ED25519)
        rc = run_msm_ps(c, it->second, off, c->stage_scalars, tot, kSlots - 1, &ie->second, extra_off, (uint32_t)n_extra);
    else
        rc = w_run_msm_any(c, it->second, off, c->stage_scalars, tot, kSlots - 1, &ie->second, extra_off, (uint32_t)n_extra);
    if (rc) return rc;
    return fetch_slot(c, kSlots - 1, out_affine);
}

int32_t vmsm_points_concat(uint64_t ctx, uint64_t a, uint64_t a_off, uint64_t a_n, uint64_t b, uint64_t b_off,
                           uint64_t b_n, uint64_t *out) {
    GET_CTX(ctx);
    if (!out) return fail(VMSM_ERR_INVALID, "null argument");
    auto ia = c->points.find(a);
    if (ia == c->points.end()) return fail(VMSM_ERR_INVALID, "invalid points handle");
    if (a_off > ia->second.n || a_n > ia->second.n - a_off) return fail(VMSM_ERR_INVALID, "range out of bounds");
    PointSet pb{};
    if (b_n) {
        auto ib = c->points.find(b);
        if (ib == c->points.end()) return fail(VMSM_ERR_INVALID, "invalid points handle");
        if (b_off > ib->second.n || b_n > ib->second.n - b_off) return fail(VMSM_ERR_INVALID, "range out of bounds");
        pb = ib->second;
    }
    PointSet ps;
    if (b_n && pb.curve != ia->second.curve) return fail(VMSM_ERR_INVALID, "point vectors are on different curves");
    if (ia->second.curve != VMSM_CURVE_ED25519) {
        const bool g2 = ia->second.curve == VMSM_CURVE_BN256_G2;
        int32_t rcw = g2 ? w_new_pointset<Fp2BN>(ia->second.curve, a_n + b_n, &ps) : w_new_pointset<FpBN>(ia->second.curve, a_n + b_n, &ps);
        if (rcw) return rcw;
        size_t wb = wire_bytes(ia->second.curve);
        cudaError_t e = cudaSuccess;
        if (a_n) {
            e = cudaMemcpyAsync(ps.w_wire, (const uint8_t *)ia->second.w_wire + a_off * wb, a_n * wb, cudaMemcpyDeviceToDevice, c->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(ps.w_base, (const uint8_t *)ia->second.w_base + a_off * wb, a_n * wb, cudaMemcpyDeviceToDevice, c->stream);
        }
        if (b_n && e == cudaSuccess) {
            e = cudaMemcpyAsync((uint8_t *)ps.w_wire + a_n * wb, (const uint8_t *)pb.w_wire + b_off * wb, b_n * wb, cudaMemcpyDeviceToDevice, c->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync((uint8_t *)ps.w_base + a_n * wb, (const uint8_t *)pb.w_base + b_off * wb, b_n * wb, cudaMemcpyDeviceToDevice, c->stream);
        }
        if (e != cudaSuccess) {
            pool_free(ps.w_wire), pool_free(ps.w_base);
            return fail(VMSM_ERR_CUDA, "concat: %s", cudaGetErrorString(e));
        }
        uint64_t idw = c->next_id++;
        c->points[idw] = ps;
        *out = idw;
        return VMSM_OK;
    }
    int32_t rc = new_pointset(c, ia->second.curve, a_n + b_n, &ps);
    if (rc) return rc;
    CudaBE be(c);
    if (a_n) {
        KCopyPoints k = {ia->second.aff + a_off, ia->second.niels + a_off, ps.aff, ps.niels};
        be.launch(k, (uint32_t)a_n);
    }
    if (b_n) {
        KCopyPoints k = {pb.aff + b_off, pb.niels + b_off, ps.aff + a_n, ps.niels + a_n};
        be.launch(k, (uint32_t)b_n);
    }
    if (be.err != cudaSuccess) {
        pool_free(ps.aff), pool_free(ps.niels);
        return fail(VMSM_ERR_CUDA, "concat: %s", cudaGetErrorString(be.err));
    }
    uint64_t id = c->next_id++;
    c->points[id] = ps;
    *out = id;
    return VMSM_OK;
}

// ---- multi-GPU: index-range split of one MSM, partials through a peer mailbox
int32_t vmsm_mailbox_create(uint64_t ctx, uint32_t world, uint8_t *ipc_handle_out) {
    GET_CTX(ctx);
    if (world < 1 || world > 64) return fail(VMSM_ERR_INVALID, "world must be in [1, 64]");
    if (c->mailbox) return fail(VMSM_ERR_INVALID, "context already has a mailbox");
    size_t bytes = (size_t)kSlots * world * sizeof(MailSlot);
    CU(cudaMalloc(&c->mailbox, bytes));
    CU(cudaMemset(c->mailbox, 0, bytes));
    c->mb_owner = true;
    c->mb_world = world;
    c->mb_rank = 0;
    if (ipc_handle_out) {
        cudaIpcMemHandle_t h;
        CU(cudaIpcGetMemHandle(&h, c->mailbox));
        static_assert(sizeof(h) == 64, "IPC handle size");
        memcpy(ipc_handle_out, &h, 64);
    }
    return VMSM_OK;
}

int32_t vmsm_mailbox_open_ipc(uint64_t ctx, const uint8_t *ipc_handle, uint32_t rank, uint32_t world) {
    GET_CTX(ctx);
    if (!ipc_handle || rank == 0 || rank >= world || world > 64) return fail(VMSM_ERR_INVALID, "bad mailbox arguments");
    if (c->mailbox) return fail(VMSM_ERR_INVALID, "context already has a mailbox");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, 64);
    void *p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->mailbox = (MailSlot *)p;
    c->mb_owner = false;
    c->mb_ipc = true;
    c->mb_world = world;
    c->mb_rank = rank;
    return VMSM_OK;
}

int32_t vmsm_mailbox_open_local(uint64_t ctx, uint64_t owner_ctx, uint32_t rank) {
    GET_CTX(ctx);
    Ctx *o = get_ctx(owner_ctx);
    if (!o || !o->mailbox || !o->mb_owner) return fail(VMSM_ERR_INVALID, "owner context has no mailbox");
    if (rank == 0 || rank >= o->mb_world) return fail(VMSM_ERR_INVALID, "bad rank");
    if (c->mailbox) return fail(VMSM_ERR_INVALID, "context already has a mailbox");
    if (o->device != c->device) {
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, c->device, o->device));
        if (!can) return fail(VMSM_ERR_UNSUPPORTED, "no peer access from device %d to %d", c->device, o->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            return fail(VMSM_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
        cudaGetLastError();
    }
    c->mailbox = o->mailbox;
    c->mb_owner = false;
    c->mb_ipc = false;
    c->mb_world = o->mb_world;
    c->mb_rank = rank;
    return VMSM_OK;
}

int32_t vmsm_msm_dev_shard(uint64_t ctx, uint64_t pts, uint64_t poff, uint64_t n, uint64_t sc, uint64_t soff,
                           uint32_t slot, uint32_t seq) {
    GET_CTX(ctx);
    if (!c->mailbox) return fail(VMSM_ERR_INVALID, "no mailbox: call vmsm_mailbox_create / _open first");
    if (seq == 0) return fail(VMSM_ERR_INVALID, "seq must be non-zero and increase from call to call");
    auto it = c->points.find(pts);
    if (it == c->points.end()) return fail(VMSM_ERR_INVALID, "invalid points handle");
    auto is = c->scalars.find(sc);
    if (is == c->scalars.end()) return fail(VMSM_ERR_INVALID, "invalid scalars handle");
    if (poff > it->second.n || n > it->second.n - poff) return fail(VMSM_ERR_INVALID, "Not enough generators.");
    if (soff > is->second.n || n > is->second.n - soff) return fail(VMSM_ERR_INVALID, "scalar range out of bounds");
    if (slot >= kSlots - 1) return fail(VMSM_ERR_INVALID, "slot must be < %u", kSlots - 1);
    if (it->second.curve != VMSM_CURVE_ED25519) return fail(VMSM_ERR_UNSUPPORTED, "sharded MSM: Ed25519 only");
    c->shard_next = seq;
    return run_msm_ps(c, it->second, poff, is->second.data + soff * 8, n, slot);
}

int32_t vmsm_msm_async(uint64_t ctx, uint64_t pts, uint64_t off, uint64_t n, const uint8_t *scalars_le32,
                       uint32_t slot) {
    GET_CTX(ctx);
    auto it = c->points.find(pts);
    if (it == c->points.end()) return fail(VMSM_ERR_INVALID, "invalid points handle");
    if (off > it->second.n || n > it->second.n - off) return fail(VMSM_ERR_INVALID, "Not enough generators.");
    if (n && !scalars_le32) return fail(VMSM_ERR_INVALID, "null argument");
    if (slot >= kSlots - 1) return fail(VMSM_ERR_INVALID, "slot must be < %u", kSlots - 1);
    const int b = (int)(c->async_seq++ & 1);
    if ((n ? n : 1) > c->astage_cap[b]) {
        if (c->astage[b]) cudaFree(c->astage[b]);  // synchronises the device
        c->astage[b] = nullptr;
        c->astage_cap[b] = 0;
        CU(cudaMalloc(&c->astage[b], (n ? n : 1) * 32));
        c->astage_cap[b] = n ? n : 1;
        c->astage_used[b] = false;
    }
    // the staging buffer may still be read by the MSM issued two calls ago
    if (c->astage_used[b]) CU(cudaStreamWaitEvent(c->copy, c->ev_consumed[b], 0));
    if (n) CU(cudaMemcpyAsync(c->astage[b], scalars_le32, n * 32, cudaMemcpyHostToDevice, c->copy));
    CU(cudaEventRecord(c->ev_copied[b], c->copy));
    if (c->async_sort) c->scalars_ready = c->ev_copied[b];  // only the sort stream reads the scalars
    else CU(cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));
    int32_t rc = it->second.curve == VMSM_CURVE_ED25519 ? run_msm_ps(c, it->second, off, c->astage[b], n, slot)
                                                        : w_run_msm_any(c, it->second, off, c->astage[b], n, slot);
    if (rc) return rc;
    // the scalars are read by the counting sort only: the staging buffer is free again once that is done
    CU(cudaEventRecord(c->ev_consumed[b], c->async_sort ? c->sort : c->stream));
    c->astage_used[b] = true;
    return VMSM_OK;
}

int32_t vmsm_msm_dev(uint64_t ctx, uint64_t pts, uint64_t poff, uint64_t n, uint64_t sc, uint64_t soff,
                     uint32_t slot) {
    GET_CTX(ctx);
    auto it = c->points.find(pts);
    if (it == c->points.end()) return fail(VMSM_ERR_INVALID, "invalid points handle");
    auto is = c->scalars.find(sc);
    if (is == c->scalars.end()) return fail(VMSM_ERR_INVALID, "invalid scalars handle");
    if (poff > it->second.n || n > it->second.n - poff) return fail(VMSM_ERR_INVALID, "Not enough generators.");
    if (soff > is->second.n || n > is->second.n - soff) return fail(VMSM_ERR_INVALID, "scalar range out of bounds");
    if (slot >= kSlots - 1) return fail(VMSM_ERR_INVALID, "slot must be < %u", kSlots - 1);
    // the sort stream reads the scalars in place: order it after a fold kernel that may have just written them
    if (c->sc_dirty && c->async_sort && !c->scalars_ready) c->scalars_ready = c->ev_sc_written;
    if (it->second.curve != VMSM_CURVE_ED25519) return w_run_msm_any(c, it->second, poff, is->second.data + soff * 8, n, slot);
    return run_msm_ps(c, it->second, poff, is->second.data + soff * 8, n, slot);
}

// ---- device-resident scalar vectors modulo the Ed25519 group order (witness / linear-form halving)
static bool scalar_below_l(const uint8_t *le32, scl *out) {
    memcpy(out->v, le32, 32);
    return scl_gt(scl_l(), *out);
}

// a kernel on the main stream is about to overwrite scalars that earlier MSMs may still be reading on the copy /
// sort streams: order it after them
static void scalars_write_barrier(Ctx *c) {
    for (int k = 0; k < 2; k++) {
        if (c->astage_used[k]) cudaStreamWaitEvent(c->stream, c->ev_copied[k], 0);
        cudaStreamWaitEvent(c->stream, c->ev_sorted[k], 0);  // no-op until first recorded
    }
}

int32_t vmsm_scalars_fold(uint64_t ctx, uint64_t sc, uint64_t half, const uint8_t *c_le32, int32_t mode) {
    GET_CTX(ctx);
    auto it = c->scalars.find(sc);
    if (it == c->scalars.end()) return fail(VMSM_ERR_INVALID, "invalid scalars handle");
    if (!c_le32) return fail(VMSM_ERR_INVALID, "null argument");
    if (mode != VMSM_FOLD_WITNESS && mode != VMSM_FOLD_FORM) return fail(VMSM_ERR_INVALID, "unknown fold mode %d", mode);
    if (half == 0 || 2 * half > it->second.n) return fail(VMSM_ERR_INVALID, "fold: need 2*half <= length");
    scl cs;
    if (!scalar_below_l(c_le32, &cs)) return fail(VMSM_ERR_INVALID, "challenge is not reduced modulo the group order");
    scalars_write_barrier(c);
    CudaBE be(c);
    KScalarAxpy k = {it->second.data, it->second.data + 8ull * half, scl_to_mont(cs), mode};
    be.launch(k, (uint32_t)half);
    be.note(cudaEventRecord(c->ev_sc_written, c->stream));
    c->sc_dirty = true;
    if (be.err != cudaSuccess) return fail(VMSM_ERR_CUDA, "scalars_fold: %s", cudaGetErrorString(be.err));
    return VMSM_OK;
}

int32_t vmsm_scalars_axpy(uint64_t ctx, uint64_t dst, uint64_t doff, uint64_t src, uint64_t soff, uint64_t n,
                          const uint8_t *c_le32, int32_t mode) {
    GET_CTX(ctx);
    auto id = c->scalars.find(dst);
    if (id == c->scalars.end()) return fail(VMSM_ERR_INVALID, "invalid scalars handle");
    if (!c_le32) return fail(VMSM_ERR_INVALID, "null argument");
    if (mode < VMSM_AXPY_ADD_SCALED || mode > VMSM_AXPY_SCALE) return fail(VMSM_ERR_INVALID, "unknown axpy mode %d", mode);
    if (doff > id->second.n || n > id->second.n - doff) return fail(VMSM_ERR_INVALID, "scalar range out of bounds");
    const uint32_t *sp = nullptr;
    if (mode != VMSM_AXPY_SCALE) {
        auto is = c->scalars.find(src);
        if (is == c->scalars.end()) return fail(VMSM_ERR_INVALID, "invalid scalars handle");
        if (soff > is->second.n || n > is->second.n - soff) return fail(VMSM_ERR_INVALID, "scalar range out of bounds");
        if (src == dst && soff < doff + n && doff < soff + n && soff != doff + n && doff != soff + n)
            return fail(VMSM_ERR_INVALID, "axpy: source and destination ranges overlap");
        sp = is->second.data + soff * 8;
    }
    scl cs;
    if (!scalar_below_l(c_le32, &cs)) return fail(VMSM_ERR_INVALID, "constant is not reduced modulo the group order");
    if (!n) return VMSM_OK;
    if (n > (1ull << 28)) return fail(VMSM_ERR_UNSUPPORTED, "too many scalars");
    scalars_write_barrier(c);
    CudaBE be(c);
    KScalarAxpy k = {id->second.data + doff * 8, sp, scl_to_mont(cs), mode};
    be.launch(k, (uint32_t)n);
    be.note(cudaEventRecord(c->ev_sc_written, c->stream));
    c->sc_dirty = true;
    if (be.err != cudaSuccess) return fail(VMSM_ERR_CUDA, "scalars_axpy: %s", cudaGetErrorString(be.err));
    return VMSM_OK;
}

// <a, b> mod l into the 32 bytes at `dst` (device), on the main stream: three latency-bound stages (T threads multiply
// and pre-sum n/T terms, T2 threads sum T/T2 of those, one thread finishes) with serial chains of comparable length
static void launch_dot(Ctx *c, CudaBE &be, const uint32_t *a, const uint32_t *b, uint64_t n, uint32_t *dst) {
    uint32_t T, T2;
    dot_stage_sizes(n, &T, &T2);
    uint32_t *p1 = c->dot_scratch, *p2 = p1 + kDotT1 * 8;
    KScalarDotPartial k1 = {a, b, (uint32_t)n, T, p1};
    be.launch(k1, T);
    KScalarSum k2 = {p1, T, T2, p2, 0};
    be.launch(k2, T2);
    KScalarSum k3 = {p2, T2, 1, dst, 1};
    be.launch(k3, 1);
}

int32_t vmsm_scalars_dot(uint64_t ctx, uint64_t a, uint64_t aoff, uint64_t b, uint64_t boff, uint64_t n,
                         uint8_t *out_le32) {
    GET_CTX(ctx);
    auto ia = c->scalars.find(a), ib = c->scalars.find(b);
    if (ia == c->scalars.end() || ib == c->scalars.end()) return fail(VMSM_ERR_INVALID, "invalid scalars handle");
    if (!out_le32) return fail(VMSM_ERR_INVALID, "null argument");
    if (aoff > ia->second.n || n > ia->second.n - aoff || boff > ib->second.n || n > ib->second.n - boff)
        return fail(VMSM_ERR_INVALID, "scalar range out of bounds");
    if (n > (1ull << 28)) return fail(VMSM_ERR_UNSUPPORTED, "too many scalars");
    memset(out_le32, 0, 32);
    if (!n) return VMSM_OK;
    CudaBE be(c);
    uint32_t *p3 = c->dot_scratch + (kDotT1 + kDotT2) * 8;
    launch_dot(c, be, ia->second.data + aoff * 8, ib->second.data + boff * 8, n, p3);
    be.note(cudaMemcpyAsync(c->pin, p3, 32, cudaMemcpyDeviceToHost, c->stream));
    be.note(cudaStreamSynchronize(c->stream));
    if (be.err != cudaSuccess) return fail(VMSM_ERR_CUDA, "scalars_dot: %s", cudaGetErrorString(be.err));
    memcpy(out_le32, c->pin, 32);
    return VMSM_OK;
}

int32_t vmsm_scalars_text_ptr(uint64_t ctx, uint64_t sc, uint64_t off, uint64_t n, int32_t is_signed,
                              const uint8_t **text, uint64_t *len) {
    GET_CTX(ctx);
    if (!len || !text) return fail(VMSM_ERR_INVALID, "null argument");
    auto it = c->scalars.find(sc);
    if (it == c->scalars.end()) return fail(VMSM_ERR_INVALID, "invalid scalars handle");
    if (off > it->second.n || n > it->second.n - off) return fail(VMSM_ERR_INVALID, "scalar range out of bounds");
    *len = 0;
    *text = nullptr;
    if (n > (1ull << 26)) return fail(VMSM_ERR_UNSUPPORTED, "too many scalars");
    int32_t rc = text_ensure(c, n ? n : 1);
    if (rc) return rc;
    *text = c->txt_host_sc;
    if (!n) return VMSM_OK;
    CudaBE be(c);
    KScalarText kt = {it->second.data + off * 8, c->txt_slots, c->txt_lens, (uint32_t)n, is_signed};
    be.launch(kt, (uint32_t)n);
    return text_finish(c, be, n, VMSM_SCALAR_TEXT_SLOT, len, c->txt_host_sc);
}

int32_t vmsm_msm_dev_ext(uint64_t ctx, uint64_t pts, uint64_t poff, uint64_t n, uint64_t sc, uint64_t soff,
                         uint64_t extra_pts, uint64_t extra_off, uint64_t n_extra, const uint8_t *extra_scalars_le32,
                         uint32_t slot) {
    GET_CTX(ctx);
    auto it = c->points.find(pts);
    auto ie = c->points.find(extra_pts);
    if (it == c->points.end() || ie == c->points.end()) return fail(VMSM_ERR_INVALID, "invalid points handle");
    auto is = c->scalars.find(sc);
    if (is == c->scalars.end()) return fail(VMSM_ERR_INVALID, "invalid scalars handle");
    if (it->second.curve != ie->second.curve) return fail(VMSM_ERR_INVALID, "msm_dev_ext: the two point vectors are on different curves");
    if (poff > it->second.n || n > it->second.n - poff) return fail(VMSM_ERR_INVALID, "Not enough generators.");
    if (soff > is->second.n || n > is->second.n - soff) return fail(VMSM_ERR_INVALID, "scalar range out of bounds");
    if (extra_off > ie->second.n || n_extra > ie->second.n - extra_off)
        return fail(VMSM_ERR_INVALID, "extra range out of bounds");
    if (n_extra && !extra_scalars_le32) return fail(VMSM_ERR_INVALID, "null argument");
    if (n_extra > 64) return fail(VMSM_ERR_INVALID, "at most 64 extra terms");
    if (slot >= kSlots - 1) return fail(VMSM_ERR_INVALID, "slot must be < %u", kSlots - 1);
    const uint64_t tot = n + n_extra;
    const int b = (int)(c->async_seq++ & 1);
    if ((tot ? tot : 1) > c->astage_cap[b]) {
        if (c->astage[b]) cudaFree(c->astage[b]);  // synchronises the device
        c->astage[b] = nullptr;
        c->astage_cap[b] = 0;
        CU(cudaMalloc(&c->astage[b], (tot ? tot : 1) * 32));
        c->astage_cap[b] = tot ? tot : 1;
        c->astage_used[b] = false;
    }
    // staging on the copy stream: after the kernel that last wrote the device scalars, and after the MSM that last
    // read this staging buffer
    if (c->sc_dirty) CU(cudaStreamWaitEvent(c->copy, c->ev_sc_written, 0));
    if (c->astage_used[b]) CU(cudaStreamWaitEvent(c->copy, c->ev_consumed[b], 0));
    if (n) CU(cudaMemcpyAsync(c->astage[b], is->second.data + soff * 8, n * 32, cudaMemcpyDeviceToDevice, c->copy));
    if (n_extra) CU(cudaMemcpyAsync(c->astage[b] + n * 8, extra_scalars_le32, n_extra * 32, cudaMemcpyHostToDevice, c->copy));
    CU(cudaEventRecord(c->ev_copied[b], c->copy));
    if (c->async_sort) c->scalars_ready = c->ev_copied[b];
    else CU(cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));
    // extra terms that simply continue the main range of the same vector (the delta terms a Pinocchio key keeps right
    // after its mid-wire elements, pynocchio.py:248-262) make one plain range, which also keeps the vector's table usable
    const bool contiguous = extra_pts == pts && extra_off == poff + n;
    PointSet *eps = contiguous ? nullptr : &ie->second;
    const uint32_t nx = contiguous ? 0u : (uint32_t)n_extra;
    int32_t rc = it->second.curve == VMSM_CURVE_ED25519 ? run_msm_ps(c, it->second, poff, c->astage[b], tot, slot, eps, extra_off, nx)
                                                        : w_run_msm_any(c, it->second, poff, c->astage[b], tot, slot, eps, extra_off, nx);
    if (rc) return rc;
    CU(cudaEventRecord(c->ev_consumed[b], c->async_sort ? c->sort : c->stream));
    c->astage_used[b] = true;
    return VMSM_OK;
}

// One commitment of a folding round without a host round trip (compressed_pivot.py:41-42: A_i = g_R^{z_L} k^{L_R(z_L)}):
// sum_{i<n} s[soff+i] P[poff+i]  +  <a[aoff..], b[boff..]> * E[extra_off], the inner product computed on the device and
// handed to the MSM as the scalar of its extra term.
int32_t vmsm_msm_dev_ext_dot(uint64_t ctx, uint64_t pts, uint64_t poff, uint64_t n, uint64_t sc, uint64_t soff,
                             uint64_t extra_pts, uint64_t extra_off, uint64_t dot_a, uint64_t dot_aoff, uint64_t dot_b,
                             uint64_t dot_boff, uint64_t dot_n, uint32_t slot) {
    GET_CTX(ctx);
    auto it = c->points.find(pts);
    auto ie = c->points.find(extra_pts);
    if (it == c->points.end() || ie == c->points.end()) return fail(VMSM_ERR_INVALID, "invalid points handle");
    auto is = c->scalars.find(sc), ia = c->scalars.find(dot_a), ib = c->scalars.find(dot_b);
    if (is == c->scalars.end() || ia == c->scalars.end() || ib == c->scalars.end())
        return fail(VMSM_ERR_INVALID, "invalid scalars handle");
    if (it->second.curve != VMSM_CURVE_ED25519 || ie->second.curve != VMSM_CURVE_ED25519)
        return fail(VMSM_ERR_UNSUPPORTED, "msm_dev_ext_dot: Ed25519 only");
    if (poff > it->second.n || n > it->second.n - poff) return fail(VMSM_ERR_INVALID, "Not enough generators.");
    if (soff > is->second.n || n > is->second.n - soff) return fail(VMSM_ERR_INVALID, "scalar range out of bounds");
    if (extra_off >= ie->second.n) return fail(VMSM_ERR_INVALID, "extra range out of bounds");
    if (dot_aoff > ia->second.n || dot_n > ia->second.n - dot_aoff || dot_boff > ib->second.n || dot_n > ib->second.n - dot_boff)
        return fail(VMSM_ERR_INVALID, "scalar range out of bounds");
    if (dot_n == 0 || dot_n > (1ull << 28)) return fail(VMSM_ERR_INVALID, "inner product length out of range");
    if (slot >= kSlots - 1) return fail(VMSM_ERR_INVALID, "slot must be < %u", kSlots - 1);
    const uint64_t tot = n + 1;
    const int b = (int)(c->async_seq++ & 1);
    if (tot > c->astage_cap[b]) {
        if (c->astage[b]) cudaFree(c->astage[b]);  // synchronises the device
        c->astage[b] = nullptr;
        c->astage_cap[b] = 0;
        CU(cudaMalloc(&c->astage[b], tot * 32));
        c->astage_cap[b] = tot;
        c->astage_used[b] = false;
    }
    // the inner product on the main stream (after the kernels that last wrote the vectors), into this parity's result word
    CudaBE be(c);
    uint32_t *dres = c->dot_scratch + (kDotT1 + kDotT2 + 1 + (uint32_t)b) * 8;
    launch_dot(c, be, ia->second.data + dot_aoff * 8, ib->second.data + dot_boff * 8, dot_n, dres);
    if (be.err != cudaSuccess) return fail(VMSM_ERR_CUDA, "msm_dev_ext_dot: %s", cudaGetErrorString(be.err));
    CU(cudaEventRecord(c->ev_dot[b], c->stream));
    // staging on the copy stream: after the inner product, after the kernel that last wrote the device scalars, and
    // after the MSM that last read this staging buffer
    CU(cudaStreamWaitEvent(c->copy, c->ev_dot[b], 0));
    if (c->sc_dirty) CU(cudaStreamWaitEvent(c->copy, c->ev_sc_written, 0));
    if (c->astage_used[b]) CU(cudaStreamWaitEvent(c->copy, c->ev_consumed[b], 0));
    if (n) CU(cudaMemcpyAsync(c->astage[b], is->second.data + soff * 8, n * 32, cudaMemcpyDeviceToDevice, c->copy));
    CU(cudaMemcpyAsync(c->astage[b] + n * 8, dres, 32, cudaMemcpyDeviceToDevice, c->copy));
    CU(cudaEventRecord(c->ev_copied[b], c->copy));
    if (c->async_sort) c->scalars_ready = c->ev_copied[b];
    else CU(cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));
    int32_t rc = run_msm_ps(c, it->second, poff, c->astage[b], tot, slot, &ie->second, extra_off, 1);
    if (rc) return rc;
    CU(cudaEventRecord(c->ev_consumed[b], c->async_sort ? c->sort : c->stream));
    c->astage_used[b] = true;
    return VMSM_OK;
}

int32_t vmsm_result_affine(uint64_t ctx, uint32_t slot, uint8_t *out_affine) {
    GET_CTX(ctx);
    if (slot >= kSlots || !out_affine) return fail(VMSM_ERR_INVALID, "bad slot / null argument");
    return fetch_slot(c, slot, out_affine);  // waits for THIS result only, not for MSMs issued after it
}

int32_t vmsm_result_extended(uint64_t ctx, uint32_t slot, uint8_t *out_extended) {
    GET_CTX(ctx);
    if (slot >= kSlots || !out_extended) return fail(VMSM_ERR_INVALID, "bad slot / null argument");
    if (c->slot_curve[slot] != VMSM_CURVE_ED25519) return fail(VMSM_ERR_UNSUPPORTED, "extended results: Ed25519 only");
    CU(join_tail(c));
    CU(cudaMemcpyAsync(c->pin, c->res_ext + slot, sizeof(ge_ext), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(out_extended, c->pin, sizeof(ge_ext));
    return VMSM_OK;
}

// ---- fold
int32_t vmsm_fold(uint64_t ctx, uint64_t pts, uint64_t half, const uint8_t *c_le32) {
    GET_CTX(ctx);
    auto it = c->points.find(pts);
    if (it == c->points.end()) return fail(VMSM_ERR_INVALID, "invalid points handle");
    if (!c_le32) return fail(VMSM_ERR_INVALID, "null argument");
    if (it->second.curve != VMSM_CURVE_ED25519) return fail(VMSM_ERR_UNSUPPORTED, "fold is defined for the Ed25519 path only");
    if (half == 0 || 2 * half > it->second.n) return fail(VMSM_ERR_INVALID, "fold: need 2*half <= length");
    int32_t rc = ensure_tmp(c, half);
    if (rc) return rc;
    uint32_t cs[8];
    memcpy(cs, c_le32, 32);
    CU(wait_bases_released(c, false));  // the fold rewrites niels in place: order it after every reader in flight
    if (it->second.pre) {  // a table describes the generators it was built from, not the folded ones
        CU(wait_bases_released(c, true));
        pool_free(it->second.pre);
        it->second.pre = nullptr;
        it->second.pre_c = it->second.pre_W = 0;
    }
    CudaBE be(c);
    fold_run(be, it->second.aff, it->second.niels, c->tmp_ext, (uint32_t)half, cs, c->fold_quad_max);
    if (be.err != cudaSuccess) return fail(VMSM_ERR_CUDA, "fold: %s", cudaGetErrorString(be.err));
    it->second.n = half;
    return VMSM_OK;
}

// ---- small linear combination of host points
int32_t vmsm_lincomb(uint64_t ctx, int32_t curve, const uint8_t *affine, const uint8_t *scalars_le32, uint64_t n,
                     uint8_t *out_affine) {
    GET_CTX(ctx);
    if (curve < VMSM_CURVE_ED25519 || curve > VMSM_CURVE_BN256_G2) return fail(VMSM_ERR_UNSUPPORTED, "unknown curve %d", curve);
    if (n > 64) return fail(VMSM_ERR_INVALID, "lincomb takes at most 64 terms");
    if (!out_affine || (n && (!affine || !scalars_le32))) return fail(VMSM_ERR_INVALID, "null argument");
    int32_t rc = ensure_stage(c, 64);
    if (rc) return rc;
    CudaBE be(c);
    if (curve != VMSM_CURVE_ED25519) {
        size_t wb = wire_bytes(curve);
        if (n) {
            be.note(cudaMemcpyAsync(c->small_w_wire, affine, n * wb, cudaMemcpyHostToDevice, c->stream));
            be.note(cudaMemcpyAsync(c->stage_scalars, scalars_le32, n * 32, cudaMemcpyHostToDevice, c->stream));
            be.zero(c->err_word, 4);
            if (curve == VMSM_CURVE_BN256_G1) {
                KUploadW<FpBN> k = {(const waff<FpBN> *)c->small_w_wire, (waff<FpBN> *)c->small_w_base, c->err_word, c->check_points ? 1u : 0u};
                be.launch(k, (uint32_t)n);
            } else {
                KUploadW<Fp2BN> k = {(const waff<Fp2BN> *)c->small_w_wire, (waff<Fp2BN> *)c->small_w_base, c->err_word, c->check_points ? 1u : 0u};
                be.launch(k, (uint32_t)n);
            }
        }
        if (be.err != cudaSuccess) return fail(VMSM_ERR_CUDA, "lincomb: %s", cudaGetErrorString(be.err));
        PointSet tmp{};
        tmp.curve = curve;
        tmp.n = n;
        tmp.w_wire = c->small_w_wire;
        tmp.w_base = c->small_w_base;
        CU(cudaEventRecord(c->ev_sort_in, c->stream));
        c->scalars_ready = c->ev_sort_in;
        rc = w_run_msm_any(c, tmp, 0, c->stage_scalars, n, kSlots - 1);
        if (rc) return rc;
        CU(cudaMemcpyAsync(c->pin + 64, c->err_word, 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (n && *reinterpret_cast<uint32_t *>(c->pin + 64)) return fail(VMSM_ERR_POINT, "invalid point in lincomb input");
        return fetch_slot(c, kSlots - 1, out_affine);
    }
    if (n) {
        be.note(cudaMemcpyAsync(c->small_aff, affine, n * sizeof(ge_aff), cudaMemcpyHostToDevice, c->stream));
        be.note(cudaMemcpyAsync(c->stage_scalars, scalars_le32, n * 32, cudaMemcpyHostToDevice, c->stream));
        be.zero(c->err_word, 4);
        KAffToNiels k = {c->small_aff, c->small_niels, c->err_word, c->check_points ? 1u : 0u};
        be.launch(k, (uint32_t)n);
    }
    if (be.err != cudaSuccess) return fail(VMSM_ERR_CUDA, "lincomb: %s", cudaGetErrorString(be.err));
    CU(cudaEventRecord(c->ev_sort_in, c->stream));
    c->scalars_ready = c->ev_sort_in;
    rc = run_msm(c, c->small_niels, c->stage_scalars, n, kSlots - 1);
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->pin + 64, c->err_word, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (n && *reinterpret_cast<uint32_t *>(c->pin + 64)) {
        CU(cudaEventSynchronize(c->ev_slot[kSlots - 1]));
        return fail(VMSM_ERR_POINT, "invalid point in lincomb input");
    }
    return fetch_slot(c, kSlots - 1, out_affine);
    return VMSM_OK;
}

int32_t vmsm_lincomb_async(uint64_t ctx, int32_t curve, const uint8_t *affine, const uint8_t *scalars_le32, uint64_t n,
                           uint32_t slot) {
    GET_CTX(ctx);
    if (curve != VMSM_CURVE_ED25519) return fail(VMSM_ERR_UNSUPPORTED, "lincomb_async: Ed25519 only");
    if (n > 64) return fail(VMSM_ERR_INVALID, "lincomb takes at most 64 terms");
    if (n && (!affine || !scalars_le32)) return fail(VMSM_ERR_INVALID, "null argument");
    if (slot >= kSlots - 1) return fail(VMSM_ERR_INVALID, "slot must be < %u", kSlots - 1);
    const uint32_t r = c->la_seq++ % kLaRing;
    ge_aff *aff = c->la_aff + 64 * r;
    ge_niels *niels = c->la_niels + 64 * r;
    uint32_t *sc = c->la_scalars + 64 * 8 * r, *err = c->la_err + r;
    CudaBE be(c);
    // the staging set may still be read by the sort of the call that used it kLaRing calls ago (its points are read
    // by that call's accumulate kernel, which precedes this copy on the main stream anyway)
    if (c->la_used[r]) be.note(cudaStreamWaitEvent(c->stream, c->ev_la_sorted[r], 0));
    // ... and its points by that call's accumulate kernel, which runs on a head stream
    if (c->la_used[r]) be.note(wait_bases_released(c, false));
    be.zero(err, 4);
    if (n) {
        be.note(cudaMemcpyAsync(aff, affine, n * sizeof(ge_aff), cudaMemcpyHostToDevice, c->stream));
        be.note(cudaMemcpyAsync(sc, scalars_le32, n * 32, cudaMemcpyHostToDevice, c->stream));
        KAffToNiels k = {aff, niels, err, c->check_points ? 1u : 0u};
        be.launch(k, (uint32_t)n);
        KPublishErr kp = {err, c->res_status_host + slot, 2u};
        be.launch(kp, 32);
    }
    if (be.err != cudaSuccess) return fail(VMSM_ERR_CUDA, "lincomb_async: %s", cudaGetErrorString(be.err));
    CU(cudaEventRecord(c->ev_sort_in, c->stream));
    c->scalars_ready = c->ev_sort_in;
    int32_t rc = run_msm(c, niels, sc, n, slot);
    if (rc) return rc;
    CU(cudaEventRecord(c->ev_la_sorted[r], c->async_sort ? c->sort : c->stream));
    c->la_used[r] = true;
    return VMSM_OK;
}

int32_t vmsm_host_alloc(uint64_t bytes, void **ptr) {
    if (!ptr) return fail(VMSM_ERR_INVALID, "null argument");
    *ptr = nullptr;
    CU(cudaHostAlloc(ptr, bytes ? bytes : 16, cudaHostAllocDefault));
    return VMSM_OK;
}
int32_t vmsm_host_free(void *ptr) {
    if (ptr) CU(cudaFreeHost(ptr));
    return VMSM_OK;
}

int32_t vmsm_selftest_fe(uint64_t ctx, int32_t op, const uint8_t *a, const uint8_t *b, uint64_t n, uint8_t *out) {
    GET_CTX(ctx);
    if (!a || !b || !out || n == 0 || n > (1u << 24)) return fail(VMSM_ERR_INVALID, "bad argument");
    fe *da = nullptr, *db = nullptr, *dout = nullptr;
    CU(cudaMalloc(&da, n * 32));
    CU(cudaMalloc(&db, n * 32));
    CU(cudaMalloc(&dout, n * 32));
    CudaBE be(c);
    be.note(cudaMemcpyAsync(da, a, n * 32, cudaMemcpyHostToDevice, c->stream));
    be.note(cudaMemcpyAsync(db, b, n * 32, cudaMemcpyHostToDevice, c->stream));
    KSelfTestFe k = {da, db, dout, op};
    be.launch(k, (uint32_t)n);
    be.note(cudaMemcpyAsync(out, dout, n * 32, cudaMemcpyDeviceToHost, c->stream));
    be.note(cudaStreamSynchronize(c->stream));
    cudaFree(da), cudaFree(db), cudaFree(dout);
    if (be.err != cudaSuccess) return fail(VMSM_ERR_CUDA, "selftest: %s", cudaGetErrorString(be.err));
    return VMSM_OK;
}

int32_t vmsm_microbench_imad(uint64_t ctx, double *tera_lp_per_s) {
    GET_CTX(ctx);
    if (!tera_lp_per_s) return fail(VMSM_ERR_INVALID, "null argument");
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, c->device));
    int blocks = prop.multiProcessorCount * 4, threads = 512;
    uint32_t *out = nullptr;
    CU(cudaMalloc(&out, (size_t)blocks * threads * 4));
    float best = 1e30f;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(c->t0, c->stream);
        vmsm_imad_peak<<<blocks, threads, 0, c->stream>>>(out, 12345u + rep, 6789u);
        cudaEventRecord(c->t1, c->stream);
        cudaError_t e = cudaEventSynchronize(c->t1);
        if (e != cudaSuccess) {
            cudaFree(out);
            return fail(VMSM_ERR_CUDA, "microbench: %s", cudaGetErrorString(e));
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, c->t0, c->t1);
        if (rep > 0 && ms < best) best = ms;  // rep 0 is the warm-up
    }
    cudaFree(out);
    double ops = (double)blocks * threads * 8.0 * MB_ITERS;
    *tera_lp_per_s = ops / (best * 1e-3) / 1e12;
    return VMSM_OK;
}

}  // extern "C"
