// Kernel bodies of the Ed25519 MSM / fold engine, written as per-thread functors.
//
// Every kernel without intra-block cooperation is a struct with `void operator()(uint32_t tid) const`; the CUDA
// backend (vmsm.cu) launches it through vmsm_kernel<F><<<...>>>, the host-emulation backend (tests/hostemu, tests
// only, never shipped) runs the very same body in a loop so that all indexing and arithmetic can be checked
// against the oracle in a container without a GPU.
//
// Pipeline (signed-window Pippenger; replaces pivot.py:139-145 vector_commitment's n independent
// double-and-add scalar multiplications):
//   KDigitsHist  scalar -> W signed c-bit digits, histogram of |digit| per (window, bucket)
//   scan         exclusive prefix sums per window -> CSR offsets              (cooperative, backend specific)
//   KScatter     second pass over the scalars: counting-sort (index|sign) into the CSR lists
//   KAccumulate  one thread per bucket: sum of its (signed) bases with 7M mixed additions, bases gathered from L2
//   KReduce      radix-R tree over buckets carrying (S, T) = (sum B_k, sum (k - lo) B_k) per node
//   KFinal       Horner over windows, one inversion, canonical affine out
#pragma once
#include "ed25519.cuh"
#include "ed25519_quad.cuh"

namespace vmsm {

#if defined(__CUDA_ARCH__)
#define VMSM_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define VMSM_ATOMIC_OR(p, v) atomicOr((p), (v))
#else
static inline uint32_t vmsm_host_atomic_add(uint32_t *p, uint32_t v) {
    uint32_t o = *p;
    *p = o + v;
    return o;
}
static inline uint32_t vmsm_host_atomic_or(uint32_t *p, uint32_t v) {
    uint32_t o = *p;
    *p = o | v;
    return o;
}
#define VMSM_ATOMIC_ADD(p, v) vmsm_host_atomic_add((p), (v))
#define VMSM_ATOMIC_OR(p, v) vmsm_host_atomic_or((p), (v))
#endif

// ---------------------------------------------------------------------------------------------- vector loads
struct alignas(16) u32x4 {
    uint32_t x, y, z, w;
};

VMSM_HD u32x4 ld128(const void *p) {
#if defined(__CUDA_ARCH__)
    uint4 t = __ldg(reinterpret_cast<const uint4 *>(p));
    u32x4 r = {t.x, t.y, t.z, t.w};
    return r;
#else
    return *reinterpret_cast<const u32x4 *>(p);
#endif
}
VMSM_HD void st128(void *p, const u32x4 &v) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint4 *>(p) = make_uint4(v.x, v.y, v.z, v.w);
#else
    *reinterpret_cast<u32x4 *>(p) = v;
#endif
}

VMSM_HD fe ld_fe(const fe *p) {
    u32x4 a = ld128(p), b = ld128(reinterpret_cast<const uint8_t *>(p) + 16);
    fe r = {{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
    return r;
}
VMSM_HD void st_fe(fe *p, const fe &v) {
    u32x4 a = {v.v[0], v.v[1], v.v[2], v.v[3]}, b = {v.v[4], v.v[5], v.v[6], v.v[7]};
    st128(p, a);
    st128(reinterpret_cast<uint8_t *>(p) + 16, b);
}
VMSM_HD ge_niels ld_niels(const ge_niels *p) {
    ge_niels r;
    r.ypx = ld_fe(&p->ypx);
    r.ymx = ld_fe(&p->ymx);
    r.t2d = ld_fe(&p->t2d);
    return r;
}
VMSM_HD void st_niels(ge_niels *p, const ge_niels &v) {
    st_fe(&p->ypx, v.ypx);
    st_fe(&p->ymx, v.ymx);
    st_fe(&p->t2d, v.t2d);
}
VMSM_HD ge_ext ld_ext(const ge_ext *p) {
    ge_ext r;
    r.X = ld_fe(&p->X);
    r.Y = ld_fe(&p->Y);
    r.Z = ld_fe(&p->Z);
    r.T = ld_fe(&p->T);
    return r;
}
// same without the read-only (non-coherent) path: for memory another GPU or stream may have just written
VMSM_HD ge_ext ld_ext_plain(const ge_ext *p) {
    ge_ext r;
#if defined(__CUDA_ARCH__)
    const volatile uint32_t *w = reinterpret_cast<const volatile uint32_t *>(p);
#pragma unroll
    for (int i = 0; i < 8; i++) r.X.v[i] = w[i], r.Y.v[i] = w[8 + i], r.Z.v[i] = w[16 + i], r.T.v[i] = w[24 + i];
#else
    r = *p;
#endif
    return r;
}
VMSM_HD void st_ext(ge_ext *p, const ge_ext &v) {
    st_fe(&p->X, v.X);
    st_fe(&p->Y, v.Y);
    st_fe(&p->Z, v.Z);
    st_fe(&p->T, v.T);
}
VMSM_HD ge_aff ld_aff(const ge_aff *p) {
    ge_aff r;
    r.x = ld_fe(&p->x);
    r.y = ld_fe(&p->y);
    return r;
}
VMSM_HD void st_aff(ge_aff *p, const ge_aff &v) {
    st_fe(&p->x, v.x);
    st_fe(&p->y, v.y);
}

// ---------------------------------------------------------------------------------------------- scalars
struct MsmGeom {
    uint32_t n;   // terms
    uint32_t c;   // window bits
    uint32_t W;   // windows, c*W >= scalar_bits + 1
    uint32_t NB;  // buckets per window = 2^(c-1), bucket b holds |digit| = b+1
    // Bucket sets.  Plain MSM: S = W, window w owns bucket set w (its sum carries the weight 2^(c*w), applied by the
    // Horner chain).  MSM over PRECOMPUTED bases (tables of 2^(c*w) * P_i, KPrecompute): the weight is already in the
    // base, so windows may share buckets -- window w goes to set w % S with k = w / S recorded in the CSR entry
    // (entry = term index | k << lg | sign << 31), S * NB buckets in all, no doublings anywhere.
    uint32_t S;   // bucket sets, 1 <= S <= W
    uint32_t lg;  // bit position of k in a CSR entry: 2^lg >= n (unused when S == W: k is always 0)
};
// position in idx where bucket set s starts: n * (number of windows in the sets before it)
VMSM_HD uint32_t geom_set_start(const MsmGeom &g, uint32_t s) {
    uint32_t q = g.W / g.S, r = g.W % g.S;
    return g.n * (s * q + (s < r ? s : r));
}

struct sc256 {
    uint32_t v[9];  // v[8] = 0 sentinel so window extraction can read one limb past the top
};

VMSM_HD sc256 ld_scalar(const uint32_t *base, uint32_t i) {
    u32x4 a = ld128(base + 8ull * i), b = ld128(base + 8ull * i + 4);
    sc256 s = {{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, 0u}};
    return s;
}

// raw c-bit window at bit offset `bit` (bit + c may run past 256: zero extended)
VMSM_HD uint32_t sc_bits(const sc256 &s, uint32_t bit, uint32_t c) {
    uint32_t limb = bit >> 5, sh = bit & 31;
    if (limb >= 8) return 0;
    uint64_t two = (uint64_t)s.v[limb] | ((uint64_t)s.v[limb + 1] << 32);
    return (uint32_t)(two >> sh) & ((1u << c) - 1u);
}

// signed digit of window w given the carry from window w-1; digits lie in [-(2^(c-1) - 1), 2^(c-1)]
VMSM_HD int32_t sc_digit(const sc256 &s, uint32_t w, uint32_t c, uint32_t &carry) {
    uint32_t raw = sc_bits(s, w * c, c) + carry;
    uint32_t half = 1u << (c - 1);
    if (raw > half) {
        carry = 1;
        return (int32_t)raw - (int32_t)(1u << c);
    }
    carry = 0;
    return (int32_t)raw;
}

// ---------------------------------------------------------------------------------------------- synthetic inputs
// Counter-based generator; the spec lives in oracle/prng.py (independent Python restatement) and DESIGN.md.
VMSM_HD uint64_t splitmix64_mix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// uniform-ish scalar below the Ed25519 group order l: 253 low bits, minus l if >= l
VMSM_HD sc256 synth_scalar_ed(uint64_t seed, uint64_t i) {
    sc256 s;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint64_t wv = splitmix64_mix(seed + 0x9E3779B97F4A7C15ull * (4 * i + j + 1));
        s.v[2 * j] = (uint32_t)wv;
        s.v[2 * j + 1] = (uint32_t)(wv >> 32);
    }
    s.v[7] &= 0x1fffffffu;
    s.v[8] = 0;
    const uint32_t L[8] = {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0u, 0u, 0u, 0x10000000u};
    uint32_t d[8];
    int64_t bw = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        bw += (int64_t)s.v[k] - (int64_t)L[k];
        d[k] = (uint32_t)bw;
        bw >>= 32;
    }
    if (bw == 0) {  // no borrow: s >= l
#pragma unroll
        for (int k = 0; k < 8; k++) s.v[k] = d[k];
    }
    return s;
}

// ---------------------------------------------------------------------------------------------- MSM kernels
struct KDigitsHist {
    enum { kBlock = 256 };
    const uint32_t *scalars;  // n x 8 limbs, each < group order
    uint32_t *counts;         // W x NB
    MsmGeom g;
    VMSM_HD void operator()(uint32_t tid) const {
        sc256 s = ld_scalar(scalars, tid);
        uint32_t carry = 0, set = 0;
        for (uint32_t w = 0; w < g.W; w++) {
            int32_t d = sc_digit(s, w, g.c, carry);
            if (d != 0) {
                uint32_t a = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
                VMSM_ATOMIC_ADD(&counts[set * g.NB + (a - 1)], 1u);
            }
            if (++set == g.S) set = 0;
        }
    }
};

struct KScatter {
    enum { kBlock = 256 };
    const uint32_t *scalars;
    uint32_t *cursor;  // W x NB, initialised to the CSR offsets; ends at offsets + counts
    uint32_t *idx;     // W x n entries: base index | sign << 31
    MsmGeom g;
    VMSM_HD void operator()(uint32_t tid) const {
        sc256 s = ld_scalar(scalars, tid);
        uint32_t carry = 0, set = 0, k = 0;
        for (uint32_t w = 0; w < g.W; w++) {
            int32_t d = sc_digit(s, w, g.c, carry);
            if (d != 0) {
                uint32_t a = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
                uint32_t pos = VMSM_ATOMIC_ADD(&cursor[set * g.NB + (a - 1)], 1u);
                idx[pos] = tid | (k << g.lg) | (d < 0 ? 0x80000000u : 0u);
            }
            if (++set == g.S) set = 0, k++;
        }
    }
};

// Block-privatised counting sort (CUDA backend: vmsm_bsort_* in vmsm.cu; restated in a loop by the host emulation).
// The two-pass sort above spends 2 x W global atomics per scalar, which is what bounds it (L2 atomic throughput, 6 % of
// the HBM roofline) and what it steals from the accumulate kernel it runs under.  Here the digits are recoded ONCE into
// a window-major array of 16-bit codes; a block then owns (bucket set, chunk of the scalars), keeps the 2^(c-1)
// counters of its set in shared memory, and both the histogram and the scatter pass use shared-memory atomics only.
// code = (|d| - 1) | (d < 0) << 15, 0xffff for a zero digit (|d| - 1 = 32767 with the sign set would be d = -2^15,
// outside the digit range [-(2^(c-1) - 1), 2^(c-1)] for every c <= 16).
struct KRecode {
    enum { kBlock = 256 };
    const uint32_t *scalars;
    uint16_t *dig;    // W rows of `stride` codes
    uint32_t stride;  // >= n, multiple of 8 (rows stay 16-byte aligned)
    MsmGeom g;
    VMSM_HD void operator()(uint32_t tid) const {
        sc256 s = ld_scalar(scalars, tid);
        uint32_t carry = 0;
        for (uint32_t w = 0; w < g.W; w++) {
            int32_t d = sc_digit(s, w, g.c, carry);
            uint32_t code = 0xffffu;
            if (d != 0) code = d < 0 ? ((uint32_t)(-d) - 1u) | 0x8000u : (uint32_t)d - 1u;
            dig[(size_t)w * stride + tid] = (uint16_t)code;
            // the last scalar's thread also fills the row's padding (the sort reads whole 16-byte vectors of codes)
            if (tid + 1 == g.n)
                for (uint32_t p = g.n; p < stride; p++) dig[(size_t)w * stride + p] = 0xffffu;
        }
    }
};

// Long buckets.  A thread of KAccumulate sums at most `cap` entries of its bucket; what is left of a longer bucket is
// cut into at most 64 segments ("overflow tasks", each a multiple of 32 entries, >= 256) that KOverflow sums with one
// warp per task, and KCombine folds the task partials back into the bucket.  This bounds the serial chain of any
// thread whatever the scalar distribution: the top window of a 253-bit scalar has only a few live bits (few, huge
// buckets), and real witnesses are full of tiny values (bits, padding), cf. circuit_sat_cb.py:46-56.
struct OverflowTask {
    uint32_t bucket, first, count;  // idx[first .. first + count)
};
struct LongBucket {
    uint32_t bucket, task_base, ntask;
};
struct OverflowCtl {
    uint32_t ntasks, nlong;
};

// One thread per bucket.  `order` (optional) lists bucket ids by decreasing population so the lanes of a warp
// run the same trip count.
// Measured dead ends (profiles/r01/accumulate_variants.md): forcing 5 or 6 blocks/SM with __launch_bounds__ (96 / 80
// registers, +1.5 % / +8 % time), and prefetch.global.L2/.L1 of the next base one addition ahead (+4 %): at 100
// registers and 16 warps/SM the gather latency is already covered by the other warps' additions.
// Where the base of a CSR entry lives.  Plain MSM: bases[i] / extra[i - n_main].  Precomputed bases (PRE): level
// w = set + k * S of the table, i.e. the point 2^(c*w) * P_i, at bases[w * stride + i] (`bases` already points at the
// first term of the MSM inside level 0); the few extra terms (h, k of a Pedersen commitment) have their own table.
struct BaseRef {
    const ge_niels *bases;
    const ge_niels *extra;    // may be null
    uint32_t n_main;
    uint32_t stride, extra_stride;  // points per table level (PRE only)
    uint32_t lg, S, log2NB;         // CSR entry layout / bucket id -> set (PRE only)
    template <bool PRE>
    VMSM_HD const ge_niels *ptr(uint32_t e, uint32_t bucket) const {
        if (!PRE) {
            uint32_t i = e & 0x7fffffffu;
            return i < n_main ? bases + i : extra + (i - n_main);
        }
        uint32_t i = e & ((1u << lg) - 1u);
        uint32_t w = (bucket >> log2NB) + ((e & 0x7fffffffu) >> lg) * S;
        return i < n_main ? bases + (size_t)w * stride + i : extra + (size_t)w * extra_stride + (i - n_main);
    }
};

template <bool PRE>
struct KAccumulateT {
    enum { kBlock = 128 };
    BaseRef br;
    const uint32_t *offsets;  // sets x NB exclusive prefix (absolute position in idx)
    const uint32_t *counts;   // sets x NB
    const uint32_t *idx;
    const uint32_t *order;    // may be null
    ge_ext *buckets;          // sets x NB
    uint32_t nbuckets;
    uint32_t cap;             // entries summed by the bucket's own thread
    OverflowCtl *ctl;
    OverflowTask *tasks;
    LongBucket *longs;
    VMSM_HD void operator()(uint32_t tid) const {
        uint32_t b = order ? order[tid] : tid;
        uint32_t pos = offsets[b], cnt = counts[b];
        if (cnt > cap) {
            uint32_t over = cnt - cap;
            uint32_t seg = (over + 63) / 64;
            if (seg < 256) seg = 256;
            seg = (seg + 31) & ~31u;
            uint32_t ntask = (over + seg - 1) / seg;
            uint32_t base = VMSM_ATOMIC_ADD(&ctl->ntasks, ntask);
            uint32_t lpos = VMSM_ATOMIC_ADD(&ctl->nlong, 1u);
            for (uint32_t k = 0; k < ntask; k++) {
                OverflowTask t = {b, pos + cap + k * seg, over - k * seg < seg ? over - k * seg : seg};
                tasks[base + k] = t;
            }
            LongBucket lb = {b, base, ntask};
            longs[lpos] = lb;
            cnt = cap;
        }
        ge_ext acc = ge_identity();
        if (cnt) {
            uint32_t e = idx[pos];
            uint32_t en = cnt > 1 ? idx[pos + 1] : 0u;
            acc = ge_from_niels(ld_niels(br.template ptr<PRE>(e, b)), (e >> 31) != 0);  // first base: 1M instead of 7M
            e = en;
            for (uint32_t k = 1; k < cnt; k++) {
                en = (k + 1 < cnt) ? idx[pos + k + 1] : 0u;  // index prefetch: one load ahead of the gather
                ge_niels q = ld_niels(br.template ptr<PRE>(e, b));
                acc = ge_madd(acc, q, (e >> 31) != 0);
                e = en;
            }
        }
        st_ext(buckets + b, acc);
    }
};
typedef KAccumulateT<false> KAccumulate;
typedef KAccumulateT<true> KAccumulatePre;

// One warp per overflow task (grid-stride over the device-side task count): lanes stride over the segment, then a
// shuffle tree adds the 32 lane sums.
template <bool PRE>
struct KOverflowT {
    enum { kBlock = 128 };
    BaseRef br;
    const uint32_t *idx;
    const OverflowCtl *ctl;
    const OverflowTask *tasks;
    ge_ext *partials;
    uint32_t nwarps;
    VMSM_HD void operator()(uint32_t tid) const {
        const uint32_t ntasks = ctl->ntasks;
#if defined(__CUDA_ARCH__)
        const uint32_t lane = tid & 31;
        for (uint32_t t = tid >> 5; t < ntasks; t += nwarps) {
            OverflowTask tk = tasks[t];
            ge_ext acc = ge_identity();
            for (uint32_t k = lane; k < tk.count; k += 32) {
                uint32_t e = idx[tk.first + k];
                acc = ge_madd(acc, ld_niels(br.template ptr<PRE>(e, tk.bucket)), (e >> 31) != 0);
            }
#pragma unroll 1
            for (int d = 16; d >= 1; d >>= 1) {
                ge_ext o;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    o.X.v[i] = __shfl_down_sync(0xffffffffu, acc.X.v[i], d);
                    o.Y.v[i] = __shfl_down_sync(0xffffffffu, acc.Y.v[i], d);
                    o.Z.v[i] = __shfl_down_sync(0xffffffffu, acc.Z.v[i], d);
                    o.T.v[i] = __shfl_down_sync(0xffffffffu, acc.T.v[i], d);
                }
                acc = ge_add(acc, o);
            }
            if (lane == 0) st_ext(partials + t, acc);
        }
#else
        if (tid & 31) return;
        for (uint32_t t = tid >> 5; t < ntasks; t += nwarps) {
            OverflowTask tk = tasks[t];
            ge_ext acc = ge_identity();
            for (uint32_t k = 0; k < tk.count; k++) {
                uint32_t e = idx[tk.first + k];
                acc = ge_madd(acc, ld_niels(br.template ptr<PRE>(e, tk.bucket)), (e >> 31) != 0);
            }
            st_ext(partials + t, acc);
        }
#endif
    }
};
typedef KOverflowT<false> KOverflow;
typedef KOverflowT<true> KOverflowPre;

// One thread per long bucket (grid-stride): bucket += its task partials.
struct KCombine {
    enum { kBlock = 128 };
    const OverflowCtl *ctl;
    const LongBucket *longs;
    const ge_ext *partials;
    ge_ext *buckets;
    uint32_t nthreads;
    VMSM_HD void operator()(uint32_t tid) const {
        const uint32_t nlong = ctl->nlong;
        for (uint32_t l = tid; l < nlong; l += nthreads) {
            LongBucket lb = longs[l];
            ge_ext acc = ld_ext(buckets + lb.bucket);
            for (uint32_t k = 0; k < lb.ntask; k++) acc = ge_add(acc, ld_ext(partials + lb.task_base + k));
            st_ext(buckets + lb.bucket, acc);
        }
    }
};

// ---------------------------------------------------------------------------------------------- segmented accumulate
// Balanced bucket accumulation.  One thread per BUCKET (KAccumulateT above) runs as fast as the machine allows only when
// there are many more buckets than resident threads and their populations are alike: a 2^17-term slice of a sharded
// MSM, or windows sharing bucket sets over precomputed bases, have neither (measured: 2^20 terms over ONE bucket set,
// 32768 buckets of 512 entries, 2.36 ms against 1.13 ms for 16 sets).  Here the sorted CSR array is cut into equal
// SEGMENTS of L entries instead, one thread each, whatever buckets they fall into:
//   * a bucket that lies inside one segment is summed and stored by that thread;
//   * a bucket that straddles segments leaves one partial sum per segment -- a thread has at most two such partials,
//     for the run that reaches it from the previous segment ("head") and the run that leaves it ("tail");
//   * KSegFixup (one thread per bucket) adds the partials of a straddling bucket, writes the identity into empty
//     buckets and hands buckets spanning more than `long_span` segments to KSegLongFix (one warp per bucket).
// Every thread does the same number of mixed additions, so the launch is sized in whole waves of resident threads
// (seg_plan) and skewed scalars (boolean witnesses: one bucket holding half the MSM) need no separate path.
// seg_bucket[t] = the bucket holding entry t * L (written by the scan that produces the CSR offsets).
template <bool PRE>
struct KAccumulateSegT {
    // 48 KB of (unused) dynamic shared memory per block: at most FOUR blocks per SM.  At 96 registers a fifth block
    // would fit, which measured no faster alone (profiles/r01/accumulate_variants.md) and leaves no registers for the
    // thin counting-sort blocks of the next MSM that run underneath this kernel (2^20 terms: 1.55 instead of 1.45 ms).
    enum { kBlock = 128, kDynSmem = 48 * 1024 };
    BaseRef br;
    const uint32_t *offsets;     // per bucket, positions in the CONTIGUOUS idx array
    const uint32_t *counts;
    const uint32_t *idx;
    const uint32_t *seg_bucket;  // per segment
    const uint32_t *total;       // device word: number of CSR entries E
    ge_ext *buckets;
    ge_ext *partials;            // [2t] head, [2t + 1] tail of segment t
    uint32_t L;
    VMSM_HD void flush(uint32_t tid, uint32_t b, const ge_ext &acc, bool cont, bool more) const {
        if (cont) st_ext(partials + 2 * (size_t)tid, acc);
        else if (more) st_ext(partials + 2 * (size_t)tid + 1, acc);
        else st_ext(buckets + b, acc);
    }
    VMSM_HD void operator()(uint32_t tid) const {
        const uint32_t E = *total;
        const uint32_t pos0 = tid * L;
        if (pos0 >= E) return;
        const uint32_t pos1 = E - pos0 < L ? E : pos0 + L;
        uint32_t b = seg_bucket[tid];
        uint32_t off = offsets[b];
        uint32_t bend = off + counts[b];
        bool cont = off < pos0;  // the first run started in an earlier segment
        ge_ext acc = ge_identity();
        uint32_t e = idx[pos0];
        for (uint32_t pos = pos0; pos < pos1; pos++) {
            uint32_t en = pos + 1 < pos1 ? idx[pos + 1] : 0u;  // index prefetch: one load ahead of the gather
            if (pos == bend) {  // the run of bucket b is complete: next non-empty bucket (rare, divergent, cheap)
                flush(tid, b, acc, cont, false);
                cont = false;
                uint32_t c;
                do {
                    b++;
                    c = counts[b];
                } while (c == 0);
                bend += c;  // offsets are contiguous: bucket b starts where its predecessor ended
                acc = ge_identity();
            }
            // the mixed addition stays outside every divergent branch (a run's first base is added to the identity:
            // 7M instead of 1M once per run, in exchange for a warp that never executes two variants of the step)
            acc = ge_madd(acc, ld_niels(br.template ptr<PRE>(e, b)), (e >> 31) != 0);
            e = en;
        }
        flush(tid, b, acc, cont, bend > pos1);
    }
};
typedef KAccumulateSegT<false> KAccumulateSeg;
typedef KAccumulateSegT<true> KAccumulateSegPre;

struct KSegFixup {
    enum { kBlock = 128 };
    const uint32_t *offsets;
    const uint32_t *counts;
    const ge_ext *partials;
    ge_ext *buckets;
    uint32_t L, long_span;
    OverflowCtl *ctl;   // nlong
    LongBucket *longs;  // {bucket, first segment, segments after the first}
    VMSM_HD void operator()(uint32_t b) const {
        const uint32_t cnt = counts[b];
        if (!cnt) {
            st_ext(buckets + b, ge_identity());
            return;
        }
        const uint32_t off = offsets[b];
        const uint32_t t0 = off / L, t1 = (off + cnt - 1) / L;
        if (t0 == t1) return;  // summed and stored by its segment's thread
        if (t1 - t0 > long_span) {
            uint32_t lpos = VMSM_ATOMIC_ADD(&ctl->nlong, 1u);
            LongBucket lb = {b, t0, t1 - t0};
            longs[lpos] = lb;
            return;
        }
        ge_ext acc = ld_ext(partials + 2 * (size_t)t0 + 1);
        for (uint32_t t = t0 + 1; t <= t1; t++) acc = ge_add(acc, ld_ext(partials + 2 * (size_t)t));
        st_ext(buckets + b, acc);
    }
};

// One warp per long bucket (grid-stride over the device-side count): lanes stride over its partials, shuffle tree.
struct KSegLongFix {
    enum { kBlock = 128 };
    const OverflowCtl *ctl;
    const LongBucket *longs;
    const ge_ext *partials;
    ge_ext *buckets;
    uint32_t nwarps;
    VMSM_HD const ge_ext *part(const LongBucket &lb, uint32_t k) const {
        return k == 0 ? partials + 2 * (size_t)lb.task_base + 1 : partials + 2 * ((size_t)lb.task_base + k);
    }
    VMSM_HD void operator()(uint32_t tid) const {
        const uint32_t nlong = ctl->nlong;
#if defined(__CUDA_ARCH__)
        const uint32_t lane = tid & 31;
        for (uint32_t l = tid >> 5; l < nlong; l += nwarps) {
            LongBucket lb = longs[l];
            ge_ext acc = ge_identity();
            for (uint32_t k = lane; k <= lb.ntask; k += 32) acc = ge_add(acc, ld_ext(part(lb, k)));
#pragma unroll 1
            for (int d = 16; d >= 1; d >>= 1) {
                ge_ext o;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    o.X.v[i] = __shfl_down_sync(0xffffffffu, acc.X.v[i], d);
                    o.Y.v[i] = __shfl_down_sync(0xffffffffu, acc.Y.v[i], d);
                    o.Z.v[i] = __shfl_down_sync(0xffffffffu, acc.Z.v[i], d);
                    o.T.v[i] = __shfl_down_sync(0xffffffffu, acc.T.v[i], d);
                }
                acc = ge_add(acc, o);
            }
            if (lane == 0) st_ext(buckets + lb.bucket, acc);
        }
#else
        if (tid & 31) return;
        for (uint32_t l = tid >> 5; l < nlong; l += nwarps) {
            LongBucket lb = longs[l];
            ge_ext acc = ge_identity();
            for (uint32_t k = 0; k <= lb.ntask; k++) acc = ge_add(acc, ld_ext(part(lb, k)));
            st_ext(buckets + lb.bucket, acc);
        }
#endif
    }
};

// Radix-R merge of bucket-tree nodes.  A node covering buckets [lo, lo + s) carries
//   S = sum B_k            T = sum (k - lo) B_k
// Leaves are the buckets themselves (s = 1, T = 0, inT == null).  Merging children i = 0..m-1 of size s:
//   S' = sum S_i           T' = sum T_i + s * sum i S_i
struct KReduce {
    enum { kBlock = 128, kMinBlocks = 4 };
    const ge_ext *inS;
    const ge_ext *inT;  // null at the leaf level
    ge_ext *outS;
    ge_ext *outT;
    uint32_t cnt_in;   // nodes per window on input
    uint32_t cnt_out;  // nodes per window on output = ceil(cnt_in / R)
    uint32_t R;        // radix (power of two)
    uint32_t log2s;    // log2 of the bucket span of one input node
    VMSM_HD void operator()(uint32_t tid) const {
        uint32_t w = tid / cnt_out, j = tid - w * cnt_out;
        uint32_t first = j * R;
        uint32_t m = cnt_in - first < R ? cnt_in - first : R;
        const ge_ext *s = inS + (size_t)w * cnt_in + first;
        ge_ext acc = ge_identity(), run = ge_identity();
        for (uint32_t i = m - 1; i >= 1; i--) {
            acc = ge_add(acc, ld_ext(s + i));
            run = ge_add(run, acc);
        }
        acc = ge_add(acc, ld_ext(s));
        for (uint32_t k = 0; k < log2s; k++) run = ge_dbl(run);
        if (inT) {
            const ge_ext *t = inT + (size_t)w * cnt_in + first;
            for (uint32_t i = 0; i < m; i++) run = ge_add(run, ld_ext(t + i));
        }
        st_ext(outS + (size_t)w * cnt_out + j, acc);
        st_ext(outT + (size_t)w * cnt_out + j, run);
    }
};

// Single thread: window w total = T_w + S_w (bucket b weighs b+1), Horner over windows, normalise.
struct KFinal {
    enum { kBlock = 32 };
    const ge_ext *S;  // W root nodes
    const ge_ext *T;
    ge_ext *out_ext;
    ge_aff *out_aff;
    uint32_t W, c;
    // when set: the extended result goes to host-mapped memory and the HOST normalises it (one inversion is ~265
    // dependent field multiplications: 0.19 ms for a lone GPU thread, ~15 us for a CPU core)
    ge_ext *out_host_ext;
    VMSM_HD void operator()(uint32_t) const {
        ge_ext acc = ge_identity();
        for (int32_t w = (int32_t)W - 1; w >= 0; w--) {
            if (w != (int32_t)W - 1)
                for (uint32_t k = 0; k < c; k++) acc = ge_dbl(acc);
            acc = ge_add(acc, ge_add(ld_ext(S + w), ld_ext(T + w)));
        }
        st_ext(out_ext, acc);
        if (out_host_ext) {
            st_ext(out_host_ext, acc);
            return;
        }
        st_aff(out_aff, ge_ext_to_aff(acc));
    }
};

// Quad-cooperative twin of KReduce for the upper (latency-bound) levels of the bucket tree: 4 lanes per output node
// (ed25519_quad.cuh).  `nodes` output nodes, launched with 4 * nodes threads rounded up to a warp; the children count
// m = min(R, cnt_in) is uniform over the launch (all sizes are powers of two), so every shuffle is warp-uniform.
struct KReduceQ {
    enum { kBlock = 128 };
    const ge_ext *inS;
    const ge_ext *inT;  // null at the leaf level
    ge_ext *outS;
    ge_ext *outT;
    uint32_t cnt_in, cnt_out, R, log2s;
    uint32_t nodes;  // W * cnt_out
    VMSM_HD void operator()(uint32_t tid) const {
#if defined(__CUDA_ARCH__)
        const int q = tid & 3;
        uint32_t node = tid >> 2;
        const bool live = node < nodes;
        if (!live) node = nodes - 1;
        uint32_t w = node / cnt_out, j = node - w * cnt_out;
        const uint32_t m = cnt_in < R ? cnt_in : R;
        const ge_ext *s = inS + (size_t)w * cnt_in + (size_t)j * m;
        fe acc = quad_identity(q), run = quad_identity(q);
        for (uint32_t i = m - 1; i >= 1; i--) {
            acc = quad_add(q, acc, quad_load(s + i, q));
            run = quad_add(q, run, acc);
        }
        acc = quad_add(q, acc, quad_load(s, q));
        for (uint32_t k = 0; k < log2s; k++) run = quad_dbl(q, run);
        if (inT) {
            const ge_ext *t = inT + (size_t)w * cnt_in + (size_t)j * m;
            for (uint32_t i = 0; i < m; i++) run = quad_add(q, run, quad_load(t + i, q));
        }
        if (live) {
            quad_store(outS + (size_t)w * cnt_out + j, q, acc);
            quad_store(outT + (size_t)w * cnt_out + j, q, run);
        }
#else
        if ((tid & 3) || (tid >> 2) >= nodes) return;
        KReduce k = {inS, inT, outS, outT, cnt_in, cnt_out, R, log2s};
        k(tid >> 2);
#endif
    }
};

// Quad-cooperative twin of KFinal (one warp; every quad computes the same thing, quad 0 stores).
struct KFinalQ {
    enum { kBlock = 32 };
    const ge_ext *S;
    const ge_ext *T;
    ge_ext *out_ext;
    ge_aff *out_aff;
    uint32_t W, c;
    ge_ext *out_host_ext;  // see KFinal
    VMSM_HD void operator()(uint32_t tid) const {
#if defined(__CUDA_ARCH__)
        const int q = tid & 3;
        fe acc = quad_identity(q);
        for (int32_t w = (int32_t)W - 1; w >= 0; w--) {
            if (w != (int32_t)W - 1)
                for (uint32_t k = 0; k < c; k++) acc = quad_dbl(q, acc);
            acc = quad_add(q, acc, quad_add(q, quad_load(S + w, q), quad_load(T + w, q)));
        }
        if (out_host_ext) {
            if (tid < 4) {
                quad_store(out_ext, q, acc);
                quad_store(out_host_ext, q, acc);
            }
            return;
        }
        fe zi = fe_inv(quad_get(acc, 2));
        fe aff = fe_canon(fe_mul(acc, zi));  // lane 0: x, lane 1: y
        if (tid < 4) quad_store(out_ext, q, acc);
        if (tid < 2) st_fe(q == 0 ? &out_aff->x : &out_aff->y, aff);
#else
        if (tid) return;
        KFinal k = {S, T, out_ext, out_aff, W, c, out_host_ext};
        k(0);
#endif
    }
};

// ---------------------------------------------------------------------------------------------- multi-GPU partials
// Index-range split of one MSM over G GPUs: every GPU pushes the 128-byte partial result of its slice
// straight into a mailbox in the owner GPU's HBM (peer store over NVLink, mapped through CUDA IPC or peer access),
// then publishes a sequence number; the owner's gather kernel waits for the G sequence numbers, adds the partials and
// normalises.  No host hop, no NCCL.  One mailbox entry per (result slot, rank).
struct MailSlot {
    ge_ext pt;
    uint32_t seq;
    uint32_t pad[31];
};

struct KPushPartial {
    enum { kBlock = 32 };
    const ge_ext *src;  // this GPU's partial (its own result slot)
    MailSlot *dst;      // mailbox entry on the owner (may be a peer pointer)
    uint32_t seq;
    VMSM_HD void operator()(uint32_t tid) const {
        if (tid) return;
        st_ext(&dst->pt, ld_ext_plain(src));
#if defined(__CUDA_ARCH__)
        __threadfence_system();
        *reinterpret_cast<volatile uint32_t *>(&dst->seq) = seq;
        __threadfence_system();
#else
        dst->seq = seq;
#endif
    }
};

struct KGatherPartials {
    enum { kBlock = 32 };
    MailSlot *box;  // `world` consecutive entries of this result slot (owner-local memory)
    uint32_t world, seq;
    ge_ext *out_ext;
    ge_aff *out_aff;
    uint32_t *status;  // host-mapped: 0 ok, 1 timeout
    ge_ext *out_host_ext;  // when set: extended sum into host-mapped memory, the host normalises (see KFinal)
    VMSM_HD void operator()(uint32_t tid) const {
        if (tid) return;
        ge_ext acc = ge_identity();
        for (uint32_t r = 0; r < world; r++) {
#if defined(__CUDA_ARCH__)
            volatile uint32_t *flag = reinterpret_cast<volatile uint32_t *>(&box[r].seq);
            long long t0 = clock64();
            while (*flag != seq) {
                if (clock64() - t0 > (1ll << 33)) {  // ~4.5 s at 1.9 GHz: a peer died
                    *status = 1u;
                    return;
                }
                __nanosleep(200);
            }
            __threadfence_system();
#else
            if (box[r].seq != seq) {
                *status = 1u;
                return;
            }
#endif
            acc = ge_add(acc, ld_ext_plain(&box[r].pt));
        }
        st_ext(out_ext, acc);
        if (out_host_ext) st_ext(out_host_ext, acc);
        else st_aff(out_aff, ge_ext_to_aff(acc));
        *status = 0u;
    }
};

// ---------------------------------------------------------------------------------------------- point-set kernels
// Upload path: canonical affine -> niels, with validation (coordinates < p, on the curve).
struct KAffToNiels {
    enum { kBlock = 128 };
    const ge_aff *aff;
    ge_niels *niels;
    uint32_t *err;  // bit 0: non-canonical coordinate, bit 1: not on curve
    uint32_t check;
    VMSM_HD void operator()(uint32_t tid) const {
        ge_aff a = ld_aff(aff + tid);
        if (check) {
            uint32_t e = 0;
            if (!fe_is_canonical(a.x) || !fe_is_canonical(a.y)) e |= 1u;
            if (!ge_aff_on_curve(a)) e |= 2u;
            if (e) VMSM_ATOMIC_OR(err, e);
        }
        st_niels(niels + tid, ge_aff_to_niels(a));
    }
};

// asynchronous calls cannot return a validation error: it is published in the result slot's status word instead
struct KPublishErr {
    enum { kBlock = 32 };
    const uint32_t *err;
    uint32_t *status;  // mapped host memory
    uint32_t value;
    VMSM_HD void operator()(uint32_t tid) const {
        if (tid == 0 && *err) *status = value;
    }
};

// device-side copy of a point range (vmsm_points_concat)
struct KCopyPoints {
    enum { kBlock = 128 };
    const ge_aff *src_aff;
    const ge_niels *src_niels;
    ge_aff *dst_aff;
    ge_niels *dst_niels;
    VMSM_HD void operator()(uint32_t tid) const {
        st_aff(dst_aff + tid, ld_aff(src_aff + tid));
        st_niels(dst_niels + tid, ld_niels(src_niels + tid));
    }
};

// extended -> canonical affine + niels (one inversion per thread)
struct KNormalize {
    enum { kBlock = 128 };
    const ge_ext *in;
    ge_aff *aff;
    ge_niels *niels;
    VMSM_HD void operator()(uint32_t tid) const {
        ge_aff a = ge_ext_to_aff(ld_ext(in + tid));
        st_aff(aff + tid, a);
        st_niels(niels + tid, ge_aff_to_niels(a));
    }
};

// Base tables for FIXED generators.  The generators of a Pedersen vector commitment are created once
// (circuit_sat_r1cs.py:47-93) and then used for every commitment of every proof (circuit_sat_cb.py:103,
// compressed_pivot.py:110), so the doublings a windowed MSM spends on the weights 2^(c*w) can be paid once per
// generator instead of once per MSM: table[w][i] = 2^(c*w) * P_i in niels form, w < W.  One thread per point walks
// the levels (c doublings + one inversion each).  With the table all W windows of all points are just n*W independent
// (digit, base) pairs: they may share buckets, the bucket tree shrinks accordingly and the Horner chain disappears.
struct KPrecompute {
    enum { kBlock = 128 };
    const ge_aff *aff;   // n points
    ge_niels *table;     // W x stride
    uint32_t stride, c, W;
    VMSM_HD void operator()(uint32_t tid) const {
        ge_aff a = ld_aff(aff + tid);
        st_niels(table + tid, ge_aff_to_niels(a));
        for (uint32_t w = 1; w < W; w++) {
            ge_ext p = ge_aff_to_ext(a);
            for (uint32_t k = 0; k < c; k++) p = ge_dbl(p);
            a = ge_ext_to_aff(p);
            st_niels(table + (size_t)w * stride + tid, ge_aff_to_niels(a));
        }
    }
};

// Generator fold (compressed_pivot.py:64 / :178):  out[j] = c * P[j] + P[half + j]  with ONE shared scalar c,
// given in non-adjacent form as two 256-bit masks (nz: digit != 0, ng: digit < 0).  Uniform control flow across
// the grid; ~253 doublings + ~85 mixed additions per element.
struct KFold {
    enum { kBlock = 128 };
    const ge_niels *niels;  // 2*half inputs
    ge_ext *out;            // half outputs
    uint32_t half;
    int32_t top;  // index of the highest non-zero NAF digit, -1 when c == 0
    uint32_t nz[9];
    uint32_t ng[9];
    VMSM_HD void operator()(uint32_t tid) const {
        ge_niels p = ld_niels(niels + tid);
        ge_ext acc = ge_identity();
        for (int32_t i = top; i >= 0; i--) {
            acc = ge_dbl(acc);
            if ((nz[i >> 5] >> (i & 31)) & 1u) acc = ge_madd(acc, p, ((ng[i >> 5] >> (i & 31)) & 1u) != 0);
        }
        acc = ge_madd(acc, ld_niels(niels + half + tid), false);
        st_ext(out + tid, acc);
    }
};

// Quad-cooperative, fused twin of KFold + KNormalize for the latency-bound rounds (small halves): four lanes per
// element, 2 multiplication slots per doubling / mixed addition instead of 8 / 7 dependent multiplications, and the
// normalisation (one inversion) in the same kernel.  Reads niels[j], niels[half + j], writes aff[j] / niels[j] -- only
// its own element, so the fold is in place.  Launched with 4 * half threads rounded up to a warp.
struct KFoldQ {
    enum { kBlock = 128 };
    ge_aff *aff;
    ge_niels *niels;
    uint32_t half;
    int32_t top;
    uint32_t nz[9];
    uint32_t ng[9];
    VMSM_HD void operator()(uint32_t tid) const {
#if defined(__CUDA_ARCH__)
        const int q = tid & 3;
        uint32_t j = tid >> 2;
        const bool live = j < half;
        if (!live) j = half - 1;
        const ge_niels p = ld_niels(niels + j);
        const fe b_pos = quad_niels_operand(q, p, false), b_neg = quad_niels_operand(q, p, true);
        fe acc = quad_identity(q);
        for (int32_t i = top; i >= 0; i--) {
            acc = quad_dbl(q, acc);
            if ((nz[i >> 5] >> (i & 31)) & 1u) acc = quad_madd(q, acc, ((ng[i >> 5] >> (i & 31)) & 1u) ? b_neg : b_pos);
        }
        acc = quad_madd(q, acc, quad_niels_operand(q, ld_niels(niels + half + j), false));
        const fe zi = fe_inv(quad_get(acc, 2));
        const fe xy = fe_canon(fe_mul(acc, zi));  // lane 0: x, lane 1: y
        ge_aff a;
        a.x = quad_get(xy, 0);
        a.y = quad_get(xy, 1);
        if (live && q == 0) {
            st_aff(aff + j, a);
            st_niels(niels + j, ge_aff_to_niels(a));
        }
#else
        if ((tid & 3) || (tid >> 2) >= half) return;
        const uint32_t j = tid >> 2;
        ge_niels p = ld_niels(niels + j);
        ge_ext acc = ge_identity();
        for (int32_t i = top; i >= 0; i--) {
            acc = ge_dbl(acc);
            if ((nz[i >> 5] >> (i & 31)) & 1u) acc = ge_madd(acc, p, ((ng[i >> 5] >> (i & 31)) & 1u) != 0);
        }
        acc = ge_madd(acc, ld_niels(niels + half + j), false);
        ge_aff a = ge_ext_to_aff(acc);
        st_aff(aff + j, a);
        st_niels(niels + j, ge_aff_to_niels(a));
#endif
    }
};

// Fixed-base batch  out[i] = r_i * B  with r_i = synth_scalar_ed(seed, i) or explicit scalars; signed 4-bit windows
// over a host-built table tbl[64][8] (j * 16^w * B in niels form): 64 mixed additions, no doublings.
// This is the shape of create_generators (circuit_sat_r1cs.py:59-74: g_i = h ** r_i with h = group.generator).
struct KFixedBase {
    enum { kBlock = 128 };
    const ge_niels *tbl;      // 64 x 8
    const uint32_t *scalars;  // null -> synthetic from seed
    uint64_t seed;
    ge_ext *out;
    VMSM_HD void operator()(uint32_t tid) const {
        sc256 s = scalars ? ld_scalar(scalars, tid) : synth_scalar_ed(seed, tid);
        ge_ext acc = ge_identity();
        uint32_t carry = 0;
        for (uint32_t w = 0; w < 64; w++) {
            int32_t d = sc_digit(s, w, 4, carry);
            uint32_t a = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
            ge_niels q = a ? ld_niels(tbl + w * 8 + (a - 1)) : ge_niels_identity();
            acc = ge_madd(acc, q, d < 0);
        }
        st_ext(out + tid, acc);
    }
};

struct KSynthScalars {
    enum { kBlock = 256 };
    uint32_t *out;
    uint64_t seed;
    VMSM_HD void operator()(uint32_t tid) const {
        sc256 s = synth_scalar_ed(seed, tid);
        u32x4 a = {s.v[0], s.v[1], s.v[2], s.v[3]}, b = {s.v[4], s.v[5], s.v[6], s.v[7]};
        st128(out + 8ull * tid, a);
        st128(out + 8ull * tid + 4, b);
    }
};

// ---------------------------------------------------------------------------------------------- transcript text
// The Fiat-Shamir pre-image of every folding round contains the decimal text of ALL current generators
// (compressed_pivot.py:51-59 hashes str([A, B, g_hat, k, Q, L_tilde]); pivot.py:131-136).  Formatting 2*N 255-bit
// integers in Python dominated the prover's latency, so the device emits the text itself: one slot per point holding
// "[x, y, 1]" (the repr of a normalised point) and its length; a scan + compaction then produces the exact
// ", "-joined string that repr(list_of_points) has between its brackets.
#define VMSM_TEXT_SLOT 176  // >= 1 + 78 + 2 + 78 + 4

// decimal digits of a 256-bit value (8 limbs, little-endian), most significant first; returns the digit count
VMSM_HD uint32_t fe_to_decimal(const fe &v, uint8_t *out) {
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = v.v[i];
    uint32_t chunks[9];  // base 10^9, least significant first
    int nchunks = 0;
    for (int pass = 0; pass < 9; pass++) {
        uint64_t rem = 0;
        uint32_t any = 0;
#pragma unroll
        for (int i = 7; i >= 0; i--) {
            uint64_t cur = (rem << 32) | w[i];
            w[i] = (uint32_t)(cur / 1000000000u);
            rem = cur % 1000000000u;
            any |= w[i];
        }
        chunks[nchunks++] = (uint32_t)rem;
        if (!any) break;
    }
    uint32_t len = 0;
    for (int k = nchunks - 1; k >= 0; k--) {
        uint32_t c = chunks[k];
        uint8_t d[9];
        for (int j = 8; j >= 0; j--) {
            d[j] = (uint8_t)('0' + c % 10u);
            c /= 10u;
        }
        int start = 0;
        if (k == nchunks - 1)
            while (start < 8 && d[start] == '0') start++;  // no leading zeros on the top chunk (keeps a lone "0")
        for (int j = start; j < 9; j++) out[len++] = d[j];
    }
    return len;
}

struct KPointText {
    enum { kBlock = 128 };
    const ge_aff *aff;
    uint8_t *slots;  // n x VMSM_TEXT_SLOT
    uint32_t *lens;  // n, length of "[x, y, 1]" plus 2 for the ", " separator that follows all but the last
    uint32_t n;
    VMSM_HD void operator()(uint32_t tid) const {
        ge_aff a = ld_aff(aff + tid);
        uint8_t *o = slots + (size_t)tid * VMSM_TEXT_SLOT;
        uint32_t len = 0;
        o[len++] = '[';
        len += fe_to_decimal(a.x, o + len);
        o[len++] = ',';
        o[len++] = ' ';
        len += fe_to_decimal(a.y, o + len);
        o[len++] = ',';
        o[len++] = ' ';
        o[len++] = '1';
        o[len++] = ']';
        if (tid + 1 < n) {
            o[len++] = ',';
            o[len++] = ' ';
        }
        lens[tid] = len;
    }
};

struct KTextCompact {
    enum { kBlock = 128 };
    const uint8_t *slots;
    const uint32_t *lens;
    const uint64_t *offsets;  // exclusive prefix of lens
    uint8_t *out;
    uint32_t slot;  // bytes per slot
    VMSM_HD void operator()(uint32_t tid) const {
        const uint8_t *s = slots + (size_t)tid * slot;
        uint8_t *d = out + offsets[tid];
        uint32_t len = lens[tid];
        for (uint32_t i = 0; i < len; i++) d[i] = s[i];
    }
};

// device self-test of fe25519.cuh (vmsm_selftest_fe): results are canonicalised
struct KSelfTestFe {
    enum { kBlock = 128 };
    const fe *a;
    const fe *b;
    fe *out;
    int32_t op;
    VMSM_HD void operator()(uint32_t tid) const {
        fe x = ld_fe(a + tid), y = ld_fe(b + tid), r;
        switch (op) {
            case 0: r = fe_add(x, y); break;
            case 1: r = fe_sub(x, y); break;
            case 2: r = fe_mul(x, y); break;
            case 3: r = fe_inv(x); break;
            case 4: r = x; break;  // canon(a)
            case 5: r = fe_sqr(x); break;
            case 6: r = fe_mul(fe_add(x, y), fe_add(fe_add(x, y), y)); break;
            case 7: r = fe_mul(fe_sub(x, y), fe_dbl(fe_add(x, y))); break;
            case 8: r = fe_sqr(fe_add(x, y)); break;
            case 9: r = fe_sub(fe_add(fe_add(x, y), x), fe_add(fe_add(y, y), y)); break;
            default: r = fe_neg(x); break;
        }
        st_fe(out + tid, fe_canon(r));
    }
};

}  // namespace vmsm
