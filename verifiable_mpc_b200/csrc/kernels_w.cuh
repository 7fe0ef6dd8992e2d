// MSM kernels for the BN256 groups (G1 over Fp, G2 over Fp2), written once over the field policy F (fbn256.cuh) on top
// of the Jacobian arithmetic of bn256.cuh.  Same pipeline and same curve-agnostic front half (digit recoding,
// histogram, scan, scatter, bucket ordering) as the Ed25519 path in kernels.cuh; only the point arithmetic differs.
// Functors are per-thread bodies, so tests/hostemu runs them on the CPU as well.
#pragma once
#include "bn256.cuh"
#include "kernels.cuh"

namespace vmsm {

// 16-byte vector copies of whole point structs (all sizes are multiples of 32 B, arrays are 256 B aligned)
template <class T>
VMSM_HD T ld_obj(const T *p) {
    T r;
    uint32_t *w = reinterpret_cast<uint32_t *>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(T) / 16); k++) {
        u32x4 t = ld128(reinterpret_cast<const uint8_t *>(p) + 16 * k);
        w[4 * k] = t.x, w[4 * k + 1] = t.y, w[4 * k + 2] = t.z, w[4 * k + 3] = t.w;
    }
    return r;
}
template <class T>
VMSM_HD void st_obj(T *p, const T &v) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(&v);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(T) / 16); k++) {
        u32x4 t = {w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]};
        st128(reinterpret_cast<uint8_t *>(p) + 16 * k, t);
    }
}

// BN256 scalars are reduced below the 256-bit group order n (top bit set): all 256 bits are live
VMSM_HD sc256 synth_scalar_bn(uint64_t seed, uint64_t i) {
    sc256 s;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint64_t wv = splitmix64_mix(seed + 0x9E3779B97F4A7C15ull * (4 * i + j + 1));
        s.v[2 * j] = (uint32_t)wv;
        s.v[2 * j + 1] = (uint32_t)(wv >> 32);
    }
    s.v[8] = 0;
    const uint32_t N[8] = {0x57ac7261u, 0x1a2ef45bu, 0xf82b3924u, 0x2e8d8e12u, 0x6184dc21u, 0xaa6fecb8u, 0x4aa387f9u, 0x8fb501e3u};
    uint32_t d[8];
    int64_t bw = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        bw += (int64_t)s.v[k] - (int64_t)N[k];
        d[k] = (uint32_t)bw;
        bw >>= 32;
    }
    if (bw == 0) {
#pragma unroll
        for (int k = 0; k < 8; k++) s.v[k] = d[k];
    }
    return s;
}

// Where the base of a CSR entry e lives: plain (stride == 0) bases[i] / extra[i - n_main] with i = e & 0x7fffffff; over
// tables of 2^(c*w) * P_i (stride != 0, see KPrecomputeW) level w = set + k * S of the table, where set = bucket >>
// log2NB and k (which of the windows sharing that bucket set, MsmGeom) sits above bit lg of the entry.
template <class F>
struct BaseRefW {
    const waff<F> *bases;
    const waff<F> *extra;
    uint32_t n_main;
    uint32_t stride, extra_stride, log2NB;
    uint32_t lg, S;
    VMSM_HD const waff<F> *ptr(uint32_t e, uint32_t bucket) const {
        if (!stride) {
            const uint32_t i = e & 0x7fffffffu;
            return i < n_main ? bases + i : extra + (i - n_main);
        }
        const uint32_t i = e & ((1u << lg) - 1u);
        const size_t w = (bucket >> log2NB) + (size_t)((e & 0x7fffffffu) >> lg) * S;
        return i < n_main ? bases + w * stride + i : extra + w * extra_stride + (i - n_main);
    }
};

template <class F>
struct KAccumulateW {
    enum { kBlock = 128 };
    BaseRefW<F> br;
    const uint32_t *offsets, *counts, *idx, *order;
    wjac<F> *buckets;
    uint32_t nbuckets, cap;
    OverflowCtl *ctl;
    OverflowTask *tasks;
    LongBucket *longs;
    uint32_t seg_min;  // shortest overflow segment (a multiple of 32)
    VMSM_HD void operator()(uint32_t tid) const {
        uint32_t b = order ? order[tid] : tid;
        uint32_t pos = offsets[b], cnt = counts[b];
        if (cnt > cap) {
            uint32_t over = cnt - cap;
            uint32_t seg = (over + 63) / 64;
            if (seg < seg_min) seg = seg_min;
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
        wjac<F> acc = wj_identity<F>();
        for (uint32_t k = 0; k < cnt; k++) {
            uint32_t e = idx[pos + k];
            acc = wj_madd(acc, ld_obj(br.ptr(e, b)), (e >> 31) != 0);
        }
        st_obj(buckets + b, acc);
    }
};

template <class F>
struct KOverflowW {
    enum { kBlock = 128 };
    BaseRefW<F> br;
    const uint32_t *idx;
    const OverflowCtl *ctl;
    const OverflowTask *tasks;
    wjac<F> *partials;
    uint32_t nwarps;
    VMSM_HD void operator()(uint32_t tid) const {
        const uint32_t ntasks = ctl->ntasks;
#if defined(__CUDA_ARCH__)
        const uint32_t lane = tid & 31;
        for (uint32_t t = tid >> 5; t < ntasks; t += nwarps) {
            OverflowTask tk = tasks[t];
            wjac<F> acc = wj_identity<F>();
            for (uint32_t k = lane; k < tk.count; k += 32) {
                uint32_t e = idx[tk.first + k];
                acc = wj_madd(acc, ld_obj(br.ptr(e, tk.bucket)), (e >> 31) != 0);
            }
#pragma unroll 1
            for (int d = 16; d >= 1; d >>= 1) {
                wjac<F> o;
                uint32_t *ow = reinterpret_cast<uint32_t *>(&o);
                const uint32_t *aw = reinterpret_cast<const uint32_t *>(&acc);
#pragma unroll
                for (int i = 0; i < (int)(sizeof(wjac<F>) / 4); i++) ow[i] = __shfl_down_sync(0xffffffffu, aw[i], d);
                acc = wj_add(acc, o);
            }
            if (lane == 0) st_obj(partials + t, acc);
        }
#else
        if (tid & 31) return;
        for (uint32_t t = tid >> 5; t < ntasks; t += nwarps) {
            OverflowTask tk = tasks[t];
            wjac<F> acc = wj_identity<F>();
            for (uint32_t k = 0; k < tk.count; k++) {
                uint32_t e = idx[tk.first + k];
                acc = wj_madd(acc, ld_obj(br.ptr(e, tk.bucket)), (e >> 31) != 0);
            }
            st_obj(partials + t, acc);
        }
#endif
    }
};

template <class F>
struct KCombineW {
    enum { kBlock = 128 };
    const OverflowCtl *ctl;
    const LongBucket *longs;
    const wjac<F> *partials;
    wjac<F> *buckets;
    uint32_t nthreads;
    VMSM_HD void operator()(uint32_t tid) const {
        const uint32_t nlong = ctl->nlong;
        for (uint32_t l = tid; l < nlong; l += nthreads) {
            LongBucket lb = longs[l];
            wjac<F> acc = ld_obj(buckets + lb.bucket);
            for (uint32_t k = 0; k < lb.ntask; k++) acc = wj_add(acc, ld_obj(partials + lb.task_base + k));
            st_obj(buckets + lb.bucket, acc);
        }
    }
};

// Balanced (segmented) accumulation over key tables whose windows share bucket sets: the Weierstrass twin of
// KAccumulateSegT / KSegFixup / KSegLongFix in kernels.cuh (equal segments of the sorted CSR array, one thread each; a
// bucket that straddles segments leaves partial sums that the fix-up kernels add).  A 2^14-term MSM with c = 13 has 20
// windows x 4096 buckets of FOUR entries each: one thread per bucket is a launch of short chains plus a bucket tree as
// dear as the accumulation itself; over S shared sets the tree shrinks by W / S and the segments keep every resident
// thread equally busy whatever the bucket populations are.
template <class F>
struct KAccumulateSegW {
    enum { kBlock = 128 };
    BaseRefW<F> br;
    const uint32_t *offsets, *counts, *idx, *seg_bucket, *total;
    wjac<F> *buckets, *partials;  // partials: [2t] head, [2t + 1] tail of segment t
    uint32_t L;
    VMSM_HD void flush(uint32_t tid, uint32_t b, const wjac<F> &acc, bool cont, bool more) const {
        if (cont) st_obj(partials + 2 * (size_t)tid, acc);
        else if (more) st_obj(partials + 2 * (size_t)tid + 1, acc);
        else st_obj(buckets + b, acc);
    }
    VMSM_HD void operator()(uint32_t tid) const {
        const uint32_t E = *total;
        const uint32_t pos0 = tid * L;
        if (pos0 >= E) return;
        const uint32_t pos1 = E - pos0 < L ? E : pos0 + L;
        uint32_t b = seg_bucket[tid];
        const uint32_t off = offsets[b];
        uint32_t bend = off + counts[b];
        bool cont = off < pos0;
        wjac<F> acc = wj_identity<F>();
        for (uint32_t pos = pos0; pos < pos1; pos++) {
            if (pos == bend) {
                flush(tid, b, acc, cont, false);
                cont = false;
                uint32_t c;
                do {
                    b++;
                    c = counts[b];
                } while (c == 0);
                bend += c;
                acc = wj_identity<F>();
            }
            const uint32_t e = idx[pos];
            acc = wj_madd(acc, ld_obj(br.ptr(e, b)), (e >> 31) != 0);
        }
        flush(tid, b, acc, cont, bend > pos1);
    }
};

template <class F>
struct KSegFixupW {
    enum { kBlock = 128 };
    const uint32_t *offsets, *counts;
    const wjac<F> *partials;
    wjac<F> *buckets;
    uint32_t L, long_span;
    OverflowCtl *ctl;
    LongBucket *longs;  // {bucket, first segment, segments after the first}
    VMSM_HD void operator()(uint32_t b) const {
        const uint32_t cnt = counts[b];
        if (!cnt) {
            st_obj(buckets + b, wj_identity<F>());
            return;
        }
        const uint32_t off = offsets[b];
        const uint32_t t0 = off / L, t1 = (off + cnt - 1) / L;
        if (t0 == t1) return;
        if (t1 - t0 > long_span) {
            uint32_t lpos = VMSM_ATOMIC_ADD(&ctl->nlong, 1u);
            LongBucket lb = {b, t0, t1 - t0};
            longs[lpos] = lb;
            return;
        }
        wjac<F> acc = ld_obj(partials + 2 * (size_t)t0 + 1);
        for (uint32_t t = t0 + 1; t <= t1; t++) acc = wj_add(acc, ld_obj(partials + 2 * (size_t)t));
        st_obj(buckets + b, acc);
    }
};

template <class F>
struct KSegLongFixW {
    enum { kBlock = 128 };
    const OverflowCtl *ctl;
    const LongBucket *longs;
    const wjac<F> *partials;
    wjac<F> *buckets;
    uint32_t nwarps;
    VMSM_HD const wjac<F> *part(const LongBucket &lb, uint32_t k) const {
        return k == 0 ? partials + 2 * (size_t)lb.task_base + 1 : partials + 2 * ((size_t)lb.task_base + k);
    }
    VMSM_HD void operator()(uint32_t tid) const {
        const uint32_t nlong = ctl->nlong;
#if defined(__CUDA_ARCH__)
        const uint32_t lane = tid & 31;
        for (uint32_t l = tid >> 5; l < nlong; l += nwarps) {
            LongBucket lb = longs[l];
            wjac<F> acc = wj_identity<F>();
            for (uint32_t k = lane; k <= lb.ntask; k += 32) acc = wj_add(acc, ld_obj(part(lb, k)));
#pragma unroll 1
            for (int d = 16; d >= 1; d >>= 1) {
                wjac<F> o;
                uint32_t *ow = reinterpret_cast<uint32_t *>(&o);
                const uint32_t *aw = reinterpret_cast<const uint32_t *>(&acc);
#pragma unroll
                for (int i = 0; i < (int)(sizeof(wjac<F>) / 4); i++) ow[i] = __shfl_down_sync(0xffffffffu, aw[i], d);
                acc = wj_add(acc, o);
            }
            if (lane == 0) st_obj(buckets + lb.bucket, acc);
        }
#else
        if (tid & 31) return;
        for (uint32_t l = tid >> 5; l < nlong; l += nwarps) {
            LongBucket lb = longs[l];
            wjac<F> acc = wj_identity<F>();
            for (uint32_t k = 0; k <= lb.ntask; k++) acc = wj_add(acc, ld_obj(part(lb, k)));
            st_obj(buckets + lb.bucket, acc);
        }
#endif
    }
};

// (S, T) bucket tree, see KReduce in kernels.cuh
template <class F>
struct KReduceW {
    enum { kBlock = 128 };
    const wjac<F> *inS, *inT;
    wjac<F> *outS, *outT;
    uint32_t cnt_in, cnt_out, R, log2s;
    VMSM_HD void operator()(uint32_t tid) const {
        uint32_t w = tid / cnt_out, j = tid - w * cnt_out;
        uint32_t first = j * R;
        uint32_t m = cnt_in - first < R ? cnt_in - first : R;
        const wjac<F> *s = inS + (size_t)w * cnt_in + first;
        wjac<F> acc = wj_identity<F>(), run = wj_identity<F>();
        for (uint32_t i = m - 1; i >= 1; i--) {
            acc = wj_add(acc, ld_obj(s + i));
            run = wj_add(run, acc);
        }
        acc = wj_add(acc, ld_obj(s));
        for (uint32_t k = 0; k < log2s; k++) run = wj_dbl(run);
        if (inT) {
            const wjac<F> *t = inT + (size_t)w * cnt_in + first;
            for (uint32_t i = 0; i < m; i++) run = wj_add(run, ld_obj(t + i));
        }
        st_obj(outS + (size_t)w * cnt_out + j, acc);
        st_obj(outT + (size_t)w * cnt_out + j, run);
    }
};

// plain (non-Montgomery) canonical affine: the wire form; identity = all zero
template <class F>
VMSM_HD waff<F> wa_to_wire(const waff<F> &m) {
    waff<F> r = {F::from_mont(m.x), F::from_mont(m.y)};
    return r;
}

template <class F>
struct KFinalW {
    enum { kBlock = 32 };
    const wjac<F> *S, *T;
    wjac<F> *out_jac;
    waff<F> *out_wire;
    uint32_t W, c;
    wjac<F> *out_host_jac;  // when set: Jacobian result into host-mapped memory, the host normalises (see KFinal)
    VMSM_HD void operator()(uint32_t tid) const {
        if (tid) return;
        wjac<F> acc = wj_identity<F>();
        for (int32_t w = (int32_t)W - 1; w >= 0; w--) {
            if (w != (int32_t)W - 1)
                for (uint32_t k = 0; k < c; k++) acc = wj_dbl(acc);
            acc = wj_add(acc, wj_add(ld_obj(S + w), ld_obj(T + w)));
        }
        st_obj(out_jac, acc);
        if (out_host_jac) {
            st_obj(out_host_jac, acc);
            return;
        }
        st_obj(out_wire, wa_to_wire(wj_to_aff(acc)));
    }
};

// ---------------------------------------------------------------------------------------------- lane-cooperative tail
// Four adjacent lanes share one Jacobian operation of the Horner chain: every lane holds the whole point, the
// independent field multiplications of a formula stage are spread over the lanes (one each) and the products are
// broadcast with width-4 shuffles.  A doubling is 3 multiplication latencies instead of 7, a full addition 5 instead
// of 16.  The operands are identical in the four lanes, so the exceptional-case branches are uniform.
#if defined(__CUDA_ARCH__)
template <class T>
VMSM_D T wq_get(const T &v, int src) {
    T r;
    const uint32_t *in = reinterpret_cast<const uint32_t *>(&v);
    uint32_t *out = reinterpret_cast<uint32_t *>(&r);
    // only the caller's own quad takes part: quads of one warp may be in different branches (different nodes)
    const uint32_t mask = 0xfu << (threadIdx.x & 28u);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 4); i++) out[i] = __shfl_sync(mask, in[i], src, 4);
    return r;
}
template <class T>
VMSM_D T wq_pick(int q, const T &a0, const T &a1, const T &a2, const T &a3) {
    T r;
    const uint32_t *p0 = reinterpret_cast<const uint32_t *>(&a0), *p1 = reinterpret_cast<const uint32_t *>(&a1);
    const uint32_t *p2 = reinterpret_cast<const uint32_t *>(&a2), *p3 = reinterpret_cast<const uint32_t *>(&a3);
    uint32_t *out = reinterpret_cast<uint32_t *>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 4); i++) {
        uint32_t lo = (q & 1) ? p1[i] : p0[i], hi = (q & 1) ? p3[i] : p2[i];
        out[i] = (q & 2) ? hi : lo;
    }
    return r;
}

// 2P (dbl-2009-l, a = 0): stage 1 {X^2, Y^2, Y*Z}, stage 2 {B^2, (X+B)^2, E^2}, stage 3 E*(D - X3) on every lane
template <class F>
static __device__ __noinline__ wjac<F> wq_dbl(int q, const wjac<F> &p) {
    typedef typename F::T T;
    T m1 = F::mul(wq_pick(q, p.X, p.Y, p.Y, p.X), wq_pick(q, p.X, p.Y, p.Z, p.X));
    T A = wq_get(m1, 0), B = wq_get(m1, 1), YZ = wq_get(m1, 2);
    T E = F::add(F::dbl(A), A);
    T XB = F::add(p.X, B);
    T m2 = F::mul(wq_pick(q, B, XB, E, B), wq_pick(q, B, XB, E, B));
    T C = wq_get(m2, 0), t = wq_get(m2, 1), Fq = wq_get(m2, 2);
    T D = F::dbl(F::sub(F::sub(t, A), C));
    wjac<F> r;
    r.X = F::sub(Fq, F::dbl(D));
    T C8 = F::dbl(F::dbl(F::dbl(C)));
    r.Y = F::sub(F::mul(E, F::sub(D, r.X)), C8);
    r.Z = F::dbl(YZ);
    return r;
}

// P + Q (add-2007-bl), five stages
template <class F>
static __device__ __noinline__ wjac<F> wq_add(int q, const wjac<F> &p, const wjac<F> &o) {
    typedef typename F::T T;
    if (F::is_zero(p.Z)) return o;
    if (F::is_zero(o.Z)) return p;
    T m1 = F::mul(wq_pick(q, p.Z, o.Z, p.Y, o.Y), wq_pick(q, p.Z, o.Z, o.Z, p.Z));
    T Z1Z1 = wq_get(m1, 0), Z2Z2 = wq_get(m1, 1), Y1Z2 = wq_get(m1, 2), Y2Z1 = wq_get(m1, 3);
    T m2 = F::mul(wq_pick(q, p.X, o.X, Y1Z2, Y2Z1), wq_pick(q, Z2Z2, Z1Z1, Z2Z2, Z1Z1));
    T U1 = wq_get(m2, 0), U2 = wq_get(m2, 1), S1 = wq_get(m2, 2), S2 = wq_get(m2, 3);
    T H = F::sub(U2, U1);
    T rr = F::sub(S2, S1);
    if (F::is_zero(H)) {
        if (F::is_zero(rr)) return wq_dbl<F>(q, p);
        return wj_identity<F>();
    }
    T H2 = F::dbl(H), r2 = F::dbl(rr), ZS = F::add(p.Z, o.Z);
    T m3 = F::mul(wq_pick(q, H2, r2, ZS, H2), wq_pick(q, H2, r2, ZS, H2));
    T I = wq_get(m3, 0), R2 = wq_get(m3, 1), ZZ = wq_get(m3, 2);
    T Zt = F::sub(F::sub(ZZ, Z1Z1), Z2Z2);
    T m4 = F::mul(wq_pick(q, H, U1, Zt, H), wq_pick(q, I, I, H, I));
    T J = wq_get(m4, 0), V = wq_get(m4, 1);
    wjac<F> r;
    r.Z = wq_get(m4, 2);
    r.X = F::sub(F::sub(R2, J), F::dbl(V));
    T m5 = F::mul(wq_pick(q, r2, S1, r2, S1), wq_pick(q, F::sub(V, r.X), J, F::sub(V, r.X), J));
    r.Y = F::sub(wq_get(m5, 0), F::dbl(wq_get(m5, 1)));
    return r;
}
// P + (neg ? -Q : Q) with Q affine (madd-2007-bl), five stages instead of eleven multiplications in a row
template <class F>
static __device__ __noinline__ wjac<F> wq_madd(int q, const wjac<F> &p, const waff<F> &a, bool neg) {
    typedef typename F::T T;
    if (wa_is_identity(a)) return p;
    T qy = neg ? F::neg(a.y) : a.y;
    if (F::is_zero(p.Z)) {
        wjac<F> r = {a.x, qy, F::one()};
        return r;
    }
    T m1 = F::mul(wq_pick(q, p.Z, qy, p.Z, qy), wq_pick(q, p.Z, p.Z, p.Z, p.Z));
    T Z1Z1 = wq_get(m1, 0), Y2Z1 = wq_get(m1, 1);
    T m2 = F::mul(wq_pick(q, a.x, Y2Z1, a.x, Y2Z1), Z1Z1);
    T U2 = wq_get(m2, 0), S2 = wq_get(m2, 1);
    T H = F::sub(U2, p.X);
    T rr = F::sub(S2, p.Y);
    if (F::is_zero(H)) {
        if (F::is_zero(rr)) return wq_dbl<F>(q, p);
        return wj_identity<F>();
    }
    T ZH = F::add(p.Z, H), r2 = F::dbl(rr);
    T m3 = F::mul(wq_pick(q, H, ZH, r2, H), wq_pick(q, H, ZH, r2, H));
    T HH = wq_get(m3, 0), ZZ = wq_get(m3, 1), R2 = wq_get(m3, 2);
    T I = F::dbl(F::dbl(HH));
    T m4 = F::mul(wq_pick(q, H, p.X, H, p.X), I);
    T J = wq_get(m4, 0), V = wq_get(m4, 1);
    wjac<F> r;
    r.X = F::sub(F::sub(R2, J), F::dbl(V));
    T m5 = F::mul(wq_pick(q, r2, p.Y, r2, p.Y), wq_pick(q, F::sub(V, r.X), J, F::sub(V, r.X), J));
    r.Y = F::sub(wq_get(m5, 0), F::dbl(wq_get(m5, 1)));
    r.Z = F::sub(F::sub(ZZ, Z1Z1), HH);
    return r;
}
#endif

// Lane-cooperative twin of KAccumulateW for the latency-bound sizes: four lanes per bucket (launched with
// 4 * nbuckets threads); lane 0 registers the overflow tasks of a long bucket and stores the result.
template <class F>
struct KAccumulateWQ {
    enum { kBlock = 128 };
    BaseRefW<F> br;
    const uint32_t *offsets, *counts, *idx, *order;
    wjac<F> *buckets;
    uint32_t nbuckets, cap;
    OverflowCtl *ctl;
    OverflowTask *tasks;
    LongBucket *longs;
    uint32_t seg_min;
    VMSM_HD void operator()(uint32_t tid) const {
#if defined(__CUDA_ARCH__)
        const int q = tid & 3;
        const uint32_t slot = tid >> 2;
        uint32_t b = order ? order[slot] : slot;
        uint32_t pos = offsets[b], cnt = counts[b];
        if (cnt > cap) {
            if (q == 0) {
                uint32_t over = cnt - cap;
                uint32_t seg = (over + 63) / 64;
                if (seg < seg_min) seg = seg_min;
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
            }
            cnt = cap;
        }
        wjac<F> acc = wj_identity<F>();
        for (uint32_t k = 0; k < cnt; k++) {
            uint32_t e = idx[pos + k];
            acc = wq_madd<F>(q, acc, ld_obj(br.ptr(e, b)), (e >> 31) != 0);
        }
        if (q == 0) st_obj(buckets + b, acc);
#else
        if (tid & 3) return;
        KAccumulateW<F> k = {br, offsets, counts, idx, order, buckets, nbuckets, cap, ctl, tasks, longs, seg_min};
        k(tid >> 2);
#endif
    }
};

// Lane-cooperative twin of KSegFixupW: four lanes per bucket (launched with 4 * nbuckets threads).  A bucket of a
// 2^14-term MSM over two shared sets holds ~40 entries in ~4 segments: three dependent full additions per bucket, the
// longest chain between the accumulate kernel and the bucket tree.
template <class F>
struct KSegFixupWQ {
    enum { kBlock = 128 };
    const uint32_t *offsets, *counts;
    const wjac<F> *partials;
    wjac<F> *buckets;
    uint32_t L, long_span;
    OverflowCtl *ctl;
    LongBucket *longs;
    VMSM_HD void operator()(uint32_t tid) const {
#if defined(__CUDA_ARCH__)
        const int q = tid & 3;
        const uint32_t b = tid >> 2;
        const uint32_t cnt = counts[b];
        if (!cnt) {
            if (q == 0) st_obj(buckets + b, wj_identity<F>());
            return;
        }
        const uint32_t off = offsets[b];
        const uint32_t t0 = off / L, t1 = (off + cnt - 1) / L;
        if (t0 == t1) return;
        if (t1 - t0 > long_span) {
            if (q == 0) {
                uint32_t lpos = VMSM_ATOMIC_ADD(&ctl->nlong, 1u);
                LongBucket lb = {b, t0, t1 - t0};
                longs[lpos] = lb;
            }
            return;
        }
        wjac<F> acc = ld_obj(partials + 2 * (size_t)t0 + 1);
        for (uint32_t t = t0 + 1; t <= t1; t++) acc = wq_add<F>(q, acc, ld_obj(partials + 2 * (size_t)t));
        if (q == 0) st_obj(buckets + b, acc);
#else
        if (tid & 3) return;
        KSegFixupW<F> k = {offsets, counts, partials, buckets, L, long_span, ctl, longs};
        k(tid >> 2);
#endif
    }
};

// Lane-cooperative twin of KReduceW: four lanes per tree node (launched with 4 * nodes threads, a multiple of 4)
template <class F>
struct KReduceWQ {
    enum { kBlock = 128 };
    const wjac<F> *inS, *inT;
    wjac<F> *outS, *outT;
    uint32_t cnt_in, cnt_out, R, log2s;
    VMSM_HD void operator()(uint32_t tid) const {
#if defined(__CUDA_ARCH__)
        const int q = tid & 3;
        const uint32_t node = tid >> 2;
        uint32_t w = node / cnt_out, j = node - w * cnt_out;
        uint32_t first = j * R;
        uint32_t m = cnt_in - first < R ? cnt_in - first : R;
        const wjac<F> *s = inS + (size_t)w * cnt_in + first;
        wjac<F> acc = wj_identity<F>(), run = wj_identity<F>();
        for (uint32_t i = m - 1; i >= 1; i--) {
            acc = wq_add<F>(q, acc, ld_obj(s + i));
            run = wq_add<F>(q, run, acc);
        }
        acc = wq_add<F>(q, acc, ld_obj(s));
        for (uint32_t k = 0; k < log2s; k++) run = wq_dbl<F>(q, run);
        if (inT) {
            const wjac<F> *t = inT + (size_t)w * cnt_in + first;
            for (uint32_t i = 0; i < m; i++) run = wq_add<F>(q, run, ld_obj(t + i));
        }
        if (q == 0) {
            st_obj(outS + (size_t)w * cnt_out + j, acc);
            st_obj(outT + (size_t)w * cnt_out + j, run);
        }
#else
        if (tid & 3) return;
        KReduceW<F> k = {inS, inT, outS, outT, cnt_in, cnt_out, R, log2s};
        k(tid >> 2);
#endif
    }
};

// Lane-cooperative twin of KFinalW (one warp: every quad computes the same chain, lane 0 stores)
template <class F>
struct KFinalWQ {
    enum { kBlock = 32 };
    const wjac<F> *S, *T;
    wjac<F> *out_jac;
    waff<F> *out_wire;
    uint32_t W, c;
    wjac<F> *out_host_jac;
    VMSM_HD void operator()(uint32_t tid) const {
#if defined(__CUDA_ARCH__)
        const int q = tid & 3;
        wjac<F> acc = wj_identity<F>();
        for (int32_t w = (int32_t)W - 1; w >= 0; w--) {
            if (w != (int32_t)W - 1)
                for (uint32_t k = 0; k < c; k++) acc = wq_dbl<F>(q, acc);
            acc = wq_add<F>(q, acc, wq_add<F>(q, ld_obj(S + w), ld_obj(T + w)));
        }
        if (tid == 0) {
            st_obj(out_jac, acc);
            if (out_host_jac) st_obj(out_host_jac, acc);
            else st_obj(out_wire, wa_to_wire(wj_to_aff(acc)));
        }
#else
        KFinalW<F> k = {S, T, out_jac, out_wire, W, c, out_host_jac};
        k(tid);
#endif
    }
};

// Tables for a FIXED key (the evaluation key of pynocchio.py:101-167 is made once per circuit and used by every
// compute_proof): table[w][i] = 2^(c*w) * P_i in Montgomery affine form, w < W.  One thread per point, c doublings and
// one inversion per level.  An MSM over the tables needs no doubling at all: every window's bucket set carries weight 1.
template <class F>
struct KPrecomputeW {
    enum { kBlock = 128 };
    const waff<F> *base;  // n points
    waff<F> *table;       // W x stride
    uint32_t stride, c, W;
    VMSM_HD void operator()(uint32_t tid) const {
        waff<F> a = ld_obj(base + tid);
        st_obj(table + tid, a);
        for (uint32_t w = 1; w < W; w++) {
            wjac<F> p = wa_to_jac(a);
            for (uint32_t k = 0; k < c; k++) p = wj_dbl(p);
            a = wj_to_aff(p);
            st_obj(table + (size_t)w * stride + tid, a);
        }
    }
};

// upload: wire (plain canonical) -> validation -> Montgomery base
template <class F>
struct KUploadW {
    enum { kBlock = 128 };
    const waff<F> *wire;
    waff<F> *base;
    uint32_t *err;  // bit 0: coordinate >= p, bit 1: not on the curve
    uint32_t check;
    VMSM_HD void operator()(uint32_t tid) const {
        waff<F> w = ld_obj(wire + tid);
        uint32_t e = 0;
        if (check && (!F::plain_ok(w.x) || !F::plain_ok(w.y))) e |= 1u;
        waff<F> m = {F::to_mont(w.x), F::to_mont(w.y)};
        if (check && !wa_on_curve(m)) e |= 2u;
        if (e) VMSM_ATOMIC_OR(err, e);
        st_obj(base + tid, m);
    }
};

template <class F>
struct KNormalizeW {
    enum { kBlock = 128 };
    const wjac<F> *in;
    waff<F> *wire, *base;
    VMSM_HD void operator()(uint32_t tid) const {
        waff<F> a = wj_to_aff(ld_obj(in + tid));
        st_obj(base + tid, a);
        st_obj(wire + tid, wa_to_wire(a));
    }
};

// out[i] = r_i * G with signed 4-bit windows over a host-built table tbl[65][8] (j * 16^w * G); r_i explicit or synthetic
template <class F>
struct KFixedBaseW {
    enum { kBlock = 128 };
    const waff<F> *tbl;
    const uint32_t *scalars;
    uint64_t seed;
    wjac<F> *out;
    VMSM_HD void operator()(uint32_t tid) const {
        sc256 s = scalars ? ld_scalar(scalars, tid) : synth_scalar_bn(seed, tid);
        wjac<F> acc = wj_identity<F>();
        uint32_t carry = 0;
        for (uint32_t w = 0; w < 65; w++) {  // 65th window holds the carry out of bit 255
            int32_t d = sc_digit(s, w, 4, carry);
            if (d != 0) {
                uint32_t a = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
                acc = wj_madd(acc, ld_obj(tbl + w * 8 + (a - 1)), d < 0);
            }
        }
        st_obj(out + tid, acc);
    }
};

template <class F>
struct KCopyW {
    enum { kBlock = 128 };
    const waff<F> *src_wire, *src_base;
    waff<F> *dst_wire, *dst_base;
    VMSM_HD void operator()(uint32_t tid) const {
        st_obj(dst_wire + tid, ld_obj(src_wire + tid));
        st_obj(dst_base + tid, ld_obj(src_base + tid));
    }
};

struct KSynthScalarsBN {
    enum { kBlock = 256 };
    uint32_t *out;
    uint64_t seed;
    VMSM_HD void operator()(uint32_t tid) const {
        sc256 s = synth_scalar_bn(seed, tid);
        u32x4 a = {s.v[0], s.v[1], s.v[2], s.v[3]}, b = {s.v[4], s.v[5], s.v[6], s.v[7]};
        st128(out + 8ull * tid, a);
        st128(out + 8ull * tid + 4, b);
    }
};

}  // namespace vmsm
