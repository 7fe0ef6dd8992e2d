// Scalar field of the Ed25519 group, Z / l with l = 2^252 + 27742317777372353535851937790883648493, for the vectors
// of the compressed pivot that are halved every round next to the generators: the witness z' = z_L + c z_R
// (verifiable_mpc/ac20/compressed_pivot.py:76) and the linear form L' = c L_L + L_R (:68-73), the cross terms
// L_R(z_L), L_L(z_R) that enter A_i and B_i (:41-42), and the decimal text of L' inside the next Fiat-Shamir
// pre-image (:51-54).  Keeping them in HBM removes the per-round host <-> device traffic and the O(n) Python loops.
//
// Elements are canonical residues on 8 x 32-bit limbs.  Multiplication is a word-serial Montgomery product
// (a * b / 2^256 mod l) in portable C: these kernels move 64 bytes per ~200 integer instructions and are far from
// any limit (n <= 2^20 per round, halving), so the same code runs on the device and in the host emulation.  Callers
// arrange the factor 2^256: a constant is passed in Montgomery form (c * 2^256), a sum of products is fixed once.
#pragma once
#include "kernels.cuh"

namespace vmsm {

struct scl {
    uint32_t v[8];
};

#define SCL_LIMBS(a, b, c, d, e, f, g, h) {{a, b, c, d, e, f, g, h}}
VMSM_HD scl scl_l() { scl r = SCL_LIMBS(0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0u, 0u, 0u, 0x10000000u); return r; }
VMSM_HD scl scl_half() { scl r = SCL_LIMBS(0x2e7ae9f6u, 0x2c09318du, 0x517bce6bu, 0x0a6f7cefu, 0u, 0u, 0u, 0x08000000u); return r; }  // l >> 1
VMSM_HD scl scl_r2() { scl r = SCL_LIMBS(0x449c0f01u, 0xa40611e3u, 0x68859347u, 0xd00e1ba7u, 0x17f5be65u, 0xceec73d2u, 0x7c309a3du, 0x0399411bu); return r; }  // 2^512 mod l
#define SCL_LINV 0x12547e1bu  // -l^-1 mod 2^32

VMSM_HD scl scl_zero() {
    scl r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}

// a > b
VMSM_HD bool scl_gt(const scl &a, const scl &b) {
    bool gt = false, decided = false;
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        if (!decided && a.v[i] != b.v[i]) {
            gt = a.v[i] > b.v[i];
            decided = true;
        }
    }
    return gt;
}

// a - b over the integers (caller guarantees a >= b)
VMSM_HD scl scl_sub_raw(const scl &a, const scl &b) {
    scl r;
    int64_t bw = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        bw += (int64_t)a.v[i] - (int64_t)b.v[i];
        r.v[i] = (uint32_t)bw;
        bw >>= 32;
    }
    return r;
}

// x (< 2l, given as 8 limbs + carry limb) -> x mod l
VMSM_HD scl scl_cond_sub(const scl &x, uint32_t top) {
    const scl l = scl_l();
    scl d;
    int64_t bw = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        bw += (int64_t)x.v[i] - (int64_t)l.v[i];
        d.v[i] = (uint32_t)bw;
        bw >>= 32;
    }
    bw += (int64_t)top;
    const bool use = bw >= 0;  // x >= l
    scl r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = use ? d.v[i] : x.v[i];
    return r;
}

VMSM_HD scl scl_add(const scl &a, const scl &b) {
    scl s;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a.v[i] + b.v[i];
        s.v[i] = (uint32_t)c;
        c >>= 32;
    }
    return scl_cond_sub(s, (uint32_t)c);
}

// a * b / 2^256 mod l (inputs < l)
VMSM_HD scl scl_mont_mul(const scl &a, const scl &b) {
    const scl l = scl_l();
    uint32_t t[10];
#pragma unroll
    for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)t[j] + (uint64_t)a.v[j] * b.v[i];
            t[j] = (uint32_t)c;
            c >>= 32;
        }
        c += t[8];
        t[8] = (uint32_t)c;
        t[9] = (uint32_t)(c >> 32);
        const uint32_t m = t[0] * SCL_LINV;
        c = ((uint64_t)t[0] + (uint64_t)m * l.v[0]) >> 32;
#pragma unroll
        for (int j = 1; j < 8; j++) {
            c += (uint64_t)t[j] + (uint64_t)m * l.v[j];
            t[j - 1] = (uint32_t)c;
            c >>= 32;
        }
        c += t[8];
        t[7] = (uint32_t)c;
        t[8] = t[9] + (uint32_t)(c >> 32);
    }
    scl r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
    return scl_cond_sub(r, t[8]);
}

// x -> x * 2^256 mod l
VMSM_HD scl scl_to_mont(const scl &x) { return scl_mont_mul(x, scl_r2()); }

VMSM_HD scl ld_scl(const uint32_t *base, uint64_t i) {
    u32x4 a = ld128(base + 8ull * i), b = ld128(base + 8ull * i + 4);
    scl s = {{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
    return s;
}
VMSM_HD void st_scl(uint32_t *base, uint64_t i, const scl &s) {
    u32x4 a = {s.v[0], s.v[1], s.v[2], s.v[3]}, b = {s.v[4], s.v[5], s.v[6], s.v[7]};
    st128(base + 8ull * i, a);
    st128(base + 8ull * i + 4, b);
}

// ---------------------------------------------------------------------------------------------- kernels
// dst[j] (op) src[j] with a constant c (c_mont = c * 2^256 mod l):
//   mode 0  dst[j] = dst[j] + c * src[j]    witness halving z' = z_L + c z_R (src = upper half), z = r + c0 x
//   mode 1  dst[j] = c * dst[j] + src[j]    linear-form halving L' = c L_L + L_R
//   mode 2  dst[j] = c * dst[j]             L_tilde = c1 * L
struct KScalarAxpy {
    enum { kBlock = 128 };
    uint32_t *dst;
    const uint32_t *src;
    scl c_mont;
    int32_t mode;
    VMSM_HD void operator()(uint32_t tid) const {
        scl d = ld_scl(dst, tid), r;
        if (mode == 2) {
            r = scl_mont_mul(c_mont, d);
        } else {
            scl t = ld_scl(src, tid);
            r = mode == 0 ? scl_add(d, scl_mont_mul(c_mont, t)) : scl_add(scl_mont_mul(c_mont, d), t);
        }
        st_scl(dst, tid, r);
    }
};

// partial[t] = sum over i = t, t + T, ... of a[i] * b[i] / 2^256 (T threads)
struct KScalarDotPartial {
    enum { kBlock = 128 };
    const uint32_t *a;
    const uint32_t *b;
    uint32_t n, T;
    uint32_t *partial;
    VMSM_HD void operator()(uint32_t tid) const {
        scl acc = scl_zero();
        for (uint32_t i = tid; i < n; i += T) acc = scl_add(acc, scl_mont_mul(ld_scl(a, i), ld_scl(b, i)));
        st_scl(partial, tid, acc);
    }
};

// out[t] = sum over i = t, t + T, ... of in[i]; with `fix`, times 2^256 (undoes the Montgomery factor of the products)
struct KScalarSum {
    enum { kBlock = 64 };
    const uint32_t *in;
    uint32_t n_in, T;
    uint32_t *out;
    int32_t fix;
    VMSM_HD void operator()(uint32_t tid) const {
        scl acc = scl_zero();
        for (uint32_t i = tid; i < n_in; i += T) acc = scl_add(acc, ld_scl(in, i));
        if (fix) acc = scl_to_mont(acc);
        st_scl(out, tid, acc);
    }
};

// decimal text of each residue, as MPyC prints field elements: the signed representative (v - l when v > l >> 1)
// for signed field classes, the residue itself otherwise; ", " follows all but the last.
#define VMSM_SCALAR_TEXT_SLOT 96  // >= 1 + 77 + 2
struct KScalarText {
    enum { kBlock = 128 };
    const uint32_t *v;
    uint8_t *slots;
    uint32_t *lens;
    uint32_t n;
    int32_t is_signed;
    VMSM_HD void operator()(uint32_t tid) const {
        scl s = ld_scl(v, tid);
        uint8_t *o = slots + (size_t)tid * VMSM_SCALAR_TEXT_SLOT;
        uint32_t len = 0;
        if (is_signed && scl_gt(s, scl_half())) {
            o[len++] = '-';
            s = scl_sub_raw(scl_l(), s);
        }
        fe f;
#pragma unroll
        for (int i = 0; i < 8; i++) f.v[i] = s.v[i];
        len += fe_to_decimal(f, o + len);
        if (tid + 1 < n) {
            o[len++] = ',';
            o[len++] = ' ';
        }
        lens[tid] = len;
    }
};

}  // namespace vmsm
