// Integer-pipe microbenchmarks for the roofline denominator ("must be confirmed on the box").
// Prints one JSON object per test: ops/s over the whole chip, ops/clk/SM from clock64, and the SM clock implied
// by cycles / wall time.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../verifiable_mpc_b200/csrc/ed25519.cuh"

using namespace vmsm;

#define ITERS 4096

// 1) IMAD.WIDE.U32 Rd = Ra * Rb + RZ (multiply only): 8 chains per thread, the multiplicand is the low word of the
//    chain's previous product so nothing is loop invariant.
__global__ void k_imad_wide_rz(uint64_t *out, uint32_t a, uint32_t b, long long *cycles) {
    uint64_t acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = a + threadIdx.x + i;
    uint32_t y = b | 1u;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = (uint64_t)((uint32_t)acc[i] + (uint32_t)(acc[i] >> 32)) * y;
    }
    long long t1 = clock64();
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// 2) IMAD.WIDE.U32 Rd = Ra * Rb + Rc64 (multiply-accumulate with a 64-bit register addend, no carry out)
__global__ void k_imad_wide_acc(uint64_t *out, uint32_t a, uint32_t b, long long *cycles) {
    uint64_t acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = a + threadIdx.x + i;
    uint32_t y = b | 1u;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = (uint64_t)((uint32_t)acc[(i + 1) & 7]) * y + acc[i];
    }
    long long t1 = clock64();
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// 3) 32-bit IMAD Rd = Ra * Rb + Rc
__global__ void k_imad_lo(uint32_t *out, uint32_t a, uint32_t b, long long *cycles) {
    uint32_t acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = a + threadIdx.x + i;
    uint32_t y = b | 1u;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = acc[(i + 1) & 7] * y + acc[i];
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// 3b) 64-bit add pairs (IADD3 + IADD3.X) on the ALU pipe
__global__ void k_iadd64(uint64_t *out, uint32_t a, uint32_t b, long long *cycles) {
    uint64_t acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = a + threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] += acc[(i + 1) & 7] ^ b;
    }
    long long t1 = clock64();
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// 4) carry-chained wide MADs exactly as fe_mul issues them: 4 products per chain, 2 chains per iteration
__global__ void k_imad_wide_cc(uint32_t *out, uint32_t a, uint32_t b, long long *cycles) {
    uint32_t r[2][9];
#pragma unroll
    for (int k = 0; k < 2; k++)
#pragma unroll
        for (int i = 0; i < 9; i++) r[k][i] = threadIdx.x + i + k;
    uint32_t x0 = a + threadIdx.x, x1 = x0 * 3, x2 = x0 * 5, x3 = x0 * 7, y = b;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < 2; k++)
            fe_mad4(r[k][0], r[k][1], r[k][2], r[k][3], r[k][4], r[k][5], r[k][6], r[k][7], r[k][8], x0, x1, x2, x3, y);
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 2; k++)
#pragma unroll
        for (int i = 0; i < 9; i++) s ^= r[k][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// 4b) FP64 FMA chains: would the double-precision pipe be a second multiplier for 26-bit limb products?
__global__ void k_dfma(double *out, uint32_t a, uint32_t b, long long *cycles) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = 1.0 + 1e-9 * (double)(a + threadIdx.x + i);
    double y = 1.0 + 1e-12 * (double)b;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = fma(acc[i], y, acc[(i + 1) & 7]);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// 4c) both at once: 4 IMAD.WIDE chains and 4 DFMA chains per thread (do the two pipes run concurrently?)
__global__ void k_dfma_imad(double *out, uint32_t a, uint32_t b, long long *cycles) {
    double acc[4];
    uint64_t iac[4];
#pragma unroll
    for (int i = 0; i < 4; i++) acc[i] = 1.0 + 1e-9 * (double)(a + threadIdx.x + i), iac[i] = a + threadIdx.x + i;
    double y = 1.0 + 1e-12 * (double)b;
    uint32_t yi = b | 1u;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            acc[i] = fma(acc[i], y, acc[(i + 1) & 3]);
            iac[i] = (uint64_t)((uint32_t)iac[i] + (uint32_t)(iac[i] >> 32)) * yi;
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) s += acc[i] + (double)(iac[i] & 0xffff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// 5) dependent field multiplications, 2 independent chains per thread
__global__ void k_fe_mul(fe *out, const fe *in, int iters, long long *cycles) {
    fe a = in[0], b = in[1];
    a.v[0] += threadIdx.x;
    b.v[1] += blockIdx.x;
    fe c = a, d = b;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        c = fe_mul(c, a);
        d = fe_mul(d, b);
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = fe_add(c, d);
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// 5b) dependent squarings
__global__ void k_fe_sqr(fe *out, const fe *in, int iters, long long *cycles) {
    fe c = in[0], d = in[1];
    c.v[0] += threadIdx.x;
    d.v[1] += blockIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        c = fe_sqr(c);
        d = fe_sqr(d);
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = fe_add(c, d);
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// 6) register-resident mixed additions (no memory traffic)
__global__ void k_madd(ge_ext *out, const fe *in, int iters, long long *cycles) {
    ge_niels q;
    q.ypx = in[0];
    q.ymx = in[1];
    q.t2d = in[2];
    q.ypx.v[0] += threadIdx.x;
    ge_ext acc = ge_identity();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) acc = ge_madd(acc, q, (it & 1) != 0);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// 7) doublings
__global__ void k_dbl(ge_ext *out, const fe *in, int iters, long long *cycles) {
    ge_ext acc = ge_identity();
    acc.X = in[0];
    acc.Y = in[1];
    acc.X.v[0] += threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) acc = ge_dbl(acc);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static double avg_cycles(long long *d, int n) {
    long long *h = (long long *)malloc(n * sizeof(long long));
    cudaMemcpy(h, d, n * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < n; i++) s += (double)h[i];
    free(h);
    return s / n;
}

template <class L>
static void run(const char *name, int blocks, int threads, double ops_per_thread, double lp_per_op, long long *cyc, L launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();  // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    double cycles = avg_cycles(cyc, blocks);
    double total = ops_per_thread * threads * (double)blocks;
    double ops_s = total / (best * 1e-3);
    int sms = 148;
    double per_clk_sm = total / sms / cycles;  // valid when every SM holds the same number of blocks
    printf("{\"test\": \"%s\", \"blocks\": %d, \"threads\": %d, \"ms\": %.4f, \"Tops_s\": %.4f, \"T_limb_products_s\": %.4f, "
           "\"ops_per_clk_per_sm\": %.2f, \"implied_sm_mhz\": %.0f, \"err\": \"%s\"}\n",
           name, blocks, threads, best, ops_s / 1e12, ops_s * lp_per_op / 1e12, per_clk_sm, cycles / (best * 1e-3) / 1e6,
           cudaGetErrorString(err));
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"clock_khz\": %d}\n", prop.name, prop.multiProcessorCount,
           prop.major, prop.minor, prop.clockRate);
    void *out;
    cudaMalloc(&out, 148 * 8 * 1024 * sizeof(ge_ext));
    long long *cyc;
    cudaMalloc(&cyc, 148 * 8 * sizeof(long long));
    fe hin[3];
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < 8; i++) hin[k].v[i] = 0x9e3779b9u * (k * 8 + i + 1);
    fe *din;
    cudaMalloc(&din, sizeof hin);
    cudaMemcpy(din, hin, sizeof hin, cudaMemcpyHostToDevice);

    for (int bps = 1; bps <= 2; bps++) {
        int blocks = 148 * bps, threads = 512;
        run("imad_wide_rz_multiply_only", blocks, threads, 8.0 * ITERS, 1.0, cyc, [&] { k_imad_wide_rz<<<blocks, threads>>>((uint64_t *)out, 12345, 6789, cyc); });
        run("imad_wide_acc64_fused", blocks, threads, 8.0 * ITERS, 1.0, cyc, [&] { k_imad_wide_acc<<<blocks, threads>>>((uint64_t *)out, 12345, 6789, cyc); });
        run("imad_lo_u32", blocks, threads, 8.0 * ITERS, 0.5, cyc, [&] { k_imad_lo<<<blocks, threads>>>((uint32_t *)out, 12345, 6789, cyc); });
        run("iadd64_pairs", blocks, threads, 8.0 * ITERS, 0.0, cyc, [&] { k_iadd64<<<blocks, threads>>>((uint64_t *)out, 12345, 6789, cyc); });
        run("imad_wide_carry_chain", blocks, threads, 8.0 * ITERS, 1.0, cyc, [&] { k_imad_wide_cc<<<blocks, threads>>>((uint32_t *)out, 12345, 6789, cyc); });
        run("dfma_fp64", blocks, threads, 8.0 * ITERS, 0.0, cyc, [&] { k_dfma<<<blocks, threads>>>((double *)out, 12345, 6789, cyc); });
        run("dfma_plus_imad_wide_4and4", blocks, threads, 8.0 * ITERS, 0.5, cyc, [&] { k_dfma_imad<<<blocks, threads>>>((double *)out, 12345, 6789, cyc); });
    }
    const int it = 512;
    for (int threads = 128; threads <= 512; threads *= 2) {
        int blocks = 148 * (512 / threads) ;
        run("fe_mul", blocks, threads, 2.0 * it, 72.0, cyc, [&] { k_fe_mul<<<blocks, threads>>>((fe *)out, din, it, cyc); });
        run("fe_sqr", blocks, threads, 2.0 * it, 44.0, cyc, [&] { k_fe_sqr<<<blocks, threads>>>((fe *)out, din, it, cyc); });
        if (threads <= 256) run("ge_madd", blocks, threads, 1.0 * it, 504.0, cyc, [&] { k_madd<<<blocks, threads>>>((ge_ext *)out, din, it, cyc); });
        run("ge_dbl", blocks, threads, 1.0 * it, 464.0, cyc, [&] { k_dbl<<<blocks, threads>>>((ge_ext *)out, din, it, cyc); });
    }
    {
        int blocks = 148 * 5, threads = 128;
        run("ge_madd_5x128", blocks, threads, 1.0 * it, 504.0, cyc, [&] { k_madd<<<blocks, threads>>>((ge_ext *)out, din, it, cyc); });
    }
    return 0;
}
