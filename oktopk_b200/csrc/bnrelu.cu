// Fused (conv-bias +) BatchNorm + ReLU for channels_last fp32 activations, training mode, forward and backward.
//
// The CNN zoo of the reference (VGG/models/vgg.py:28-36 and the ResNets) is stacks of  Conv2d -> BatchNorm2d -> ReLU.
// Run through stock framework ops, a VGG-16 step at 16 images per GPU spends a quarter of its time in the glue around the
// convolutions (ncu launch list, profiles/launches_vgg_r2.md): per layer a bias-add kernel, a batch-norm kernel, a clamp
// kernel and a counter increment in the forward pass; a threshold kernel, a batch-norm backward kernel and a bias-gradient
// reduction in the backward pass -- all of them latency-bound passes over tensors of 32 K .. 1 M elements.  Here the
// whole block is two small kernels forward and two backward:
//
//   forward : (1) per-block partial sums of x and x^2 per channel;
//             (2) every block combines the partials in double (mean, 1/std), then streams  y = max(0, a_c x + b_c),
//                 block 0 also updates running_mean / running_var / num_batches_tracked and saves mean and 1/std;
//   backward: (3) per-block partials of  dbeta = sum(dy * [z>0])  and  dgamma = sum(dy * [z>0] * xhat);
//             (4) every block combines them and streams  dx = a_c (dy[z>0] - dbeta/M - xhat dgamma/M).
//
// A bias added before a batch-norm cancels exactly: BN(x + b) = BN(x) with the batch mean shifted by b.  The forward pass
// therefore never adds it (only running_mean sees it), and its gradient -- identically zero, since the loss does not depend
// on it -- is not "computed" by a reduction over dy that can only return rounding noise.
//
// Layout: x is [M, C] row-major (NHWC with M = N*H*W), C a multiple of 4; each thread owns 4 consecutive channels
// (128-bit accesses) and strides over rows; blocks own contiguous row ranges.
#include "common.cuh"
#include "oktopk.cuh"

namespace okt {

constexpr int kBnThreads = 256;
constexpr int kBnMaxBlocks = 512;

struct BnGeom {
    int M, C, cv;          // rows, channels, float4 columns (C/4)
    int tpr, rpi;          // threads per row, rows per block iteration
    int rows_per_block, nblk;
};

__host__ __device__ inline BnGeom bn_geom(int M, int C) {
    BnGeom g;
    g.M = M; g.C = C; g.cv = C >> 2;
    g.tpr = g.cv < kBnThreads ? g.cv : kBnThreads;
    g.rpi = kBnThreads / g.tpr;
    if (g.rpi < 1) g.rpi = 1;
    // ~8 K elements per block: enough blocks to spread a 1 M-element tensor over the GPU, a few blocks for the tiny layers
    long long want = ((long long)M * C + 8191) / 8192;
    if (want < 1) want = 1;
    if (want > kBnMaxBlocks) want = kBnMaxBlocks;
    int rpb = (int)((M + want - 1) / want);
    rpb = (rpb + g.rpi - 1) / g.rpi * g.rpi;
    if (rpb < g.rpi) rpb = g.rpi;
    g.rows_per_block = rpb;
    g.nblk = (M + rpb - 1) / rpb;
    return g;
}

__device__ __forceinline__ float4 f4_fma(const float4& x, const float4& a, const float4& b) {
    return make_float4(fmaf(x.x, a.x, b.x), fmaf(x.y, a.y, b.y), fmaf(x.z, a.z, b.z), fmaf(x.w, a.w, b.w));
}

// Combine the per-block partials [nblk][2C] into s_tot[2C] with ALL threads of the block: a tile of up to 256 float4
// columns at a time, the threads that share a column stride over the blocks, then one shared-memory reduction.  (A loop of
// "one thread per channel walks all nblk partials" is nblk dependent L2 round trips: 64 x ~0.6 us on the widest layers.)
__device__ __forceinline__ void bn_combine_partials(const float* __restrict__ partial, int nblk, int C, float* s_tot,
                                                    float4* s_scr) {
    const int cols = (2 * C) >> 2;                          // float4 columns of one partial row
    const float4* p4 = reinterpret_cast<const float4*>(partial);
    for (int c0 = 0; c0 < cols; c0 += kBnThreads) {
        const int w = min(kBnThreads, cols - c0);
        const int groups = kBnThreads / w;
        const int tx = threadIdx.x % w, ty = threadIdx.x / w;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ty < groups)
            for (int b = ty; b < nblk; b += groups) {
                const float4 v = __ldg(p4 + (size_t)b * cols + c0 + tx);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        s_scr[threadIdx.x] = acc;
        __syncthreads();
        if (ty == 0) {
            for (int j = 1; j < groups; ++j) {
                const float4 v = s_scr[j * w + tx];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            reinterpret_cast<float4*>(s_tot)[c0 + tx] = acc;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------- (1) partial statistics
// partial layout: [nblk][2][C]  (sum, sum of squares)
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const float* __restrict__ x, float* __restrict__ partial, BnGeom g) {
    extern __shared__ float4 s_red[];                 // [2][rpi][tpr]
    const int tx = threadIdx.x % g.tpr, ty = threadIdx.x / g.tpr;
    const int row0 = blockIdx.x * g.rows_per_block, row1 = min(g.M, row0 + g.rows_per_block);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (int c0 = 0; c0 < g.cv; c0 += g.tpr) {          // column tiles (one tile unless C > 1024)
        const int col = c0 + tx;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
        if (ty < g.rpi && col < g.cv) {
            auto acc = [&](const float4& v) {
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
            };
            // four independent 128-bit loads in flight per thread (the loop is latency-bound otherwise); same summation order
            int r = row0 + ty;
            const size_t st1 = (size_t)g.rpi * g.cv;
            for (; r + 3 * g.rpi < row1; r += 4 * g.rpi) {
                const float4* q0 = x4 + (size_t)r * g.cv + col;
                const float4 v0 = __ldg(q0), v1 = __ldg(q0 + st1), v2 = __ldg(q0 + 2 * st1), v3 = __ldg(q0 + 3 * st1);
                acc(v0); acc(v1); acc(v2); acc(v3);
            }
            for (; r < row1; r += g.rpi) acc(__ldg(x4 + (size_t)r * g.cv + col));
        }
        if (ty < g.rpi) { s_red[ty * g.tpr + tx] = s; s_red[(g.rpi + ty) * g.tpr + tx] = q; }
        __syncthreads();
        if (ty == 0 && col < g.cv) {
            for (int j = 1; j < g.rpi; ++j) {
                const float4 a = s_red[j * g.tpr + tx], b = s_red[(g.rpi + j) * g.tpr + tx];
                s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
                q.x += b.x; q.y += b.y; q.z += b.z; q.w += b.w;
            }
            float4* p4 = reinterpret_cast<float4*>(partial + (size_t)blockIdx.x * 2 * g.C);
            p4[col] = s;
            p4[g.cv + col] = q;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------- (2) normalise + affine + ReLU
__global__ void __launch_bounds__(kBnThreads) bn_apply_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                               const float* __restrict__ partial, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, const float* __restrict__ cbias,
                                                               float* __restrict__ save_mean, float* __restrict__ save_invstd,
                                                               float* __restrict__ rmean, float* __restrict__ rvar,
                                                               long long* __restrict__ nbt, float momentum, float eps, int relu,
                                                               BnGeom g) {
    extern __shared__ float4 s_ab[];                  // [2][cv]: a = gamma/std, b = beta - mean a; then [2][cv] totals; scratch
    __shared__ float4 s_scr[kBnThreads];
    const int tx = threadIdx.x % g.tpr, ty = threadIdx.x / g.tpr;
    float* s_tot = reinterpret_cast<float*>(s_ab) + 2 * g.C;
    // every block combines the partial sums of ALL blocks (nblk x 2C values, all threads), then mean / 1/std in double
    bn_combine_partials(partial, g.nblk, g.C, s_tot, s_scr);
    for (int c = threadIdx.x; c < g.C; c += kBnThreads) {
        const double s = (double)s_tot[c], q = (double)s_tot[g.C + c];
        const double mean = s / (double)g.M;
        double var = q / (double)g.M - mean * mean;
        if (var < 0.0) var = 0.0;
        const float invstd = (float)(1.0 / sqrt(var + (double)eps));
        const float a = __fmul_rn(__ldg(gamma + c), invstd);
        reinterpret_cast<float*>(s_ab)[c] = a;
        reinterpret_cast<float*>(s_ab)[g.C + c] = __fsub_rn(__ldg(beta + c), __fmul_rn((float)mean, a));   // no fma: backward recomputes it
        if (blockIdx.x == 0) {
            save_mean[c] = (float)mean;
            save_invstd[c] = invstd;
            if (rmean != nullptr) {                     // running statistics (unbiased variance), bias-shifted mean
                const float mb = (float)mean + (cbias != nullptr ? __ldg(cbias + c) : 0.f);
                const double unb = g.M > 1 ? var * (double)g.M / (double)(g.M - 1) : var;
                rmean[c] = (1.f - momentum) * rmean[c] + momentum * mb;
                rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && nbt != nullptr) *nbt += 1;
    __syncthreads();
    const int row0 = blockIdx.x * g.rows_per_block, row1 = min(g.M, row0 + g.rows_per_block);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    for (int c0 = 0; c0 < g.cv; c0 += g.tpr) {
        const int col = c0 + tx;
        if (ty >= g.rpi || col >= g.cv) continue;
        const float4 a = s_ab[col], b = s_ab[g.cv + col];
        auto out = [&](size_t off, const float4& xin) {
            float4 v = f4_fma(xin, a, b);
            if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            y4[off] = v;
        };
        int r = row0 + ty;
        const size_t st1 = (size_t)g.rpi * g.cv;
        for (; r + 3 * g.rpi < row1; r += 4 * g.rpi) {
            const size_t o = (size_t)r * g.cv + col;
            const float4 v0 = __ldg(x4 + o), v1 = __ldg(x4 + o + st1), v2 = __ldg(x4 + o + 2 * st1), v3 = __ldg(x4 + o + 3 * st1);
            out(o, v0); out(o + st1, v1); out(o + 2 * st1, v2); out(o + 3 * st1, v3);
        }
        for (; r < row1; r += g.rpi) { const size_t o = (size_t)r * g.cv + col; out(o, __ldg(x4 + o)); }
    }
}

// ---------------------------------------------------------------------------------------------- (3) backward partials
// partial layout: [nblk][2][C]  (dbeta, dgamma)
__global__ void __launch_bounds__(kBnThreads) bn_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                    const float* __restrict__ save_mean,
                                                                    const float* __restrict__ save_invstd,
                                                                    float* __restrict__ partial, int relu, BnGeom g) {
    extern __shared__ float4 s_red[];                 // [2][rpi][tpr]
    const int tx = threadIdx.x % g.tpr, ty = threadIdx.x / g.tpr;
    const int row0 = blockIdx.x * g.rows_per_block, row1 = min(g.M, row0 + g.rows_per_block);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* d4 = reinterpret_cast<const float4*>(dy);
    for (int c0 = 0; c0 < g.cv; c0 += g.tpr) {
        const int col = c0 + tx;
        float4 sb = make_float4(0.f, 0.f, 0.f, 0.f), sg = sb;
        if (ty < g.rpi && col < g.cv) {
            const float4 mean = __ldg(reinterpret_cast<const float4*>(save_mean) + col);
            const float4 istd = __ldg(reinterpret_cast<const float4*>(save_invstd) + col);
            const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + col);
            const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + col);
            const float4 a = make_float4(__fmul_rn(gm.x, istd.x), __fmul_rn(gm.y, istd.y), __fmul_rn(gm.z, istd.z), __fmul_rn(gm.w, istd.w));
            const float4 b = make_float4(__fsub_rn(bt.x, __fmul_rn(mean.x, a.x)), __fsub_rn(bt.y, __fmul_rn(mean.y, a.y)),
                                     __fsub_rn(bt.z, __fmul_rn(mean.z, a.z)), __fsub_rn(bt.w, __fmul_rn(mean.w, a.w)));
            auto acc = [&](const float4& v, float4 d) {
                const float4 xh = make_float4((v.x - mean.x) * istd.x, (v.y - mean.y) * istd.y, (v.z - mean.z) * istd.z,
                                              (v.w - mean.w) * istd.w);
                if (relu) {       // the forward output was max(0, fma(x, a, b)): same a, b, same fma => the same mask, bit for bit
                    if (!(fmaf(v.x, a.x, b.x) > 0.f)) d.x = 0.f;
                    if (!(fmaf(v.y, a.y, b.y) > 0.f)) d.y = 0.f;
                    if (!(fmaf(v.z, a.z, b.z) > 0.f)) d.z = 0.f;
                    if (!(fmaf(v.w, a.w, b.w) > 0.f)) d.w = 0.f;
                }
                sb.x += d.x; sb.y += d.y; sb.z += d.z; sb.w += d.w;
                sg.x = fmaf(d.x, xh.x, sg.x); sg.y = fmaf(d.y, xh.y, sg.y); sg.z = fmaf(d.z, xh.z, sg.z); sg.w = fmaf(d.w, xh.w, sg.w);
            };
            int r = row0 + ty;
            const size_t st1 = (size_t)g.rpi * g.cv;
            for (; r + 3 * g.rpi < row1; r += 4 * g.rpi) {         // 8 independent 128-bit loads in flight per thread
                const size_t o = (size_t)r * g.cv + col;
                const float4 v0 = __ldg(x4 + o), v1 = __ldg(x4 + o + st1), v2 = __ldg(x4 + o + 2 * st1), v3 = __ldg(x4 + o + 3 * st1);
                const float4 e0 = __ldg(d4 + o), e1 = __ldg(d4 + o + st1), e2 = __ldg(d4 + o + 2 * st1), e3 = __ldg(d4 + o + 3 * st1);
                acc(v0, e0); acc(v1, e1); acc(v2, e2); acc(v3, e3);
            }
            for (; r < row1; r += g.rpi) { const size_t o = (size_t)r * g.cv + col; acc(__ldg(x4 + o), __ldg(d4 + o)); }
        }
        if (ty < g.rpi) { s_red[ty * g.tpr + tx] = sb; s_red[(g.rpi + ty) * g.tpr + tx] = sg; }
        __syncthreads();
        if (ty == 0 && col < g.cv) {
            for (int j = 1; j < g.rpi; ++j) {
                const float4 a = s_red[j * g.tpr + tx], b = s_red[(g.rpi + j) * g.tpr + tx];
                sb.x += a.x; sb.y += a.y; sb.z += a.z; sb.w += a.w;
                sg.x += b.x; sg.y += b.y; sg.z += b.z; sg.w += b.w;
            }
            float4* p4 = reinterpret_cast<float4*>(partial + (size_t)blockIdx.x * 2 * g.C);
            p4[col] = sb;
            p4[g.cv + col] = sg;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------- (4) input gradient
__global__ void __launch_bounds__(kBnThreads) bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                   float* __restrict__ dx, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta,
                                                                   const float* __restrict__ save_mean,
                                                                   const float* __restrict__ save_invstd,
                                                                   const float* __restrict__ partial, float* __restrict__ dgamma,
                                                                   float* __restrict__ dbeta, int relu, BnGeom g) {
    extern __shared__ float4 s_c[];                   // [2][cv]: dbeta/M, dgamma/M; then [2][cv] totals
    __shared__ float4 s_scr[kBnThreads];
    const int tx = threadIdx.x % g.tpr, ty = threadIdx.x / g.tpr;
    float* s_tot = reinterpret_cast<float*>(s_c) + 2 * g.C;
    bn_combine_partials(partial, g.nblk, g.C, s_tot, s_scr);
    for (int c = threadIdx.x; c < g.C; c += kBnThreads) {
        const double sb = (double)s_tot[c], sg = (double)s_tot[g.C + c];
        reinterpret_cast<float*>(s_c)[c] = (float)(sb / (double)g.M);
        reinterpret_cast<float*>(s_c)[g.C + c] = (float)(sg / (double)g.M);
        if (blockIdx.x == 0) { dbeta[c] = (float)sb; dgamma[c] = (float)sg; }
    }
    __syncthreads();
    const int row0 = blockIdx.x * g.rows_per_block, row1 = min(g.M, row0 + g.rows_per_block);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* d4 = reinterpret_cast<const float4*>(dy);
    float4* o4 = reinterpret_cast<float4*>(dx);
    for (int c0 = 0; c0 < g.cv; c0 += g.tpr) {
        const int col = c0 + tx;
        if (ty >= g.rpi || col >= g.cv) continue;
        const float4 mean = __ldg(reinterpret_cast<const float4*>(save_mean) + col);
        const float4 istd = __ldg(reinterpret_cast<const float4*>(save_invstd) + col);
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + col);
        const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + col);
        const float4 mb = s_c[col], mg = s_c[g.cv + col];
        const float4 a = make_float4(__fmul_rn(gm.x, istd.x), __fmul_rn(gm.y, istd.y), __fmul_rn(gm.z, istd.z), __fmul_rn(gm.w, istd.w));
        const float4 b = make_float4(__fsub_rn(bt.x, __fmul_rn(mean.x, a.x)), __fsub_rn(bt.y, __fmul_rn(mean.y, a.y)),
                                     __fsub_rn(bt.z, __fmul_rn(mean.z, a.z)), __fsub_rn(bt.w, __fmul_rn(mean.w, a.w)));
        auto out = [&](size_t off, const float4& v, float4 d) {
            const float4 xh = make_float4((v.x - mean.x) * istd.x, (v.y - mean.y) * istd.y, (v.z - mean.z) * istd.z,
                                          (v.w - mean.w) * istd.w);
            if (relu) {
                if (!(fmaf(v.x, a.x, b.x) > 0.f)) d.x = 0.f;
                if (!(fmaf(v.y, a.y, b.y) > 0.f)) d.y = 0.f;
                if (!(fmaf(v.z, a.z, b.z) > 0.f)) d.z = 0.f;
                if (!(fmaf(v.w, a.w, b.w) > 0.f)) d.w = 0.f;
            }
            o4[off] = make_float4(a.x * (d.x - mb.x - xh.x * mg.x), a.y * (d.y - mb.y - xh.y * mg.y),
                                  a.z * (d.z - mb.z - xh.z * mg.z), a.w * (d.w - mb.w - xh.w * mg.w));
        };
        int r = row0 + ty;
        const size_t st1 = (size_t)g.rpi * g.cv;
        for (; r + 3 * g.rpi < row1; r += 4 * g.rpi) {
            const size_t o = (size_t)r * g.cv + col;
            const float4 v0 = __ldg(x4 + o), v1 = __ldg(x4 + o + st1), v2 = __ldg(x4 + o + 2 * st1), v3 = __ldg(x4 + o + 3 * st1);
            const float4 e0 = __ldg(d4 + o), e1 = __ldg(d4 + o + st1), e2 = __ldg(d4 + o + 2 * st1), e3 = __ldg(d4 + o + 3 * st1);
            out(o, v0, e0); out(o + st1, v1, e1); out(o + 2 * st1, v2, e2); out(o + 3 * st1, v3, e3);
        }
        for (; r < row1; r += g.rpi) { const size_t o = (size_t)r * g.cv + col; out(o, __ldg(x4 + o), __ldg(d4 + o)); }
    }
}

// ---------------------------------------------------------------------------------------------- launchers
int bn_num_blocks(int M, int C) { return bn_geom(M, C).nblk; }

cudaError_t launch_bn_forward(const float* x, float* y, float* partial, const float* gamma, const float* beta, const float* cbias,
                              float* save_mean, float* save_invstd, float* rmean, float* rvar, long long* nbt, float momentum,
                              float eps, int relu, int M, int C, cudaStream_t stream) {
    const BnGeom g = bn_geom(M, C);
    const size_t sm1 = sizeof(float4) * 2 * g.rpi * g.tpr, sm2 = sizeof(float) * 4 * C;
    bn_stats_kernel<<<g.nblk, kBnThreads, sm1, stream>>>(x, partial, g);
    bn_apply_kernel<<<g.nblk, kBnThreads, sm2, stream>>>(x, y, partial, gamma, beta, cbias, save_mean, save_invstd, rmean, rvar,
                                                         nbt, momentum, eps, relu, g);
    return cudaGetLastError();
}

cudaError_t launch_bn_backward(const float* x, const float* dy, float* dx, float* partial, const float* gamma, const float* beta,
                               const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta, int relu, int M,
                               int C, cudaStream_t stream) {
    const BnGeom g = bn_geom(M, C);
    const size_t sm1 = sizeof(float4) * 2 * g.rpi * g.tpr, sm2 = sizeof(float) * 4 * C;
    bn_bwd_reduce_kernel<<<g.nblk, kBnThreads, sm1, stream>>>(x, dy, gamma, beta, save_mean, save_invstd, partial, relu, g);
    bn_bwd_apply_kernel<<<g.nblk, kBnThreads, sm2, stream>>>(x, dy, dx, gamma, beta, save_mean, save_invstd, partial, dgamma,
                                                             dbeta, relu, g);
    return cudaGetLastError();
}

}  // namespace okt

// ==============================================================================================================
// 2x2 / stride-2 max pooling, channels_last fp32, forward + backward.
// The stock NHWC pooling kernels take 11-13 us per call on the VGG activations (16x32x32x64 ... 16x2x2x512) plus a
// zero-fill of the input gradient; the windows of a stride-2 2x2 pool do not overlap, so the backward pass can WRITE all
// four positions of every window (gradient at the arg-max, zeros elsewhere) without atomics or a memset.  One thread =
// one output pixel x 4 channels (128-bit accesses); the arg-max (first maximum in row-major window order, like
// at::max_pool2d) is kept as one byte per output element.
// ==============================================================================================================
namespace okt {

constexpr int kPoolThreads = 256;

__global__ void __launch_bounds__(kPoolThreads) maxpool2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                     unsigned char* __restrict__ arg, int N, int H, int W,
                                                                     int C) {
    const int cv = C >> 2, Ho = H >> 1, Wo = W >> 1;
    const long long total = (long long)N * Ho * Wo * cv;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    uchar4* a4 = reinterpret_cast<uchar4*>(arg);
    for (long long t = (long long)blockIdx.x * kPoolThreads + threadIdx.x; t < total; t += (long long)gridDim.x * kPoolThreads) {
        const int c = (int)(t % cv);
        long long p = t / cv;
        const int wo = (int)(p % Wo); p /= Wo;
        const int ho = (int)(p % Ho);
        const int n = (int)(p / Ho);
        const size_t base = (((size_t)n * H + 2 * ho) * W + 2 * wo) * cv + c;
        const float4 v0 = __ldg(x4 + base), v1 = __ldg(x4 + base + cv);
        const float4 v2 = __ldg(x4 + base + (size_t)W * cv), v3 = __ldg(x4 + base + (size_t)W * cv + cv);
        float4 m = v0;
        uchar4 a = make_uchar4(0, 0, 0, 0);
#define OKT_POOL_STEP(V, K)                                                   \
        if (V.x > m.x || V.x != V.x) { m.x = V.x; a.x = K; }              \
        if (V.y > m.y || V.y != V.y) { m.y = V.y; a.y = K; }              \
        if (V.z > m.z || V.z != V.z) { m.z = V.z; a.z = K; }              \
        if (V.w > m.w || V.w != V.w) { m.w = V.w; a.w = K; }
        OKT_POOL_STEP(v1, 1)
        OKT_POOL_STEP(v2, 2)
        OKT_POOL_STEP(v3, 3)
#undef OKT_POOL_STEP
        y4[t] = m;
        a4[t] = a;
    }
}

__global__ void __launch_bounds__(kPoolThreads) maxpool2_bwd_kernel(const float* __restrict__ dy,
                                                                     const unsigned char* __restrict__ arg,
                                                                     float* __restrict__ dx, int N, int H, int W, int C) {
    const int cv = C >> 2, Ho = H >> 1, Wo = W >> 1;
    const long long total = (long long)N * Ho * Wo * cv;
    const float4* d4 = reinterpret_cast<const float4*>(dy);
    const uchar4* a4 = reinterpret_cast<const uchar4*>(arg);
    float4* o4 = reinterpret_cast<float4*>(dx);
    for (long long t = (long long)blockIdx.x * kPoolThreads + threadIdx.x; t < total; t += (long long)gridDim.x * kPoolThreads) {
        const int c = (int)(t % cv);
        long long p = t / cv;
        const int wo = (int)(p % Wo); p /= Wo;
        const int ho = (int)(p % Ho);
        const int n = (int)(p / Ho);
        const size_t base = (((size_t)n * H + 2 * ho) * W + 2 * wo) * cv + c;
        const float4 d = __ldg(d4 + t);
        const uchar4 a = a4[t];
        auto pick = [&](int k) {
            return make_float4(a.x == k ? d.x : 0.f, a.y == k ? d.y : 0.f, a.z == k ? d.z : 0.f, a.w == k ? d.w : 0.f);
        };
        o4[base] = pick(0);
        o4[base + cv] = pick(1);
        o4[base + (size_t)W * cv] = pick(2);
        o4[base + (size_t)W * cv + cv] = pick(3);
    }
}

static inline int pool_grid(long long total) {
    long long g = (total + kPoolThreads - 1) / kPoolThreads;
    if (g < 1) g = 1;
    if (g > 148 * 8) g = 148 * 8;
    return (int)g;
}

cudaError_t launch_maxpool2_fwd(const float* x, float* y, unsigned char* arg, int N, int H, int W, int C, cudaStream_t stream) {
    const long long total = (long long)N * (H / 2) * (W / 2) * (C / 4);
    maxpool2_fwd_kernel<<<pool_grid(total), kPoolThreads, 0, stream>>>(x, y, arg, N, H, W, C);
    return cudaGetLastError();
}
cudaError_t launch_maxpool2_bwd(const float* dy, const unsigned char* arg, float* dx, int N, int H, int W, int C,
                                cudaStream_t stream) {
    const long long total = (long long)N * (H / 2) * (W / 2) * (C / 4);
    maxpool2_bwd_kernel<<<pool_grid(total), kPoolThreads, 0, stream>>>(dy, arg, dx, N, H, W, C);
    return cudaGetLastError();
}

}  // namespace okt
