// dcn_backward.cuh — DCNv2 backward, batched over N, both contractions on tcgen05.
//
// Replaces modulated_deform_conv_cuda_backward and its three kernels
// (/root/reference/basicsr/models/ops/dcn/src/deform_conv_cuda.cpp:571-685,
//  deform_conv_cuda_kernel.cu:499-568,635-767).  Stages (host orchestration in capi.cu):
//   1. gcol[p][(k,c)] = sum_co W[co][c][k] * gO[p][co]      -> a 1x1 implicit-GEMM conv over NHWC gO
//      (conv_igemm.cuh), gcol kept in fp16 NHWC: every consumer reads it pixel-major.
//   2. coord+scatter kernel: one thread per (pixel, deformable group, tap) reads its gcol slice once and
//      produces grad_offset, grad_mask (direct stores) and the grad_input scatter (vector red.add into an
//      NHWC fp32 buffer) — the reference needs two kernels and two passes over the columns.
//   3. col^T[(k,c)][p] (fp16, pixel-contiguous) + gO^T[co][p] -> split-K tcgen05 GEMM over pixels,
//      accumulated into grad_weight with fp32 atomics; grad_bias by a plain reduction.
// grad_offset follows the reference at the -1 edge: a sample coordinate <= -1 or >= H/W yields 0.
#pragma once
#include "common.cuh"
#include "elementwise.cuh"

namespace eb {

struct DcnBwdGeom {
    int N, C, H, W, Cout, kh, kw, stride, pad, dil, dg, Ho, Wo;   // stride / pad / dil along H
    int stride_w, pad_w, dil_w;                                  // along W
};

// ---- W^T packed as the weight of a 1x1 conv with Cin = Cout64 and packed rows r = k*C + c
__global__ void pack_wT_kernel(const float* __restrict__ w, int Cout, int C, int K, int Cout64, int BN,
                               __half* __restrict__ out) {
    const int nchunks = Cout64 / 64, ntiles = K * C / BN;
    const long long total = static_cast<long long>(ntiles) * nchunks * 8 * BN;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long r = i;
        const int n = r % BN; r /= BN;
        const int kc = r % 8; r /= 8;
        const int chunk = r % nchunks; r /= nchunks;
        const int row = static_cast<int>(r) * BN + n;
        const int k = row / C, c = row % C;
        H8 v;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int co = chunk * 64 + kc * 8 + e;
            v.v[e] = co < Cout ? w[(static_cast<size_t>(co) * C + c) * K + k] : 0.f;
        }
        h8_store(out + i * 8, v);
    }
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

// ---- stage 2: grad_offset, grad_mask, grad_input scatter
__global__ void dcn_bwd_coord_scatter_kernel(const DcnBwdGeom G, const __half* __restrict__ x16,
                                             const float* __restrict__ offset, const float* __restrict__ mask,
                                             const __half* __restrict__ gcol16, float* __restrict__ gx32,
                                             float* __restrict__ grad_offset, float* __restrict__ grad_mask) {
    const int K = G.kh * G.kw, HW = G.Ho * G.Wo, cpg = G.C / G.dg;
    const long long total = static_cast<long long>(G.N) * G.dg * K * HW;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int p = i % HW;
        const int k = (i / HW) % K;
        const int g = (i / HW / K) % G.dg;
        const int n = i / HW / K / G.dg;
        const int ho = p / G.Wo, wo = p - ho * G.Wo;
        const size_t ob = (static_cast<size_t>(n) * G.dg + g) * 2 * K * HW + p;
        const float dh = __ldg(offset + ob + static_cast<size_t>(2 * k) * HW);
        const float dw = __ldg(offset + ob + static_cast<size_t>(2 * k + 1) * HW);
        const size_t mb = ((static_cast<size_t>(n) * G.dg + g) * K + k) * HW + p;
        const float mk = mask ? __ldg(mask + mb) : 1.f;
        const int ki = k / G.kw, kj = k - ki * G.kw;
        const float h_im = static_cast<float>(ho * G.stride - G.pad + ki * G.dil) + dh;
        const float w_im = static_cast<float>(wo * G.stride_w - G.pad_w + kj * G.dil_w) + dw;
        float goh = 0.f, gow = 0.f, gm = 0.f;
        if (h_im > -1.f && w_im > -1.f && h_im < static_cast<float>(G.H) && w_im < static_cast<float>(G.W)) {
            const float hf = floorf(h_im), wf = floorf(w_im);
            const int hl = static_cast<int>(hf), wl = static_cast<int>(wf);
            const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
            const bool t = hl >= 0, b = hl + 1 <= G.H - 1, l = wl >= 0, r = wl + 1 <= G.W - 1;
            const long long base = ((static_cast<long long>(n) * G.H + hl) * G.W + wl) * G.C + g * cpg;
            const long long dW = G.C, dH = static_cast<long long>(G.W) * G.C;
            const __half* gc_ptr = gcol16 + (static_cast<size_t>(n) * HW + p) * (static_cast<size_t>(K) * G.C) +
                                   static_cast<size_t>(k) * G.C + g * cpg;
            for (int c8 = 0; c8 < cpg; c8 += 8) {
                const H8 gc = h8_load(gc_ptr + c8);
                H8 z;
#pragma unroll
                for (int e = 0; e < 8; ++e) z.v[e] = 0.f;
                const H8 v1 = (t && l) ? h8_load(x16 + base + c8) : z;
                const H8 v2 = (t && r) ? h8_load(x16 + base + dW + c8) : z;
                const H8 v3 = (b && l) ? h8_load(x16 + base + dH + c8) : z;
                const H8 v4 = (b && r) ? h8_load(x16 + base + dH + dW + c8) : z;
                float s[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float gce = gc.v[e];
                    gm = fmaf(gce, hh * hw * v1.v[e] + hh * lw * v2.v[e] + lh * hw * v3.v[e] + lh * lw * v4.v[e], gm);
                    goh = fmaf(gce, hw * (v3.v[e] - v1.v[e]) + lw * (v4.v[e] - v2.v[e]), goh);
                    gow = fmaf(gce, hh * (v2.v[e] - v1.v[e]) + lh * (v4.v[e] - v3.v[e]), gow);
                    s[e] = gce * mk;
                }
                float* gx = gx32 + base + c8;
                if (t && l) { const float q = hh * hw; red_add_v4(gx, q * s[0], q * s[1], q * s[2], q * s[3]); red_add_v4(gx + 4, q * s[4], q * s[5], q * s[6], q * s[7]); }
                if (t && r) { const float q = hh * lw; red_add_v4(gx + dW, q * s[0], q * s[1], q * s[2], q * s[3]); red_add_v4(gx + dW + 4, q * s[4], q * s[5], q * s[6], q * s[7]); }
                if (b && l) { const float q = lh * hw; red_add_v4(gx + dH, q * s[0], q * s[1], q * s[2], q * s[3]); red_add_v4(gx + dH + 4, q * s[4], q * s[5], q * s[6], q * s[7]); }
                if (b && r) { const float q = lh * lw; red_add_v4(gx + dH + dW, q * s[0], q * s[1], q * s[2], q * s[3]); red_add_v4(gx + dH + dW + 4, q * s[4], q * s[5], q * s[6], q * s[7]); }
            }
            goh *= mk;
            gow *= mk;
        }
        grad_offset[ob + static_cast<size_t>(2 * k) * HW] = goh;
        grad_offset[ob + static_cast<size_t>(2 * k + 1) * HW] = gow;
        if (grad_mask != nullptr) grad_mask[mb] = gm;
    }
}

// ---- NHWC fp32 -> NCHW fp32 (grad_input back to the reference layout)
__global__ void nhwc_f32_to_nchw_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int HW) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int p = p0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (p < HW && c < C) ? src[(static_cast<size_t>(n) * HW + p) * C + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, p = p0 + threadIdx.x;
        if (c < C && p < HW) dst[(static_cast<size_t>(n) * C + c) * HW + p] = tile[threadIdx.x][i];
    }
}

// ---- stage 3a: col^T[(k*C + c)][n*HW + p] fp16 (masked bilinear samples), pixel-contiguous rows
__global__ void dcn_bwd_colT_kernel(const DcnBwdGeom G, const __half* __restrict__ x16,
                                    const float* __restrict__ offset, const float* __restrict__ mask,
                                    __half* __restrict__ colT, long long Ppad) {
    const int K = G.kh * G.kw, HW = G.Ho * G.Wo, cpg = G.C / G.dg, C8 = G.C / 8;
    const long long total = static_cast<long long>(G.N) * C8 * K * HW;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int p = i % HW;
        const int k = (i / HW) % K;
        const int c8 = (i / HW / K) % C8;
        const int n = i / HW / K / C8;
        const int g = (c8 * 8) / cpg;
        const int ho = p / G.Wo, wo = p - ho * G.Wo;
        const size_t ob = (static_cast<size_t>(n) * G.dg + g) * 2 * K * HW + p;
        const float dh = __ldg(offset + ob + static_cast<size_t>(2 * k) * HW);
        const float dw = __ldg(offset + ob + static_cast<size_t>(2 * k + 1) * HW);
        const float mk = mask ? __ldg(mask + ((static_cast<size_t>(n) * G.dg + g) * K + k) * HW + p) : 1.f;
        const int ki = k / G.kw, kj = k - ki * G.kw;
        const float h_im = static_cast<float>(ho * G.stride - G.pad + ki * G.dil) + dh;
        const float w_im = static_cast<float>(wo * G.stride_w - G.pad_w + kj * G.dil_w) + dw;
        float val[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (h_im > -1.f && w_im > -1.f && h_im < static_cast<float>(G.H) && w_im < static_cast<float>(G.W)) {
            const float hf = floorf(h_im), wf = floorf(w_im);
            const int hl = static_cast<int>(hf), wl = static_cast<int>(wf);
            const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
            const bool t = hl >= 0, b = hl + 1 <= G.H - 1, l = wl >= 0, r = wl + 1 <= G.W - 1;
            const long long base = ((static_cast<long long>(n) * G.H + hl) * G.W + wl) * G.C + c8 * 8;
            const long long dW = G.C, dH = static_cast<long long>(G.W) * G.C;
            H8 z;
#pragma unroll
            for (int e = 0; e < 8; ++e) z.v[e] = 0.f;
            const H8 v1 = (t && l) ? h8_load(x16 + base) : z;
            const H8 v2 = (t && r) ? h8_load(x16 + base + dW) : z;
            const H8 v3 = (b && l) ? h8_load(x16 + base + dH) : z;
            const H8 v4 = (b && r) ? h8_load(x16 + base + dH + dW) : z;
#pragma unroll
            for (int e = 0; e < 8; ++e)
                val[e] = (hh * hw * v1.v[e] + hh * lw * v2.v[e] + lh * hw * v3.v[e] + lh * lw * v4.v[e]) * mk;
        }
        __half* o = colT + (static_cast<size_t>(k) * G.C + c8 * 8) * Ppad + static_cast<size_t>(n) * HW + p;
#pragma unroll
        for (int e = 0; e < 8; ++e) o[static_cast<size_t>(e) * Ppad] = __float2half_rn(val[e]);
    }
}

// ---- stage 3b: gO^T[co][n*HW + p] fp16 (rows >= Cout and the row tail stay zero: buffer is pre-zeroed)
__global__ void dcn_bwd_goT_kernel(const float* __restrict__ go, __half* __restrict__ goT, int N, int Cout,
                                   int HW, long long Ppad) {
    const long long total = static_cast<long long>(N) * Cout * HW;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int p = i % HW;
        const int co = (i / HW) % Cout;
        const int n = i / HW / Cout;
        goT[static_cast<size_t>(co) * Ppad + static_cast<size_t>(n) * HW + p] = __float2half_rn(go[i]);
    }
}

// ---- stage 3c: grad_weight[co][c][k] += sum_p gO^T[co][p] * col^T[k*C+c][p]   (split-K over pixels)
constexpr int WG_STAGES = 3;
constexpr int WG_SMEM_BYTES = WG_STAGES * (128 * 128 + 128 * 128);

__global__ void __launch_bounds__(128, 1)
dcn_bwd_wgrad_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ gW,
                     int Cout, int C, int K, long long Ppad, int BN, int steps_per_split, float scale) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t stage_free[WG_STAGES];
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int nt = blockIdx.x, mt = blockIdx.y;
    const long long nsteps_total = Ppad / 64;
    const long long s0 = static_cast<long long>(blockIdx.z) * steps_per_split;
    long long s1 = s0 + steps_per_split;
    if (s1 > nsteps_total) s1 = nsteps_total;
    const int nsteps = static_cast<int>(s1 - s0);

    if (tid == 0) {
        for (int i = 0; i < WG_STAGES; ++i) mbar_init(&stage_free[i], 1);
        mbar_init(&done_bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 128);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t idesc = umma_idesc_f16(128, BN);
    const uint32_t lbo_b = static_cast<uint32_t>(BN) * 16u;

    const __half* Arow = A + (static_cast<size_t>(mt) * 128 + tid) * Ppad;   // one A row per thread
    for (int i = 0; i < nsteps; ++i) {
        const int s = i % WG_STAGES;
        if (i >= WG_STAGES) mbar_wait(&stage_free[s], ((i / WG_STAGES) - 1) & 1);
        uint8_t* a_s = smem + s * (2 * 128 * 128);
        uint8_t* b_s = a_s + 128 * 128;
        const long long p0 = (s0 + i) * 64;
        uint4 va[8];
#pragma unroll
        for (int kc = 0; kc < 8; ++kc) va[kc] = ldg_nc_v4(Arow + p0 + kc * 8);
#pragma unroll
        for (int kc = 0; kc < 8; ++kc) sts_v4(smem_u32(a_s) + kc * 2048 + tid * 16, va[kc]);
        if (tid < BN) {
            const __half* Brow = B + (static_cast<size_t>(nt) * BN + tid) * Ppad + p0;
#pragma unroll
            for (int kc = 0; kc < 8; ++kc) va[kc] = ldg_nc_v4(Brow + kc * 8);
#pragma unroll
            for (int kc = 0; kc < 8; ++kc) sts_v4(smem_u32(b_s) + kc * lbo_b + tid * 16, va[kc]);
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after_sync();
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16) {
                const uint64_t ad = umma_desc_nosw(smem_u32(a_s) + k16 * 2 * 2048, 2048, 128);
                const uint64_t bd = umma_desc_nosw(smem_u32(b_s) + k16 * 2 * lbo_b, lbo_b, 128);
                umma_f16(tmem_base, ad, bd, idesc, (i | k16) != 0 ? 1u : 0u);
            }
            umma_commit(&stage_free[s]);
        }
    }
    if (tid == 0) umma_commit(&done_bar);
    if (nsteps > 0) {
        mbar_wait(&done_bar, 0);
        tc_fence_after_sync();
        const int co = mt * 128 + tid;
        for (int cc = 0; cc < BN; cc += 32) {
            float v[32];
            tmem_ld32(tmem_base + (static_cast<uint32_t>(32 * warp) << 16) + cc, v);
            if (co < Cout) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = nt * BN + cc + j;
                    const int k = n / C, c = n - k * C;
                    atomicAdd(gW + (static_cast<size_t>(co) * C + c) * K + k, v[j] * scale);
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// ---- grad_bias[co] += sum_{n,p} gO[n][co][p]; one block per output channel
__global__ void dcn_bwd_bias_kernel(const float* __restrict__ go, float* __restrict__ gb, int N, int Cout, int HW) {
    const int co = blockIdx.x;
    float s = 0.f;
    for (int n = 0; n < N; ++n) {
        const float* r = go + (static_cast<size_t>(n) * Cout + co) * HW;
        for (int p = threadIdx.x; p < HW; p += blockDim.x) s += r[p];
    }
    __shared__ float red[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) gb[co] += s;
    }
}

}  // namespace eb
