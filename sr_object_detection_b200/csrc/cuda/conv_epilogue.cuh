// Shared epilogue piece of the slab / pair convolution kernels.
#pragma once
#include "conv_plan.cuh"

namespace y2 {

// one 32-column chunk of one accumulator row: affine + activation + store
template <int ACT>
__device__ __forceinline__ void slab_epilogue_chunk(const SlabParams &prm, const uint32_t (&v)[32], const float2 *sab,
                                                    int c0, int n0, int p, int b, int y, int x, bool in_range,
                                                    bool valid)
{
    float f[32];
    const float4 *ab4 = reinterpret_cast<const float4 *>(sab + c0);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float4 q = ab4[j];  // (alpha, beta) of two filters
        float t0 = fmaf(__uint_as_float(v[2 * j]), q.x, q.y);
        float t1 = fmaf(__uint_as_float(v[2 * j + 1]), q.z, q.w);
        if (ACT == Y2_ACT_LEAKY) {  // max(t, 0.1 t) == (t > 0 ? t : 0.1 t) for every finite t
            t0 = fmaxf(t0, 0.1f * t0);
            t1 = fmaxf(t1, 0.1f * t1);
        } else if (ACT == Y2_ACT_LOGISTIC) {
            t0 = 1.f / (1.f + __expf(-t0));
            t1 = 1.f / (1.f + __expf(-t1));
        }
        f[2 * j] = t0;
        f[2 * j + 1] = t1;
    }
    const int ch0 = n0 + c0;
    if (prm.out_mode == Y2_OUT_BF16_PADDED) {
        if (in_range) {
            __nv_bfloat16 *o = reinterpret_cast<__nv_bfloat16 *>(prm.out) + (size_t)p * prm.out_cs + ch0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (ch0 + q * 8 < prm.cout) {
                    uint4 w;
                    if (valid) {
                        w.x = pack_bf16x2(f[q * 8 + 0], f[q * 8 + 1]);
                        w.y = pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]);
                        w.z = pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]);
                        w.w = pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]);
                    } else {
                        w = make_uint4(0u, 0u, 0u, 0u);
                    }
                    *reinterpret_cast<uint4 *>(o + q * 8) = w;
                }
            }
        }
    } else {  // Y2_OUT_F32_FLAT: [B][h*w][out_cs]
        if (valid) {
            float *o = reinterpret_cast<float *>(prm.out) +
                       ((size_t)b * prm.h * prm.w + (size_t)y * prm.w + x) * prm.out_cs + ch0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (ch0 + j < prm.cout) o[j] = f[j];
        }
    }
}

// affine + activation of one 32-column chunk, packed to bf16 (zeros when the row is a pad position)
template <int ACT>
__device__ __forceinline__ void slab_affine_pack(const uint32_t (&v)[32], const float2 *sab, int c0, bool valid,
                                                 uint4 *w)
{
    const float4 *ab4 = reinterpret_cast<const float4 *>(sab + c0);
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float4 q = ab4[j];  // (alpha, beta) of two filters
        float t0 = fmaf(__uint_as_float(v[2 * j]), q.x, q.y);
        float t1 = fmaf(__uint_as_float(v[2 * j + 1]), q.z, q.w);
        if (ACT == Y2_ACT_LEAKY) {
            t0 = fmaxf(t0, 0.1f * t0);
            t1 = fmaxf(t1, 0.1f * t1);
        } else if (ACT == Y2_ACT_LOGISTIC) {
            t0 = 1.f / (1.f + __expf(-t0));
            t1 = 1.f / (1.f + __expf(-t1));
        }
        pk[j] = valid ? pack_bf16x2(t0, t1) : 0u;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) w[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
}

// 32 rows x 64 channels of one warp -> its 4 KB staging tile in the SWIZZLE_128B layout of the output
// tensor map (16-byte chunk c of row r at chunk c ^ (r & 7): conflict free), then ONE TMA store: the
// copy engine writes full 128-byte lines, the LSU sees 8 shared-memory stores per thread instead of 8
// global stores that touch 32 different lines each.
__device__ __forceinline__ void slab_store_tma(const void *tm_out, uint4 *stage, const uint4 (&w)[8], int lane,
                                               int p_first, int ch0)
{
    // the previous box of this warp must have been read out of the staging tile
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 8; ++c) stage[lane * 8 + (c ^ (lane & 7))] = w[c];
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        tma_store_2d(tm_out, stage, ch0, p_first);
        tma_store_commit();
    }
}

} // namespace y2
