// First-layer kernel: 3x3 stride-1 'same' convolution over a C_in <= 3 fp32 NCHW image, fused with
// batchnorm + bias + leaky-ReLU AND the 2x2/2 maxpool that follows it in every north-star cfg.
//
// Replaces, for layers 0 and 1 of yolo-voc / tiny-yolo-voc / yolo9000 / darknet19:
//   cuda_make_array(input) + fill + im2col_ongpu + gemm_ongpu + normalize + scale_bias + add_bias +
//   activate_array (convolutional_kernels.cu:77-131) + forward_maxpool_layer_kernel
//   (maxpool_layer_kernels.cu:10-48)
// by one persistent launch that reads the caller's fp32 planar image once and writes only the
// pooled bf16 padded-NHWC tensor.  The layer is HBM-bound (K = 27): per image it must move
// 3*H*W*4 bytes in and (H/2)*(W/2)*32*2 bytes out; the full-resolution activation (H*W*32) never
// exists in memory.
//
// Work decomposition.  A tile is 16 x 8 pool windows (32 x 16 pixels).  Four producer warps stage
// the 34 x 18 x 3 fp32 halo patch with cp.async (double buffered), then each producer thread owns
// one window and writes the four im2col rows of its 2x2 pixels (K = c*9 + r*3 + s, the order of
// im2col.c:16-39, padded 27 -> 32) as bf16 into four 128 x 32 K-major SWIZZLE_64B operand tiles
// A_q, q = 2*dy + dx.  One thread issues 4 x 2 tcgen05.mma (M = 128, N = 32, K = 16) into four TMEM
// accumulators D_q.  Four epilogue warps read the four accumulators of their window from their own
// TMEM lane, take the max, apply y = leaky(alpha * m + beta) and store 64 bytes.
//
// max before the affine map is exact: the host makes every alpha_f >= 0 (a filter with negative
// alpha has its weights and alpha negated, which leaves alpha*acc bit-identical), and
// x -> leaky(fma(alpha, x, beta)) is then non-decreasing in fp32, so it commutes with max.
#include "y2_common.cuh"

namespace y2 {

constexpr int kStemThreads = 288;  // warps 0-3 producers, 4 MMA, 5-8 epilogue
constexpr int kStemN = 32;         // filters (padded)
constexpr int kStemK = 32;         // 27 taps*channels padded
constexpr int kWinX = 16, kWinY = 8;
constexpr int kPatchW = 2 * kWinX + 2, kPatchH = 2 * kWinY + 2;  // 34 x 18
constexpr int kPitch = 36;                                        // floats per staged row
constexpr int kStageFloats = 3 * kPatchH * kPitch;
constexpr int kATile = 128 * kStemK * 2;  // 8 KB per sub-position
constexpr uint32_t kStemIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kStemN >> 3) << 17) |
                                ((uint32_t)(128 >> 4) << 24);

struct StemParams {
    const float *in;   // fp32 [B][c][h][w]
    int batch, c, h, w;
    int oh, ow;        // pooled extent
    int tiles_x, tiles_y, total_tiles;
    const __nv_bfloat16 *wt;  // [32][32], K-major
    const float *alpha, *beta;
    int act;
    __nv_bfloat16 *out;  // padded NHWC [B][oh+1][ow+1][out_cs]
    int out_cs;
};

struct StemSmem {
    alignas(1024) uint8_t a[2][4][kATile];  // 64 KB, swizzle atoms need 512 B alignment
    alignas(1024) uint8_t w[kStemN * kStemK * 2];
    alignas(16) float stage[2][kStageFloats];
    float alpha[kStemN], beta[kStemN];
    alignas(8) uint64_t a_full[2], a_empty[2], t_full[2], t_empty[2];
    uint32_t tmem_slot;
};

__device__ __forceinline__ void cp_async_f32(float *dst, const float *src, bool valid)
{
    const uint32_t n = valid ? 4u : 0u;  // src-size 0 -> zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(n) : "memory");
}

// byte offset of 16-byte chunk `chunk` of row `row` in a K-major SWIZZLE_64B tile (64-byte rows):
// the hardware XORs address bits [4,6) with bits [7,9)
__device__ __forceinline__ uint32_t swz64(int row, int chunk)
{
    return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}

__device__ __forceinline__ void stem_issue_loads(const StemParams &p, float *stage, int tile, int ptid)
{
    const int per_img = p.tiles_x * p.tiles_y;
    const int b = tile / per_img;
    const int t = tile - b * per_img;
    const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
    const int y0 = ty * (2 * kWinY) - 1, x0 = tx * (2 * kWinX) - 1;
    const float *img = p.in + (size_t)b * p.c * p.h * p.w;
    for (int e = ptid; e < 3 * kPatchH * kPatchW; e += 128) {
        const int ci = e / (kPatchH * kPatchW);
        const int rem = e - ci * (kPatchH * kPatchW);
        const int ry = rem / kPatchW, rx = rem - ry * kPatchW;
        const int gy = y0 + ry, gx = x0 + rx;
        const bool valid = ci < p.c && gy >= 0 && gy < p.h && gx >= 0 && gx < p.w;
        const float *src = valid ? img + ((size_t)ci * p.h + gy) * p.w + gx : p.in;
        cp_async_f32(stage + (ci * kPatchH + ry) * kPitch + rx, src, valid);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(kStemThreads, 2) stem_conv_pool_kernel(const StemParams p)
{
    extern __shared__ uint8_t stem_raw[];
    StemSmem &sm = *reinterpret_cast<StemSmem *>(stem_raw + ((1024u - (smem_u32(stem_raw) & 1023u)) & 1023u));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sm.a_full[i], 128);
            mbar_init(&sm.a_empty[i], 1);
            mbar_init(&sm.t_full[i], 1);
            mbar_init(&sm.t_empty[i], 128);
        }
        fence_barrier_init();
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_slot)),
                     "r"(256u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // weights -> swizzled smem (generic proxy writes, fenced for the tensor core's async proxy)
    if (threadIdx.x < 128) {
        const int row = threadIdx.x >> 2, chunk = threadIdx.x & 3;
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p.wt + row * kStemK + chunk * 8));
        *reinterpret_cast<uint4 *>(sm.w + swz64(row, chunk)) = v;
    } else if (threadIdx.x < 128 + kStemN) {
        const int f = threadIdx.x - 128;
        sm.alpha[f] = p.alpha[f];
        sm.beta[f] = p.beta[f];
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_slot;

    if (warp < 4) {
        // ===================== producers: stage the patch, build the im2col rows =====================
        const int ptid = threadIdx.x;  // 0..127 == window index inside the tile
        const int wy = ptid >> 4, wx = ptid & 15;
        int it = 0;
        if ((int)blockIdx.x < p.total_tiles) stem_issue_loads(p, sm.stage[0], blockIdx.x, ptid);
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int s = it & 1;
            const int next = tile + gridDim.x;
            if (next < p.total_tiles) stem_issue_loads(p, sm.stage[s ^ 1], next, ptid);
            else asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
            // 4 x 4 pixel neighbourhood of this window, 3 channels
            float v[3][4][4];
            const float *st = sm.stage[s];
#pragma unroll
            for (int ci = 0; ci < 3; ++ci)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 *src =
                        reinterpret_cast<const float2 *>(st + (ci * kPatchH + 2 * wy + j) * kPitch + 2 * wx);
                    const float2 lo = src[0], hi = src[1];
                    v[ci][j][0] = lo.x; v[ci][j][1] = lo.y; v[ci][j][2] = hi.x; v[ci][j][3] = hi.y;
                }
            mbar_wait(&sm.a_empty[s], ((it >> 1) & 1) ^ 1, 10);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int dy = q >> 1, dx = q & 1;
                uint32_t w32[16];
#pragma unroll
                for (int k2 = 0; k2 < 16; ++k2) {
                    float e[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int k = 2 * k2 + h;  // K index = ci*9 + r*3 + s
                        e[h] = (k < 27) ? v[k / 9][dy + (k % 9) / 3][dx + k % 3] : 0.f;
                    }
                    w32[k2] = pack_bf16x2(e[0], e[1]);
                }
                uint8_t *dst = sm.a[s][q];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<uint4 *>(dst + swz64(ptid, c)) =
                        make_uint4(w32[4 * c], w32[4 * c + 1], w32[4 * c + 2], w32[4 * c + 3]);
            }
            fence_proxy_async();
            mbar_arrive(&sm.a_full[s]);
            asm volatile("bar.sync 1, 128;" ::: "memory");  // staging buffer s is free for tile it+2
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if (warp == 4) {
        // ===================== MMA issuer (whole warp waits, one elected lane issues) ===========
        const uint64_t bdesc = make_kmajor_desc(smem_u32(sm.w), 512, 4);
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int s = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            mbar_wait(&sm.t_empty[s], ph ^ 1, 11);
            mbar_wait(&sm.a_full[s], ph, 12);
            tc_fence_after();
            if (elect_one_sync()) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint64_t adesc = make_kmajor_desc(smem_u32(sm.a[s][q]), 512, 4);
                    const uint32_t d = tmem_base + (uint32_t)(s * 4 * kStemN + q * kStemN);
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        umma_bf16(d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), kStemIdesc, (uint32_t)k);
                }
                umma_commit(&sm.a_empty[s]);
                umma_commit(&sm.t_full[s]);
            }
            __syncwarp();
        }
    } else {
        // ===================== epilogue: max over the window, affine, leaky, store =====================
        const int quarter = warp & 3;
        const int win = quarter * 32 + lane;
        const int wy = win >> 4, wx = win & 15;
        const int per_img = p.tiles_x * p.tiles_y;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int s = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            mbar_wait(&sm.t_full[s], ph, 13);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * 4 * kStemN);
            uint32_t m[32], u[32];
            tmem_ld32(taddr, m);
            tmem_ld32(taddr + kStemN, u);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) m[j] = __float_as_uint(fmaxf(__uint_as_float(m[j]), __uint_as_float(u[j])));
            tmem_ld32(taddr + 2 * kStemN, u);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) m[j] = __float_as_uint(fmaxf(__uint_as_float(m[j]), __uint_as_float(u[j])));
            tmem_ld32(taddr + 3 * kStemN, u);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&sm.t_empty[s]);
#pragma unroll
            for (int j = 0; j < 32; ++j) m[j] = __float_as_uint(fmaxf(__uint_as_float(m[j]), __uint_as_float(u[j])));

            const int b = tile / per_img;
            const int t = tile - b * per_img;
            const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
            const int oy = ty * kWinY + wy, ox = tx * kWinX + wx;
            if (oy < p.oh && ox < p.ow) {
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float y = fmaf(__uint_as_float(m[j]), sm.alpha[j], sm.beta[j]);
                    if (p.act == Y2_ACT_LEAKY) y = (y > 0.f) ? y : 0.1f * y;
                    f[j] = y;
                }
                __nv_bfloat16 *o = p.out + (((size_t)b * (p.oh + 1) + oy) * (p.ow + 1) + ox) * p.out_cs;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint4 w;
                    w.x = pack_bf16x2(f[8 * c + 0], f[8 * c + 1]);
                    w.y = pack_bf16x2(f[8 * c + 2], f[8 * c + 3]);
                    w.z = pack_bf16x2(f[8 * c + 4], f[8 * c + 5]);
                    w.w = pack_bf16x2(f[8 * c + 6], f[8 * c + 7]);
                    *reinterpret_cast<uint4 *>(o + 8 * c) = w;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

} // namespace y2

using namespace y2;

// raise the dynamic shared-memory limit (plan time, outside any graph capture)
extern "C" int y2_stem_prepare(void)
{
    static bool attr_done[64] = {false};
    int dev = 0;
    Y2_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
        Y2_CUDA_CHECK(cudaFuncSetAttribute(stem_conv_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(sizeof(StemSmem) + 1024)));
        attr_done[dev] = true;
    }
    return Y2_OK;
}

extern "C" int y2_stem_conv_pool(const float *in, int batch, int c, int h, int w, const void *wt, int npad,
                                 const float *alpha, const float *beta, int act, void *out, int out_cs,
                                 y2_stream_t s)
{
    if (!in || !wt || !alpha || !beta || !out || batch <= 0 || c <= 0 || c > 3 || h < 2 || w < 2 || npad != kStemN ||
        out_cs % 8 || out_cs < kStemN || ((uintptr_t)out & 15) || ((uintptr_t)wt & 15) ||
        (act != Y2_ACT_LEAKY && act != Y2_ACT_LINEAR)) {
        set_error("y2_stem_conv_pool: invalid arguments (c=%d h=%d w=%d npad=%d out_cs=%d act=%d)", c, h, w, npad,
                  out_cs, act);
        return Y2_EINVAL;
    }
    StemParams p;
    p.in = in;
    p.batch = batch;
    p.c = c;
    p.h = h;
    p.w = w;
    p.oh = h / 2;
    p.ow = w / 2;
    p.tiles_x = (p.ow + kWinX - 1) / kWinX;
    p.tiles_y = (p.oh + kWinY - 1) / kWinY;
    const long long total = (long long)batch * p.tiles_x * p.tiles_y;
    if (total > 0x7fffffffLL) return Y2_EINVAL;
    p.total_tiles = (int)total;
    p.wt = (const __nv_bfloat16 *)wt;
    p.alpha = alpha;
    p.beta = beta;
    p.act = act;
    p.out = (__nv_bfloat16 *)out;
    p.out_cs = out_cs;
    const size_t smem = sizeof(StemSmem) + 1024;
    int rc = y2_stem_prepare();
    if (rc != Y2_OK) return rc;
    const int sms = sm_count();
    // two co-resident CTAs per SM (83 KB smem, 256 TMEM columns, <= 112 registers each): the producer warps
    // of one CTA are latency-bound, a second CTA fills the bubbles
    const int grid = p.total_tiles < 2 * sms ? p.total_tiles : 2 * sms;
    stem_conv_pool_kernel<<<grid, kStemThreads, smem, to_stream(s)>>>(p);
    Y2_LAUNCH_CHECK();
    return Y2_OK;
}
