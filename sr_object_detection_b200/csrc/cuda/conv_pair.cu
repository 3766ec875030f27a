// Implicit-GEMM 3x3 convolution on CTA pairs: tcgen05.mma.cta_group::2 over a 2-CTA cluster.
//
// Replaces forward_convolutional_layer_gpu (reference convolutional_kernels.cu:77-131) for the wide
// 3x3 layers (>= 256 filters, C_in a multiple of 64) - two thirds of YOLOv2's FLOPs.
//
// Why pairs.  With one CTA per tile (conv_slab.cu) a 128 x 256 x 64 MMA step moves 34 KB into shared
// memory and reads 48 KB of operands back out of it in 512 tensor-pipe clocks: the ~128 B/clk/SM of
// shared-memory bandwidth, not the tensor pipe, bounds the layer at ~60% of peak.  A CTA pair
// computes a 256 x 256 tile with ONE instruction stream: each CTA holds its own 128 positions of the
// halo slab (A) and HALF of the 256-filter weight tile (B); the tensor cores of both SMs read the
// two halves across the pair.  Per SM and step that is 18 KB of fill and 32 KB of operand reads.
//
// Protocol (leader = even CTA of the pair):
//   * both CTAs TMA-load their slab / weight half into their own shared memory; the bytes are
//     credited to the LEADER's full barriers (.cta_group::2 TMA, peer bit cleared in the mbarrier
//     address), whose single arrival is the leader's expect_tx for both halves;
//   * only the leader issues MMAs; tcgen05.commit multicasts the "stage free" / "accumulator
//     full" arrivals to the barriers at the same offset in both CTAs;
//   * each CTA drains its own 128 TMEM lanes; "accumulator drained" arrivals of both CTAs land on
//     the leader's barrier (remote mbarrier.arrive through mapa).
// Everything else (halo slab, row-shifted descriptors, double-buffered accumulators, 8 epilogue
// warps) is conv_slab.cu's design.
//
// Work distribution.  Whole tiles round-robin leave the last wave partly empty when the tile count is
// a small non-multiple of the 74 pairs (13x13 layers at batch 64: 196 tiles = 2.65 waves, run as 3).
// With prm.streamk the K loops of all tiles are laid end to end in units of one channel block
// (64 input channels x 9 taps) and every pair takes an equal contiguous share: a pair's share starts
// with the tail of a tile, continues with whole tiles and ends with the head of a tile.  The pair that
// computes a tile's TAIL does so first thing and parks its raw fp32 accumulator in global scratch;
// the pair that computes the HEAD does so last, adds the parked partial in its epilogue and finishes
// the tile.  The dependency always points from a pair's first segment to another pair's last one, so
// with all pairs resident (grid = SM count) nothing can wait on work that has not been scheduled.
#include "conv_epilogue.cuh"

#include <stdlib.h>
#include <string.h>

namespace y2 {

constexpr int kPairThreads = 352;
constexpr int kPairEpiThreads = 256;
constexpr int kPairMaxStagesA = 4;
constexpr int kPairMaxStagesB = 10;
constexpr int kPairBK = 64;
constexpr int kPairN = 256;                          // filters per pair tile
constexpr int kPairRowBytes = kPairBK * 2;           // 128
constexpr int kPairBHalfBytes = 128 * kPairBK * 2;   // this CTA's half of a weight tile
constexpr uint32_t kPairDescHi = ((8u * kPairRowBytes) >> 4) | (1u << 14) | (2u << 29);  // SBO 1024, v1, SWIZZLE_128B
// c = F32, a = b = BF16, K-major, N = 256, M = 256 (both CTAs)
constexpr uint32_t kPairIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kPairN >> 3) << 17) |
                                ((uint32_t)(256 >> 4) << 24);

// one stretch of K iterations of one tile: channel blocks [cb0, cb1)
struct PairSeg {
    int tile, cb0, cb1;
};

struct PairSched {
    int streamk, n_pairs, cblocks, cur, end;
    __device__ PairSched(const SlabParams &prm, int pair, int n_pairs_)
        : streamk(prm.streamk), n_pairs(n_pairs_), cblocks(prm.cblocks)
    {
        const int total_tiles = prm.tiles_m * prm.tiles_n;
        if (streamk) {
            const long long units = (long long)total_tiles * cblocks;
            cur = (int)(units * pair / n_pairs);
            end = (int)(units * (pair + 1) / n_pairs);
        } else {
            cur = pair;
            end = total_tiles;
        }
    }
    __device__ bool next(PairSeg &s)
    {
        if (cur >= end) return false;
        if (streamk) {
            s.tile = cur / cblocks;
            s.cb0 = cur - s.tile * cblocks;
            const int left = end - cur, room = cblocks - s.cb0;
            const int n = left < room ? left : room;
            s.cb1 = s.cb0 + n;
            cur += n;
        } else {
            s.tile = cur;
            s.cb0 = 0;
            s.cb1 = cblocks;
            cur += n_pairs;
        }
        return true;
    }
    __device__ bool peek(PairSeg &s) const
    {
        PairSched copy = *this;
        return copy.next(s);
    }
};

__device__ __forceinline__ int ld_acquire_gpu(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_gpu(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// TAPS = 9: 3x3 layer (halo slab, nine row-shifted descriptors per channel block);
// TAPS = 1: 1x1 layer (the "slab" is the plain 128-position tile, one descriptor per channel block)
template <int TAPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_out, const SlabParams prm)
{
    extern __shared__ uint8_t smem_raw[];
    // identical carve-up in both CTAs: descriptors and barrier offsets name the peer's memory too
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

    const int stages_a = prm.stages_a, stages_b = prm.stages_b;
    uint8_t *smem_a = smem;
    uint8_t *smem_b = smem + (size_t)stages_a * prm.slab_bytes;
    uint8_t *aux = smem_b + (size_t)stages_b * kPairBHalfBytes;
    float2 *s_ab = reinterpret_cast<float2 *>(aux);  // [2 buf][256] (alpha, beta)
    uint64_t *bars = reinterpret_cast<uint64_t *>(aux + 2 * kPairN * 8);
    uint64_t *a_full = bars;
    uint64_t *a_empty = bars + kPairMaxStagesA;
    uint64_t *b_full = bars + 2 * kPairMaxStagesA;
    uint64_t *b_empty = b_full + kPairMaxStagesB;
    uint64_t *tfull_bar = b_empty + kPairMaxStagesB;
    uint64_t *tempty_bar = tfull_bar + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty_bar + 2);
    // [8 epilogue warps][32 rows][128 B] staging of the TMA stores, 1024-byte aligned (swizzle atom)
    uint4 *s_stage = reinterpret_cast<uint4 *>(aux + ((2 * kPairN * 8 + 512 + 1023) / 1024) * 1024);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1;
    const int n_pairs = gridDim.x >> 1;
    const int cblocks = prm.cblocks;
    PairSched sched(prm, pair, n_pairs);
    PairSeg seg;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a);
        tma_prefetch_desc(&tm_b);
        for (int i = 0; i < stages_a; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < stages_b; ++i) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 2 * kPairEpiThreads);  // the epilogue threads of both CTAs
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();  // barriers of both CTAs initialised before any remote arrival
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    pdl_launch_dependents();
    if (warp == 0) {
        // ===================== slab producer (both CTAs, own 128 positions) =====================
        pdl_wait();  // the activations are the previous layer's output; weights (warp 2) are prefetched meanwhile
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t slab_tx = (uint32_t)prm.slab_loads * prm.box_rows * kPairRowBytes;
        const uint32_t load_bytes = (uint32_t)prm.box_rows * kPairRowBytes;
        while (sched.next(seg)) {
            const int m_tile = seg.tile / prm.tiles_n;
            const int row0 = m_tile * 256 + (int)rank * 128 - prm.halo;
            for (int cb = seg.cb0; cb < seg.cb1; ++cb) {
                mbar_wait(&a_empty[stage], phase ^ 1, 1);
                if (elect_one_sync()) {
                    uint8_t *sa = smem_a + (size_t)stage * prm.slab_bytes;
                    if (leader) mbar_expect_tx(&a_full[stage], 2 * slab_tx);
                    for (int i = 0; i < prm.slab_loads; ++i)
                        tma_load_2d_pair(&tm_a, &a_full[stage], sa + i * load_bytes, cb * kPairBK,
                                         row0 + i * prm.box_rows);
                }
                __syncwarp();
                if (++stage == stages_a) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 2) {
        // ===================== weight producer (both CTAs, own 128 filters) =====================
        int stage = 0;
        uint32_t phase = 0;
        int issued = 0;
        while (sched.next(seg)) {
            const int m_tile = seg.tile / prm.tiles_n;
            const int n0 = (seg.tile - m_tile * prm.tiles_n) * kPairN + (int)rank * 128;
            for (int cb = seg.cb0; cb < seg.cb1; ++cb) {
#pragma unroll 1
                for (int tap = 0; tap < TAPS; ++tap) {
                    mbar_wait(&b_empty[stage], phase ^ 1, 2);
                    const bool skip = (prm.dbg & 1) && (tap & 1) && issued >= 2 * stages_b;
                    ++issued;
                    if (elect_one_sync()) {
                        if (skip) {  // timing experiment: the stage "completes" with stale bytes
                            if (leader) mbar_arrive(&b_full[stage]);
                        } else {
                            if (leader) mbar_expect_tx(&b_full[stage], 2u * kPairBHalfBytes);
                            tma_load_2d_pair(&tm_b, &b_full[stage], smem_b + (size_t)stage * kPairBHalfBytes,
                                             (tap * cblocks + cb) * kPairBK, n0);
                        }
                    }
                    __syncwarp();
                    if (++stage == stages_b) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            int sa_i = 0, sb_i = 0;
            uint32_t pa = 0, pb = 0;
            int it = 0;
            const uint32_t a_lo0 = ((smem_u32(smem_a) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t b_lo0 = ((smem_u32(smem_b) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t slab16 = (uint32_t)prm.slab_bytes >> 4;
            constexpr uint32_t kRow16 = kPairRowBytes >> 4;
            const uint32_t wp16 = (uint32_t)prm.wp * kRow16;
            for (; sched.next(seg); ++it) {
                const int buf = it & 1;
                const uint32_t buf_phase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[buf], buf_phase ^ 1, 3);
                tc_fence_after();
                const uint32_t d0 = tmem_base + (uint32_t)(buf * kPairN);
                for (int cb = seg.cb0; cb < seg.cb1; ++cb) {
                    mbar_wait(&a_full[sa_i], pa, 4);
                    const uint32_t a_lo = a_lo0 + (uint32_t)sa_i * slab16;
                    const uint32_t acc_first = cb != seg.cb0;
#pragma unroll
                    for (int tap = 0; tap < TAPS; ++tap) {
                        mbar_wait(&b_full[sb_i], pb, 5);
                        tc_fence_after();
                        if (elect_one_sync()) {
                            const uint32_t b_lo = b_lo0 + (uint32_t)sb_i * (kPairBHalfBytes >> 4);
                            const uint32_t a_tap = a_lo + (uint32_t)(tap / 3) * wp16 + (uint32_t)(tap % 3) * kRow16;
#pragma unroll
                            for (int k = 0; k < kPairBK / 16; ++k)
                                umma_bf16_pair(d0, ((uint64_t)kPairDescHi << 32) | (uint64_t)(a_tap + (uint32_t)(k * 2)),
                                               ((uint64_t)kPairDescHi << 32) | (uint64_t)(b_lo + (uint32_t)(k * 2)),
                                               kPairIdesc, (tap == 0 && k == 0) ? acc_first : 1u);
                            umma_commit_pair(&b_empty[sb_i]);
                            if (tap == TAPS - 1) {
                                umma_commit_pair(&a_empty[sa_i]);
                                if (cb == seg.cb1 - 1) umma_commit_pair(&tfull_bar[buf]);
                            }
                        }
                        __syncwarp();
                        if (++sb_i == stages_b) { sb_i = 0; pb ^= 1; }
                    }
                    if (++sa_i == stages_a) { sa_i = 0; pa ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue (warps 3..10 of both CTAs, own 128 TMEM lanes) ==========
        const int quarter = warp & 3;
        const int half = (warp - 3) >> 2;
        const int et = threadIdx.x - 96;  // 0..255
        const int col0 = half * 128;
        const int img_pos = prm.hp * prm.wp;
        int it = 0;
        const int ewarp = (int)rank * 8 + (warp - 3);  // 0..15 within the pair
        if (sched.peek(seg)) {
            const int n0 = (seg.tile % prm.tiles_n) * kPairN;
            s_ab[et] = make_float2(__ldg(prm.alpha + n0 + et), __ldg(prm.beta + n0 + et));
        }
        for (; sched.next(seg); ++it) {
            const int tile = seg.tile;
            const int buf = it & 1;
            const uint32_t buf_phase = (it >> 1) & 1;
            const int m_tile = tile / prm.tiles_n;
            const int n0 = (tile - m_tile * prm.tiles_n) * kPairN;
            // stream-K roles of this segment: the tail of a tile is parked, the head adds it and finishes
            const bool park = seg.cb0 > 0;
            const bool join = seg.cb0 == 0 && seg.cb1 < cblocks;
            const float2 *sab = s_ab + (prm.tiles_n > 1 ? buf * kPairN : 0);
            const bool reload = prm.tiles_n > 1 || it == 0;
            if (reload) asm volatile("bar.sync 1, 256;" ::: "memory");
            PairSeg nseg;
            const bool has_next = sched.peek(nseg);
            float2 ab_next = make_float2(1.f, 0.f);
            if (prm.tiles_n > 1 && has_next) {
                const int nn = (nseg.tile % prm.tiles_n) * kPairN;
                ab_next = make_float2(__ldg(prm.alpha + nn + et), __ldg(prm.beta + nn + et));
            }
            // scratch of a split tile belongs to the pair that parks it: this pair, or the next one
            float *part = prm.sk_partial + (size_t)(park ? pair : pair + 1) * (kPairN * 256) + (int)rank * 128 +
                          quarter * 32 + lane;
            int *flag = prm.sk_flags + (park ? pair : pair + 1) * 16 + ewarp;
            const int p = m_tile * 256 + (int)rank * 128 + quarter * 32 + lane;
            const bool in_range = p < prm.total_pos;
            const int b = p / img_pos;
            const int rem = p - b * img_pos;
            const int y = rem / prm.wp;
            const int x = rem - y * prm.wp;
            const bool valid = in_range && (y < prm.h) && (x < prm.w);
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * kPairN + col0);

            mbar_wait(&tfull_bar[buf], buf_phase, 6);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 128; c += 64) {
                uint32_t v0[32], v1[32];
                tmem_ld32(taddr + (uint32_t)c, v0);
                tmem_ld32(taddr + (uint32_t)(c + 32), v1);
                tmem_ld_wait();
                if (c == 64) {  // accumulator drained into registers: hand it back to the leader's MMA warp
                    tc_fence_before();
                    mbar_arrive_cluster(&tempty_bar[buf], 0);
                }
                if (park) {  // raw partial sums, [filter][position]: a warp stores 128 contiguous bytes per filter
                    float *dst = part + (size_t)(col0 + c) * 256;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        dst[j * 256] = __uint_as_float(v0[j]);
                        dst[(32 + j) * 256] = __uint_as_float(v1[j]);
                    }
                    continue;
                }
                if (join) {
                    if (c == 0) {
                        if (lane == 0)
                            while (ld_acquire_gpu(flag) == 0) __nanosleep(64);
                        __syncwarp();
                    }
                    const float *src = part + (size_t)(col0 + c) * 256;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        v0[j] = __float_as_uint(__uint_as_float(v0[j]) + __ldcg(src + j * 256));
                        v1[j] = __float_as_uint(__uint_as_float(v1[j]) + __ldcg(src + (32 + j) * 256));
                    }
                }
                if (prm.tma_store) {  // bf16 tensor: staged, one TMA store per warp and 64 channels
                    uint4 w[8];
                    if (prm.act == Y2_ACT_LEAKY) {
                        slab_affine_pack<Y2_ACT_LEAKY>(v0, sab, col0 + c, valid, w);
                        slab_affine_pack<Y2_ACT_LEAKY>(v1, sab, col0 + c + 32, valid, w + 4);
                    } else if (prm.act == Y2_ACT_LINEAR) {
                        slab_affine_pack<Y2_ACT_LINEAR>(v0, sab, col0 + c, valid, w);
                        slab_affine_pack<Y2_ACT_LINEAR>(v1, sab, col0 + c + 32, valid, w + 4);
                    } else {
                        slab_affine_pack<Y2_ACT_LOGISTIC>(v0, sab, col0 + c, valid, w);
                        slab_affine_pack<Y2_ACT_LOGISTIC>(v1, sab, col0 + c + 32, valid, w + 4);
                    }
                    slab_store_tma(&tm_out, s_stage + (warp - 3) * 256, w, lane, p - lane, n0 + col0 + c);
                } else if (prm.act == Y2_ACT_LEAKY) {
                    slab_epilogue_chunk<Y2_ACT_LEAKY>(prm, v0, sab, col0 + c, n0, p, b, y, x, in_range, valid);
                    slab_epilogue_chunk<Y2_ACT_LEAKY>(prm, v1, sab, col0 + c + 32, n0, p, b, y, x, in_range, valid);
                } else if (prm.act == Y2_ACT_LINEAR) {
                    slab_epilogue_chunk<Y2_ACT_LINEAR>(prm, v0, sab, col0 + c, n0, p, b, y, x, in_range, valid);
                    slab_epilogue_chunk<Y2_ACT_LINEAR>(prm, v1, sab, col0 + c + 32, n0, p, b, y, x, in_range, valid);
                } else {
                    slab_epilogue_chunk<Y2_ACT_LOGISTIC>(prm, v0, sab, col0 + c, n0, p, b, y, x, in_range, valid);
                    slab_epilogue_chunk<Y2_ACT_LOGISTIC>(prm, v1, sab, col0 + c + 32, n0, p, b, y, x, in_range, valid);
                }
            }
            if (park) {  // publish: every lane's stores, then the warp's flag
                __syncwarp();
                if (lane == 0) {
                    __threadfence();
                    st_release_gpu(flag, 1);
                }
            } else if (join) {  // consumed: leave the flag clear for the next launch
                __syncwarp();
                if (lane == 0) *reinterpret_cast<volatile int *>(flag) = 0;
            }
            if (prm.tiles_n > 1 && has_next) s_ab[(buf ^ 1) * kPairN + et] = ab_next;
        }
        if (prm.tma_store && lane == 0) tma_store_wait_all();  // the copies read this CTA's shared memory
    }

    tc_fence_before();
    cluster_sync_all();  // no CTA leaves while its peer may still address its shared memory / barriers
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// -------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------
int pair_plan_init(y2_conv_plan *pl, const y2_conv_desc *d)
{
    if ((d->ksize != 3 && d->ksize != 1) || d->block_k != kPairBK || d->block_n != 256 || d->npad % kPairN ||
        d->cin % kPairBK)
        return Y2_EINVAL;
    const int taps = d->ksize * d->ksize;
    const int hp = d->h + 1, wp = d->w + 1;
    const long long total = (long long)d->batch * hp * wp;
    // 1x1 layers: only bf16 tensors through the TMA-store epilogue, and only when there is enough work to
    // fill the pairs (small layers stay on the single-CTA kernels, which have more CTAs to spread over)
    if (taps == 1 && (d->out_mode != Y2_OUT_BF16_PADDED || d->cout % 64)) return Y2_EINVAL;
    // Measured (yolo-voc L13 / L19, batch 64): the one-tap pair kernel runs 20.9 / 21.0 us against 18.8 / 20.1
    // of the slab kernel - these layers are bound by L2 -> shared-memory operand traffic and by their few
    // tiles, not by the tensor pipe - so it is only used when asked for (Y2_CONV_VARIANT=pair; tests).
    if (taps == 1) {
        const char *forced = getenv("Y2_CONV_VARIANT");
        if (!forced || strcmp(forced, "pair")) return Y2_EINVAL;
    }
    const int halo = taps == 9 ? wp + 1 : 0;
    const int slab_rows = 128 + 2 * halo;
    const int loads = (slab_rows + 255) / 256;
    int box_rows = (slab_rows + loads - 1) / loads;
    box_rows = (box_rows + 15) / 16 * 16;
    if (box_rows > 256) return Y2_EINVAL;
    const int slab_bytes = loads * box_rows * kPairRowBytes;
    const int aux = ((2 * kPairN * 8 + 512 + 1023) / 1024) * 1024 + 8 * 4096;  // alpha/beta, barriers | store staging
    const int budget = 227 * 1024 - 1024 - aux;
    const int stages_a = taps == 9 ? 2 : kPairMaxStagesA;  // 1x1: a fresh A tile every K step
    int stages_b = (budget - stages_a * slab_bytes) / kPairBHalfBytes;
    if (stages_b > kPairMaxStagesB) stages_b = kPairMaxStagesB;
    if (stages_b < 4) return Y2_EINVAL;
    const int sms = sm_count();
    if (sms < 2) return Y2_EINVAL;
    const int ktot = taps * d->cin;
    int rc = encode_2d_bf16(&pl->tm_a, d->in, (uint64_t)d->cin, (uint64_t)total, (uint64_t)d->in_cs * 2,
                            (uint32_t)kPairBK, (uint32_t)box_rows, kPairBK);
    if (rc == Y2_OK)
        rc = encode_2d_bf16(&pl->tm_b, d->wt, (uint64_t)ktot, (uint64_t)d->npad, (uint64_t)ktot * 2, (uint32_t)kPairBK,
                            128u, kPairBK);
    if (rc != Y2_OK) return rc;
    SlabParams &p = pl->slab;
    p.cblocks = d->cin / kPairBK;
    p.wp = wp;
    p.hp = hp;
    p.h = d->h;
    p.w = d->w;
    p.total_pos = (int)total;
    p.tiles_m = (int)((total + 255) / 256);
    p.tiles_n = d->npad / kPairN;
    p.halo = halo;
    p.slab_loads = loads;
    p.box_rows = box_rows;
    p.slab_bytes = slab_bytes;
    p.stages_a = stages_a;
    p.stages_b = stages_b;
    p.cout = d->cout;
    p.act = d->act;
    p.out_mode = d->out_mode;
    p.out_cs = d->out_cs;
    p.couple = 0;
    p.tma_store = 0;
    p.dbg = getenv("Y2_PAIR_DBG") ? atoi(getenv("Y2_PAIR_DBG")) : 0;
    memset(&pl->tm_out, 0, sizeof(pl->tm_out));
    if (d->out_mode == Y2_OUT_BF16_PADDED && d->cout % 64 == 0 && !getenv("Y2_SLAB_NO_TMA_STORE")) {
        rc = encode_2d_bf16(&pl->tm_out, d->out, (uint64_t)d->cout, (uint64_t)total, (uint64_t)d->out_cs * 2, 64u, 32u, 64);
        if (rc != Y2_OK) return rc;
        p.tma_store = 1;
    }
    p.alpha = d->alpha;
    p.beta = d->beta;
    p.out = d->out;
    pl->variant = kVariantPair;
    pl->block_n = kPairN;
    pl->block_k = kPairBK;
    pl->taps = taps;
    pl->smem_bytes = (size_t)stages_a * slab_bytes + (size_t)stages_b * kPairBHalfBytes + aux + 1024;
    const int tiles = p.tiles_m * p.tiles_n;
    const int pairs = tiles < sms / 2 ? tiles : sms / 2;
    pl->grid = 2 * pairs;
    // stream-K when whole tiles would leave the last wave badly filled (and there is a K loop to split)
    p.streamk = 0;
    p.sk_partial = nullptr;
    p.sk_flags = nullptr;
    // Opt-in (Y2_PAIR_STREAMK=1).  Measured on yolo-voc b64: it removes 5 % of the L23 kernel's cycles and gains
    // 1 % on the step during the first ~70 ms after idle, but under sustained load the chip sits at its 1 kW cap,
    // where only energy per step counts: the parked partial sums cost more than the idle SMs did and the
    // sustained step is 0.8 % SLOWER with it (1.912 / 1.932 vs 1.901 / 1.913 ms, tools/ab_sustained.sh).
    if (tiles > pairs && p.cblocks >= 2 && taps == 9) {  // short K loops: the parked partials would cost even more
        const char *e = getenv("Y2_PAIR_STREAMK");
        p.streamk = e && atoi(e) != 0;
    }
    if (p.streamk) {
        const size_t partial_bytes = (size_t)(pairs + 1) * kPairN * 256 * sizeof(float);
        const size_t flag_bytes = (size_t)(pairs + 1) * 16 * sizeof(int);
        Y2_CUDA_CHECK(cudaMalloc(&pl->sk_buf, partial_bytes + flag_bytes));
        Y2_CUDA_CHECK(cudaMemset(pl->sk_buf, 0, partial_bytes + flag_bytes));
        p.sk_partial = (float *)pl->sk_buf;
        p.sk_flags = (int *)((char *)pl->sk_buf + partial_bytes);
    }
    static bool attr_done[64] = {false};
    int dev = 0;
    Y2_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
        Y2_CUDA_CHECK(cudaFuncSetAttribute(conv_pair_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        Y2_CUDA_CHECK(cudaFuncSetAttribute(conv_pair_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done[dev] = true;
    }
    return Y2_OK;
}

int pair_plan_launch(const y2_conv_plan *pl, cudaStream_t st)
{
    if (pl->taps == 9)
        Y2_CUDA_CHECK(launch_pdl(conv_pair_kernel<9>, dim3(pl->grid), dim3(kPairThreads), pl->smem_bytes, st, pl->tm_a,
                                 pl->tm_b, pl->tm_out, pl->slab));
    else
        Y2_CUDA_CHECK(launch_pdl(conv_pair_kernel<1>, dim3(pl->grid), dim3(kPairThreads), pl->smem_bytes, st, pl->tm_a,
                                 pl->tm_b, pl->tm_out, pl->slab));
    Y2_LAUNCH_CHECK();
    return Y2_OK;
}

} // namespace y2
