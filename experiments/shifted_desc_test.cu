// Experiment (not product code): does a K-major swizzled UMMA A-descriptor whose start address is
// shifted by whole rows (row pitch 128 B / 64 B) read the rows TMA wrote there?  i.e. is the
// swizzle XOR a function of the absolute shared-memory address.  D = A[shift .. shift+127] * I.
#include "y2_common.cuh"
#include <vector>
#include <cstdlib>
#include <cstring>
#include <cmath>

namespace y2 { void set_error(const char *, ...) {} int sm_count() { return 148; } }
using namespace y2;

template <int BK>
__global__ void __launch_bounds__(128, 1)
shift_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int shift,
             int base_off, float *out)
{
    extern __shared__ uint8_t raw[];
    uint8_t *smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    constexpr int ROWS = 160;
    uint8_t *sa = smem;
    uint8_t *sb = smem + 32768;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 49152);
    uint32_t *slot = reinterpret_cast<uint32_t *>(smem + 49152 + 64);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar[0], ROWS * BK * 2 + 64 * BK * 2);
        tma_load_2d(&tm_a, &bar[0], sa, 0, 0);
        tma_load_2d(&tm_b, &bar[0], sb, 0, 0);
        mbar_wait(&bar[0], 0, 1);
        tc_fence_after();
        constexpr uint32_t sbo = 8 * BK * 2;
        constexpr uint32_t layout = (BK == 64) ? 2u : 4u;
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        uint64_t adesc = make_kmajor_desc(smem_u32(sa) + shift * BK * 2, sbo, layout);
        adesc |= (uint64_t)(base_off & 7) << 49;
        const uint64_t bdesc = make_kmajor_desc(smem_u32(sb), sbo, layout);
        for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)(k != 0));
        umma_commit(&bar[1]);
        mbar_wait(&bar[1], 0, 2);
    }
    __syncthreads();
    tc_fence_after();
    const int row = threadIdx.x;
    for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[row * 64 + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7fffu + ((u >> 16) & 1u); return (uint16_t)(u >> 16); }
static float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

template <int BK>
int run()
{
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const int ROWS = 160;
    std::vector<uint16_t> A(ROWS * BK), B(64 * BK, 0);
    srand(7);
    for (auto &a : A) a = f2bf((rand() % 2001 - 1000) / 64.0f);
    for (int n = 0; n < 64 && n < BK; ++n) B[n * BK + n] = f2bf(1.0f);
    void *dA, *dB; float *dO;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dO, 128 * 64 * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap ta, tb;
    CUtensorMapSwizzle sw = BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    { cuuint64_t gd[2] = {(cuuint64_t)BK, (cuuint64_t)ROWS}; cuuint64_t gs[1] = {(cuuint64_t)BK * 2}; cuuint32_t bx[2] = {(cuuint32_t)BK, (cuuint32_t)ROWS}; cuuint32_t es[2] = {1, 1};
      CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode A failed %d\n", (int)r); return 1; } }
    { cuuint64_t gd[2] = {(cuuint64_t)BK, 64}; cuuint64_t gs[1] = {(cuuint64_t)BK * 2}; cuuint32_t bx[2] = {(cuuint32_t)BK, 64}; cuuint32_t es[2] = {1, 1};
      CUresult r = enc(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode B failed %d\n", (int)r); return 1; } }
    cudaFuncSetAttribute(shift_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    std::vector<float> O(128 * 64);
    const int ncols = BK < 64 ? BK : 64;
    for (int mode = 0; mode < 2; ++mode)
        for (int shift = 0; shift <= 19; ++shift) {
            const int bo = mode ? ((shift * BK * 2) >> 7) & 7 : 0;
            shift_kernel<BK><<<1, 128, 64 * 1024>>>(ta, tb, shift, bo, dO);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("BK=%d shift=%d launch error %s\n", BK, shift, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < ncols; ++n)
                    if (O[m * 64 + n] != bf2f(A[(m + shift) * BK + n])) ++bad;
            printf("BK=%d base_offset_mode=%d shift=%2d base_off=%d mismatches=%d\n", BK, mode, shift, bo, bad);
        }
    return 0;
}

int main()
{
    if (run<64>()) return 1;
    if (run<32>()) return 1;
    return 0;
}
