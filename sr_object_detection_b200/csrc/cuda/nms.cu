// Per-class sorted non-maximum suppression on the GPU, bit-exact with do_nms_sort
// (reference box.c:249-277 with nms_comparator 239-247 and box_iou 67-97).
//
// What the reference does, per class k in ascending order: stable-sort (glibc qsort is a
// merge sort) ALL boxes by probs[.][k] descending, starting from the order the previous
// class left behind; then walk the sorted list, and every box whose prob is still non-zero
// zeroes the class-k prob of every later box with IoU > thresh.
//
// Two facts make this parallel over (image, class) without changing a single result:
//  * boxes with prob 0 never suppress and being suppressed changes nothing for them, so only
//    the non-zero "candidates" of a class matter;
//  * the carried-over order is a closed form: class k sees the boxes ordered by
//    (p[k-1] desc, p[k-2] desc, ..., p[0] desc, index asc) of the ORIGINAL probabilities
//    (each class is sorted before its own suppression runs), so ties in p[k] are broken by
//    walking back through the previous classes.
// Suppressed entries are first marked by flipping their sign (|p| keeps the original value
// for the tie-breaks of other classes running concurrently); a second pass writes the zeros.
//
// Compiled with -fmad=false: IoU must round exactly like the C code.
#include "y2_common.cuh"

namespace y2 {

__device__ __forceinline__ float overlap_ref(float x1, float w1, float x2, float w2)
{
    const float l1 = x1 - w1 / 2;
    const float l2 = x2 - w2 / 2;
    const float left = l1 > l2 ? l1 : l2;
    const float r1 = x1 + w1 / 2;
    const float r2 = x2 + w2 / 2;
    const float right = r1 < r2 ? r1 : r2;
    return right - left;
}

__device__ __forceinline__ float box_iou_ref(const float4 a, const float4 b)
{
    // float4 = (x, y, w, h)
    const float w = overlap_ref(a.x, a.z, b.x, b.z);
    const float h = overlap_ref(a.y, a.w, b.y, b.w);
    float inter;
    if (w < 0 || h < 0) inter = 0;
    else inter = w * h;
    const float uni = a.z * a.w + b.z * b.w - inter;
    return inter / uni;
}

// does candidate (ia, pa) sort before candidate (ib, pb) for class k?
__device__ __forceinline__ bool sorts_before(const float *__restrict__ probs_img, int classes, int k,
                                             int ia, float pa, int ib, float pb)
{
    if (pa != pb) return pa > pb;
    for (int c = k - 1; c >= 0; --c) {
        const float qa = fabsf(probs_img[(size_t)ia * classes + c]);
        const float qb = fabsf(probs_img[(size_t)ib * classes + c]);
        if (qa != qb) return qa > qb;
    }
    return ia < ib;
}

// count non-zero entries per (image, class); coalesced over the probs matrix
__global__ void nms_count_kernel(const float *__restrict__ probs, int *__restrict__ cnt, int batch, int total,
                                 int classes)
{
    const long long n = (long long)batch * total * classes;
    const long long per_img = (long long)total * classes;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n;
         t += (long long)gridDim.x * blockDim.x) {
        if (probs[t] != 0.f) {
            const int b = (int)(t / per_img);
            const int k = (int)(t % classes);
            atomicAdd(&cnt[(size_t)b * classes + k], 1);
        }
    }
}

// Persistent grid: every block walks the (image, class) counter array with a grid stride and works on the
// pairs that have at least two candidates (yolo9000: 9418 classes x batch pairs, a handful of them live).
// Dynamic smem: idx[cap] int, key[cap] float, sorted[cap] int, alive[cap] int, box[cap] float4.
__global__ void nms_mark_kernel(const float4 *__restrict__ boxes, float *probs, const int *__restrict__ cnt,
                                int batch, int total, int classes, float thresh, int cap)
{
    extern __shared__ __align__(16) unsigned char nms_smem[];
    const long long pairs = (long long)batch * classes;
  for (long long pair = blockIdx.x; pair < pairs; pair += gridDim.x) {
    if (cnt[pair] <= 1) continue; // a lone candidate only "suppresses" zeros (uniform per block)
    const int b = (int)(pair / classes), k = (int)(pair - (long long)b * classes);
    float4 *s_box = reinterpret_cast<float4 *>(nms_smem);
    int *s_idx = reinterpret_cast<int *>(s_box + cap);
    float *s_key = reinterpret_cast<float *>(s_idx + cap);
    int *s_sorted = reinterpret_cast<int *>(s_key + cap);
    int *s_alive = s_sorted + cap;
    __shared__ int s_warp_tot[32];
    __shared__ int s_base;

    float *pimg = probs + (size_t)b * total * classes;
    const float4 *bimg = boxes + (size_t)b * total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;

    // 1. ordered compaction of the candidates (index order)
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int start = 0; start < total; start += blockDim.x) {
        const int i = start + threadIdx.x;
        float p = 0.f;
        if (i < total) p = pimg[(size_t)i * classes + k];
        const bool is = (p != 0.f);
        const unsigned m = __ballot_sync(0xffffffffu, is);
        const int within = __popc(m & ((1u << lane) - 1));
        if (lane == 0) s_warp_tot[warp] = __popc(m);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += s_warp_tot[w];
        if (is) {
            const int pos = off + within;
            if (pos < cap) {
                s_idx[pos] = i;
                s_key[pos] = fabsf(p);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < nw; ++w) t += s_warp_tot[w];
            s_base += t;
        }
        __syncthreads();
    }
    const int m = s_base < cap ? s_base : cap;

    // 2. rank sort: position = number of candidates that sort before me
    for (int a = threadIdx.x; a < m; a += blockDim.x) {
        const int ia = s_idx[a];
        const float pa = s_key[a];
        int rank = 0;
        for (int o = 0; o < m; ++o) {
            if (o == a) continue;
            if (sorts_before(pimg, classes, k, s_idx[o], s_key[o], ia, pa)) ++rank;
        }
        s_sorted[rank] = ia;
    }
    __syncthreads();
    for (int a = threadIdx.x; a < m; a += blockDim.x) {
        s_box[a] = bimg[s_sorted[a]];
        s_alive[a] = 1;
    }
    __syncthreads();

    // 3. greedy suppression in sorted order; a barrier only after a live box acted
    for (int i = 0; i < m - 1; ++i) {
        if (!s_alive[i]) continue; // uniform: everyone reads the same flag
        const float4 a = s_box[i];
        for (int j = i + 1 + threadIdx.x; j < m; j += blockDim.x) {
            if (s_alive[j] && box_iou_ref(a, s_box[j]) > thresh) s_alive[j] = 0;
        }
        __syncthreads();
    }

    // 4. mark the suppressed entries (sign flip keeps |p| for concurrent tie-breaks)
    for (int a = threadIdx.x; a < m; a += blockDim.x) {
        if (!s_alive[a]) {
            float *q = pimg + (size_t)s_sorted[a] * classes + k;
            *q = -fabsf(*q);
        }
    }
    __syncthreads(); // the shared arrays are reused by this block's next pair
  }
}

// negative == suppressed -> 0
__global__ void nms_clear_kernel(float *__restrict__ probs, long long n)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n;
         t += (long long)gridDim.x * blockDim.x) {
        if (probs[t] < 0.f) probs[t] = 0.f;
    }
}

// ---- do_nms (box.c:279-297), the unsorted variant --------------------------------------------------
// bit j of mask[i][j / 32] : j > i and box_iou(boxes[i], boxes[j]) > thresh
__global__ void iou_mask_kernel(const float4 *__restrict__ boxes, int total, float thresh, unsigned *__restrict__ mask,
                                int words)
{
    const long long n = (long long)total * words;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / words), w = (int)(t - (long long)i * words);
        const float4 a = boxes[i];
        unsigned bits = 0;
        for (int bit = 0; bit < 32; ++bit) {
            const int j = w * 32 + bit;
            if (j > i && j < total && box_iou_ref(a, boxes[j]) > thresh) bits |= 1u << bit;
        }
        mask[t] = bits;
    }
}

// One block.  The reference walks rows i in order, skips a row none of whose probabilities is positive AT THAT
// MOMENT (a test that couples all classes), then for every later overlapping row j zeroes, per class, the
// smaller of the two entries.  Here a thread owns whole class columns (no two threads ever touch the same
// entry); only the row test needs the block.
__global__ void __launch_bounds__(1024) nms_unsorted_kernel(float *probs, const unsigned *__restrict__ mask, int total,
                                                            int classes, int words)
{
    for (int i = 0; i < total; ++i) {
        float *pi = probs + (size_t)i * classes;
        int mine = 0;
        for (int k = threadIdx.x; k < classes; k += blockDim.x) mine |= (pi[k] > 0);
        if (!__syncthreads_or(mine)) continue;
        for (int w = (i + 1) >> 5; w < words; ++w) {
            unsigned bits = mask[(size_t)i * words + w];
            while (bits) {
                const int j = (w << 5) + __ffs(bits) - 1;
                bits &= bits - 1;
                float *pj = probs + (size_t)j * classes;
                for (int k = threadIdx.x; k < classes; k += blockDim.x) {
                    if (pi[k] < pj[k]) pi[k] = 0;
                    else pj[k] = 0;
                }
            }
        }
    }
}

} // namespace y2

using namespace y2;

extern "C" int y2_nms_unsorted(const float *boxes, float *probs, int total, int classes, float thresh, y2_stream_t s)
{
    if (!boxes || !probs || total <= 0 || classes <= 0) return Y2_EINVAL;
    cudaStream_t st = to_stream(s);
    const int words = (total + 31) / 32;
    unsigned *mask = nullptr;
    Y2_CUDA_CHECK(cudaMalloc(&mask, (size_t)total * words * sizeof(unsigned)));
    const long long n = (long long)total * words;
    long long blocks = (n + 255) / 256;
    const long long capb = (long long)sm_count() * 16;
    if (blocks > capb) blocks = capb;
    iou_mask_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const float4 *>(boxes), total, thresh, mask, words);
    int rc = cudaGetLastError() == cudaSuccess ? Y2_OK : Y2_ECUDA;
    if (rc == Y2_OK) {
        int threads = classes < 1024 ? (classes + 31) / 32 * 32 : 1024;
        nms_unsorted_kernel<<<1, threads, 0, st>>>(probs, mask, total, classes, words);
        if (cudaGetLastError() != cudaSuccess) rc = Y2_ECUDA;
    }
    cudaStreamSynchronize(st);
    cudaFree(mask);
    return rc;
}

static int nms_mark_launch(const float *boxes, float *probs, const int *cnt, int batch, int total, int classes,
                           float thresh, cudaStream_t st)
{
    const int cap = total;
    const size_t smem = (size_t)cap * (sizeof(float4) + 4 * sizeof(int));
    if (smem > 200 * 1024) {
        set_error("y2_nms_sort: %d boxes per image exceed the shared-memory staging", total);
        return Y2_EINVAL;
    }
    int dev = 0;
    Y2_CUDA_CHECK(cudaGetDevice(&dev));
    static bool attr_done[64] = {false};
    if (dev >= 0 && dev < 64 && !attr_done[dev] && smem > 48 * 1024) {
        Y2_CUDA_CHECK(cudaFuncSetAttribute(nms_mark_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           200 * 1024));
        attr_done[dev] = true;
    }
    const long long pairs = (long long)batch * classes;
    const int per_sm = smem > 100 * 1024 ? 1 : smem > 48 * 1024 ? 2 : 4;
    long long blocks = (long long)sm_count() * per_sm;
    if (blocks > pairs) blocks = pairs;
    nms_mark_kernel<<<(int)blocks, 256, smem, st>>>(reinterpret_cast<const float4 *>(boxes), probs, cnt, batch, total,
                                                    classes, thresh, cap);
    Y2_LAUNCH_CHECK();
    return Y2_OK;
}

// Suppression only, for callers that own the candidate counters (the network's detection path: the box
// decode counts while it writes the probabilities, the final pick reads "negative" as suppressed).
extern "C" int y2_nms_mark(const float *boxes, float *probs, const int *cnt, int batch, int total, int classes,
                           float thresh, y2_stream_t s)
{
    if (!boxes || !probs || !cnt || batch <= 0 || total <= 0 || classes <= 0) return Y2_EINVAL;
    return nms_mark_launch(boxes, probs, cnt, batch, total, classes, thresh, to_stream(s));
}

// Stand-alone do_nms_sort: owns nothing between calls (the counters live for the duration of the call), so any
// number of networks, streams and host threads may use it concurrently.
extern "C" int y2_nms_sort(const float *boxes, float *probs, int batch, int total, int classes, float thresh,
                           y2_stream_t s)
{
    if (!boxes || !probs || batch <= 0 || total <= 0 || classes <= 0) return Y2_EINVAL;
    cudaStream_t st = to_stream(s);
    const size_t need = (size_t)batch * classes;
    int *cnt = nullptr;
    Y2_CUDA_CHECK(cudaMalloc(&cnt, need * sizeof(int)));
    cudaError_t e = cudaMemsetAsync(cnt, 0, need * sizeof(int), st);
    int rc = Y2_OK;
    if (e != cudaSuccess) rc = Y2_ECUDA;
    const long long n = (long long)batch * total * classes;
    long long blocks = (n + 255) / 256;
    const long long capb = (long long)sm_count() * 16;
    if (blocks > capb) blocks = capb;
    if (rc == Y2_OK) {
        nms_count_kernel<<<(int)blocks, 256, 0, st>>>(probs, cnt, batch, total, classes);
        if (cudaGetLastError() != cudaSuccess) rc = Y2_ECUDA;
    }
    if (rc == Y2_OK) rc = nms_mark_launch(boxes, probs, cnt, batch, total, classes, thresh, st);
    if (rc == Y2_OK) {
        nms_clear_kernel<<<(int)blocks, 256, 0, st>>>(probs, n);
        if (cudaGetLastError() != cudaSuccess) rc = Y2_ECUDA;
    }
    cudaStreamSynchronize(st); // the counters are freed below
    cudaFree(cnt);
    return rc;
}
