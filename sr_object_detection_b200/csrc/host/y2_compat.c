/* Remaining symbols of the reference's C surface that callers outside the forward pass link against:
 * host and device BLAS-ish vector helpers (blas.h:16-19, 43-53), activate_array[_ongpu]
 * (activations.h:16-18), check_error / cuda_gridsize (cuda.h:22,32), the per-layer resize_*_layer
 * functions (convolutional_layer.h:29, maxpool_layer.h:13, reorg_layer.h:10, route_layer.h:11,
 * region_layer.h:13, avgpool_layer.h:13).
 *
 * gemm_ongpu / gemm_gpu (gemm.c:173-213) and im2col_ongpu (im2col_kernels.cu:48-61) are provided as plain fp32 device
 * kernels for callers of the helper surface; the network's own convolutions never use them (implicit GEMM inside the
 * tcgen05 kernels, no im2col buffer).
 *
 * Not provided on purpose: the CPU forward_*_layer functions, gemm / gemm_cpu and im2col_cpu (there is no CPU
 * execution path; a caller that links them fails at link time rather than silently computing on the host) and
 * blas_handle (cuda.h:23): there is no cuBLAS handle in this library. */
#include "y2_host.h"

#include <assert.h>
#include <math.h>

/* ---- host vector helpers (blas.c:150-190) --------------------------------------------------------- */
void fill_cpu(int N, float ALPHA, float *X, int INCX)
{
    for (int i = 0; i < N; ++i) X[(size_t)i * INCX] = ALPHA;
}

void copy_cpu(int N, float *X, int INCX, float *Y, int INCY)
{
    for (int i = 0; i < N; ++i) Y[(size_t)i * INCY] = X[(size_t)i * INCX];
}

void axpy_cpu(int N, float ALPHA, float *X, int INCX, float *Y, int INCY)
{
    for (int i = 0; i < N; ++i) Y[(size_t)i * INCY] += ALPHA * X[(size_t)i * INCX];
}

void scal_cpu(int N, float ALPHA, float *X, int INCX)
{
    for (int i = 0; i < N; ++i) X[(size_t)i * INCX] *= ALPHA;
}

/* ---- device vector helpers on cuda_make_array buffers (blas_kernels.cu:402-470, 560-616) ----------- */
void fill_ongpu(int N, float ALPHA, float *X, int INCX)
{
    Y2_CHECK(y2_vec_fill(N, ALPHA, X, INCX, 0));
}

void copy_ongpu_offset(int N, float *X, int OFFX, int INCX, float *Y, int OFFY, int INCY)
{
    Y2_CHECK(y2_vec_copy(N, X + OFFX, INCX, Y + OFFY, INCY, 0));
}

void copy_ongpu(int N, float *X, int INCX, float *Y, int INCY)
{
    copy_ongpu_offset(N, X, 0, INCX, Y, 0, INCY);
}

void axpy_ongpu_offset(int N, float ALPHA, float *X, int OFFX, int INCX, float *Y, int OFFY, int INCY)
{
    Y2_CHECK(y2_vec_axpy(N, ALPHA, X + OFFX, INCX, Y + OFFY, INCY, 0));
}

void axpy_ongpu(int N, float ALPHA, float *X, int INCX, float *Y, int INCY)
{
    axpy_ongpu_offset(N, ALPHA, X, 0, INCX, Y, 0, INCY);
}

void scal_ongpu(int N, float ALPHA, float *X, int INCX)
{
    Y2_CHECK(y2_vec_scal(N, ALPHA, X, INCX, 0));
}

/* ---- activations (activations.h:22-60, activations.c:64-101) --------------------------------------
 * float argument, double math where the reference's expressions promote (exp, the .1 / .01 constants),
 * one rounding back to float at the return. */
float activate(float x, ACTIVATION a)
{
    switch (a) {
    case LINEAR: return x;
    case LOGISTIC: return (float)(1. / (1. + exp(-x)));
    case LOGGY: return (float)(2. / (1. + exp(-x)) - 1);
    case RELU: return x * (x > 0);
    case ELU: return (float)((x >= 0) * x + (x < 0) * (exp(x) - 1));
    case RELIE: return (x > 0) ? x : (float)(.01 * x);
    case RAMP: return (float)(x * (x > 0) + .1 * x);
    case LEAKY: return (x > 0) ? x : (float)(.1 * x);
    case TANH: return (float)((exp(2 * x) - 1) / (exp(2 * x) + 1));
    case PLSE:
        if (x < -4) return (float)(.01 * (x + 4));
        if (x > 4) return (float)(.01 * (x - 4) + 1);
        return (float)(.125 * x + .5);
    case STAIR: {
        int n = (int)floor(x);
        if (n % 2 == 0) return (float)floor(x / 2.);
        return (float)((x - n) + floor(x / 2.));
    }
    case HARDTAN: return x < -1 ? -1 : x > 1 ? 1 : x;
    case LHTAN:
        if (x < 0) return (float)(.001 * x);
        if (x > 1) return (float)(.001 * (x - 1) + 1);
        return x;
    }
    return 0;
}

void activate_array(float *x, const int n, const ACTIVATION a)
{
    for (int i = 0; i < n; ++i) x[i] = activate(x[i], a);
}

void activate_array_ongpu(float *x, int n, ACTIVATION a)
{
    Y2_CHECK(y2_vec_activate(x, n, (int)a, 0));
}

/* ---- cuda.c:27-62 ---------------------------------------------------------------------------------- */
void check_error(int status)
{
    int status2 = y2_cuda_last_status();
    if (status != 0) {
        char buffer[256];
        const char *s = y2_cuda_error_string(status);
        fprintf(stderr, "CUDA Error: %s\n", s);
        snprintf(buffer, sizeof(buffer), "CUDA Error: %s", s);
        assert(0);
        error(buffer);
    }
    if (status2 != 0) {
        char buffer[256];
        const char *s = y2_cuda_error_string(status2);
        fprintf(stderr, "CUDA Error Prev: %s\n", s);
        snprintf(buffer, sizeof(buffer), "CUDA Error Prev: %s", s);
        assert(0);
        error(buffer);
    }
}

/* launch extent for n work items in blocks of BLOCK threads; folds into y above 65535 blocks */
y2_dim3 cuda_gridsize(size_t n)
{
    size_t k = (n - 1) / BLOCK + 1;
    size_t x = k, y = 1;
    if (x > 65535) {
        x = (size_t)ceil(sqrt((double)k));
        y = (n - 1) / (x * BLOCK) + 1;
    }
    y2_dim3 d = {(unsigned)x, (unsigned)y, 1};
    return d;
}

/* ---- per-layer resize: host-side extents only; resize_network re-plans the device side -------------- */
void resize_convolutional_layer(convolutional_layer *l, int w, int h)
{
    l->w = w;
    l->h = h;
    l->out_w = (w + 2 * l->pad - l->size) / l->stride + 1; /* convolutional_layer.c:75-83 */
    l->out_h = (h + 2 * l->pad - l->size) / l->stride + 1;
    l->outputs = l->out_h * l->out_w * l->out_c;
    l->inputs = l->w * l->h * l->c;
}

void resize_maxpool_layer(maxpool_layer *l, int w, int h)
{
    l->w = w;
    l->h = h;
    l->inputs = h * w * l->c;
    l->out_w = (w + 2 * l->pad) / l->stride; /* maxpool_layer.c:30-31, 62-63 */
    l->out_h = (h + 2 * l->pad) / l->stride;
    l->outputs = l->out_w * l->out_h * l->c;
}

void resize_region_layer(layer *l, int w, int h)
{
    l->w = w;
    l->h = h;
    l->outputs = h * w * l->n * (l->classes + l->coords + 1);
    l->inputs = l->outputs;
}

void resize_route_layer(route_layer *l, network *net)
{
    layer first = net->layers[l->input_layers[0]];
    l->out_w = first.out_w;
    l->out_h = first.out_h;
    l->out_c = first.out_c;
    l->outputs = first.outputs;
    l->input_sizes[0] = first.outputs;
    for (int k = 1; k < l->n; ++k) {
        layer next = net->layers[l->input_layers[k]];
        l->outputs += next.outputs;
        l->input_sizes[k] = next.outputs;
        if (next.out_w == first.out_w && next.out_h == first.out_h) l->out_c += next.out_c;
        else l->out_h = l->out_w = l->out_c = 0; /* route_layer.c:58-63 */
    }
    l->inputs = l->outputs;
}

void resize_reorg_layer(layer *l, int w, int h)
{
    l->w = w;
    l->h = h;
    if (l->reverse) {
        l->out_w = w * l->stride;
        l->out_h = h * l->stride;
    } else {
        l->out_w = w / l->stride;
        l->out_h = h / l->stride;
    }
    l->outputs = l->out_h * l->out_w * l->out_c;
    l->inputs = l->outputs;
}

void resize_avgpool_layer(avgpool_layer *l, int w, int h)
{
    l->w = w;
    l->h = h;
    l->inputs = h * w * l->c;
}

/* ---- gemm.c:173-213, im2col_kernels.cu:48-61 on cuda_make_array buffers ------------------------------------------- */
void gemm_ongpu(int TA, int TB, int M, int N, int K, float ALPHA, float *A_gpu, int lda, float *B_gpu, int ldb,
                float BETA, float *C_gpu, int ldc)
{
    Y2_CHECK(y2_sgemm(TA, TB, M, N, K, ALPHA, A_gpu, lda, B_gpu, ldb, BETA, C_gpu, ldc, 0));
}

/* host arrays in, host array out (gemm.c:185-213) */
void gemm_gpu(int TA, int TB, int M, int N, int K, float ALPHA, float *A, int lda, float *B, int ldb, float BETA,
              float *C, int ldc)
{
    float *A_gpu = cuda_make_array(A, (size_t)(TA ? lda * K : lda * M));
    float *B_gpu = cuda_make_array(B, (size_t)(TB ? ldb * N : ldb * K));
    float *C_gpu = cuda_make_array(C, (size_t)ldc * M);
    gemm_ongpu(TA, TB, M, N, K, ALPHA, A_gpu, lda, B_gpu, ldb, BETA, C_gpu, ldc);
    cuda_pull_array(C_gpu, C, (size_t)ldc * M);
    cuda_free(A_gpu);
    cuda_free(B_gpu);
    cuda_free(C_gpu);
}

void im2col_ongpu(float *im, int channels, int height, int width, int ksize, int stride, int pad, float *data_col)
{
    Y2_CHECK(y2_im2col_f32(im, channels, height, width, ksize, stride, pad, data_col, 0));
}
