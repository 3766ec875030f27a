// Device runtime behind the C-ABI: memory, streams, events, CUDA-graph capture.
// Replaces the reference's cuda.c:12-158 (cuda_set_device / cuda_make_array /
// cuda_push_array / cuda_pull_array / cuda_free) and the per-predict
// cudaMalloc+cudaFree of network_kernels.cu:392-407.
#include "y2_common.cuh"

#include <stdarg.h>
#include <string.h>

namespace y2 {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

} // namespace y2

using namespace y2;

extern "C" const char *y2_last_error(void) { return g_err; }
extern "C" const char *y2_version(void) { return "yolo2-b200 0.1 (sm_100a)"; }

extern "C" int y2_device_count(int *count)
{
    if (!count) return Y2_EINVAL;
    Y2_CUDA_CHECK(cudaGetDeviceCount(count));
    return Y2_OK;
}
extern "C" int y2_set_device(int dev)
{
    Y2_CUDA_CHECK(cudaSetDevice(dev));
    return Y2_OK;
}
extern "C" int y2_malloc(void **dptr, size_t bytes)
{
    if (!dptr) return Y2_EINVAL;
    Y2_CUDA_CHECK(cudaMalloc(dptr, bytes ? bytes : 16));
    return Y2_OK;
}
extern "C" int y2_free(void *dptr)
{
    if (dptr) Y2_CUDA_CHECK(cudaFree(dptr));
    return Y2_OK;
}
extern "C" int y2_memset(void *dptr, int value, size_t bytes, y2_stream_t s)
{
    Y2_CUDA_CHECK(cudaMemsetAsync(dptr, value, bytes, to_stream(s)));
    return Y2_OK;
}
extern "C" int y2_host_alloc(void **hptr, size_t bytes)
{
    if (!hptr) return Y2_EINVAL;
    Y2_CUDA_CHECK(cudaHostAlloc(hptr, bytes ? bytes : 16, cudaHostAllocDefault));
    return Y2_OK;
}
extern "C" int y2_host_free(void *hptr)
{
    if (hptr) Y2_CUDA_CHECK(cudaFreeHost(hptr));
    return Y2_OK;
}
extern "C" int y2_memcpy_h2d(void *dst, const void *src, size_t bytes, y2_stream_t s)
{
    Y2_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, to_stream(s)));
    return Y2_OK;
}
extern "C" int y2_memcpy_d2h(void *dst, const void *src, size_t bytes, y2_stream_t s)
{
    Y2_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, to_stream(s)));
    return Y2_OK;
}
extern "C" int y2_stream_create(y2_stream_t *s)
{
    if (!s) return Y2_EINVAL;
    cudaStream_t st;
    Y2_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    *s = (y2_stream_t)st;
    return Y2_OK;
}
extern "C" int y2_stream_destroy(y2_stream_t s)
{
    if (s) Y2_CUDA_CHECK(cudaStreamDestroy(to_stream(s)));
    return Y2_OK;
}
extern "C" int y2_stream_sync(y2_stream_t s)
{
    Y2_CUDA_CHECK(cudaStreamSynchronize(to_stream(s)));
    return Y2_OK;
}
extern "C" int y2_device_sync(void)
{
    Y2_CUDA_CHECK(cudaDeviceSynchronize());
    return Y2_OK;
}

extern "C" int y2_graph_begin(y2_stream_t s)
{
    Y2_CUDA_CHECK(cudaStreamBeginCapture(to_stream(s), cudaStreamCaptureModeThreadLocal));
    return Y2_OK;
}
extern "C" int y2_graph_end(y2_stream_t s, y2_graph_t *g)
{
    if (!g) return Y2_EINVAL;
    cudaGraph_t graph = nullptr;
    Y2_CUDA_CHECK(cudaStreamEndCapture(to_stream(s), &graph));
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        set_error("cudaGraphInstantiate -> %s", cudaGetErrorString(e));
        return Y2_ECUDA;
    }
    *g = (y2_graph_t)exec;
    return Y2_OK;
}
extern "C" int y2_graph_launch(y2_graph_t g, y2_stream_t s)
{
    Y2_CUDA_CHECK(cudaGraphLaunch((cudaGraphExec_t)g, to_stream(s)));
    return Y2_OK;
}
extern "C" int y2_graph_destroy(y2_graph_t g)
{
    if (g) Y2_CUDA_CHECK(cudaGraphExecDestroy((cudaGraphExec_t)g));
    return Y2_OK;
}

extern "C" int y2_event_create(y2_event_t *e)
{
    if (!e) return Y2_EINVAL;
    cudaEvent_t ev;
    Y2_CUDA_CHECK(cudaEventCreate(&ev));
    *e = (y2_event_t)ev;
    return Y2_OK;
}
extern "C" int y2_event_record(y2_event_t e, y2_stream_t s)
{
    Y2_CUDA_CHECK(cudaEventRecord((cudaEvent_t)e, to_stream(s)));
    return Y2_OK;
}
extern "C" int y2_event_elapsed_ms(y2_event_t a, y2_event_t b, float *ms)
{
    Y2_CUDA_CHECK(cudaEventSynchronize((cudaEvent_t)b));
    Y2_CUDA_CHECK(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
    return Y2_OK;
}
extern "C" int y2_event_sync(y2_event_t e)
{
    Y2_CUDA_CHECK(cudaEventSynchronize((cudaEvent_t)e));
    return Y2_OK;
}
extern "C" int y2_stream_wait_event(y2_stream_t s, y2_event_t e)
{
    Y2_CUDA_CHECK(cudaStreamWaitEvent(to_stream(s), (cudaEvent_t)e, 0));
    return Y2_OK;
}
extern "C" int y2_event_destroy(y2_event_t e)
{
    if (e) Y2_CUDA_CHECK(cudaEventDestroy((cudaEvent_t)e));
    return Y2_OK;
}

/* cudaGetErrorString for the reference's check_error(cudaError_t) (cuda.c:27-49) */
extern "C" const char *y2_cuda_error_string(int cuda_status)
{
    return cudaGetErrorString((cudaError_t)cuda_status);
}

/* cudaGetLastError, cleared: what check_error is usually fed (cuda.c: check_error(cudaPeekAtLastError())) */
extern "C" int y2_cuda_last_status(void)
{
    return (int)cudaGetLastError();
}
