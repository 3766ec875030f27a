/*
 * Data-parallel detection over the GPUs of one box from a single C caller: one `network`
 * replica per GPU, one host thread per GPU, images sharded contiguously, nothing but the per-image detection lists
 * comes back - they are written straight into the caller's arrays in image order, so there is no gather step and no
 * collective on the path.
 *
 * Reference idiom followed (behavioural spec only): network_kernels.cu:133-151 / 346-376 - train_networks starts one
 * pthread per `network` replica (train_network_in_thread), each thread selects its replica's device, and the caller
 * joins them.  The reference uses this for training only; its inference entry points are single-GPU, batch 1.
 */
#include "y2_host.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* n replicas of one cfg / weights file on the listed GPUs, each planned for `batch` images per call
 * (batch <= 0 keeps the cfg's own batch).  Replicas are built one after the other on the calling thread:
 * parse_network_cfg reads the process-global gpu_index (parser.c:591), which is restored afterwards. */
network *parse_network_cfg_multi(char *cfgfile, char *weightfile, int *gpus, int ngpus, int batch)
{
    if (ngpus <= 0 || !gpus) error("parse_network_cfg_multi: empty GPU list");
    network *nets = (network *)calloc(ngpus, sizeof(network));
    const int saved = gpu_index;
    for (int i = 0; i < ngpus; ++i) {
        cuda_set_device(gpus[i]);
        nets[i] = parse_network_cfg(cfgfile);
        if (weightfile && weightfile[0]) load_weights(&nets[i], weightfile);
        if (batch > 0 && batch != nets[i].batch) set_batch_network(&nets[i], batch);
        nets[i].gpu_index = gpus[i];
    }
    if (saved >= 0) cuda_set_device(saved);
    gpu_index = saved;
    return nets;
}

void free_network_multi(network *nets, int n)
{
    if (!nets) return;
    for (int i = 0; i < n; ++i) free_network(nets[i]);
    free(nets);
}

int network_multi_batch(network *nets, int n)
{
    int total = 0;
    for (int i = 0; i < n; ++i) total += nets[i].batch;
    return total;
}

enum { JOB_DETECT, JOB_DETECT_U8, JOB_SUBMIT, JOB_SUBMIT_U8, JOB_WAIT };

typedef struct {
    network net;
    int kind;
    const void *input;
    float thresh, nms;
    y2_detection *dets;
    int *counts;
    int max_det;
} multi_job;

static void *multi_worker(void *ptr)
{
    multi_job *j = (multi_job *)ptr;
    y2_net_rt *rt = y2_rt(j->net);
    if (!rt) error("network_*_multi: replica has no device plan");
    /* this thread serves one GPU: stay on the CPUs next to it (staging copies, NUMA) and select the device */
    y2_bind_thread_to_device(rt->device);
    Y2_CHECK(y2_set_device(rt->device));
    switch (j->kind) {
    case JOB_DETECT:
        network_detect_batch(j->net, (const float *)j->input, j->thresh, j->nms, j->dets, j->counts, j->max_det);
        break;
    case JOB_DETECT_U8:
        network_detect_batch_u8(j->net, (const unsigned char *)j->input, j->thresh, j->nms, j->dets, j->counts,
                                j->max_det);
        break;
    case JOB_SUBMIT:
        network_detect_submit(j->net, (const float *)j->input, j->thresh, j->nms, j->max_det);
        break;
    case JOB_SUBMIT_U8:
        network_detect_submit_u8(j->net, (const unsigned char *)j->input, j->thresh, j->nms, j->max_det);
        break;
    case JOB_WAIT:
        network_detect_wait(j->net, j->dets, j->counts, j->max_det);
        break;
    }
    return 0;
}

/* one thread per replica, replica i owning images [sum batch_0..i-1, + batch_i) of the global batch */
static void run_multi(network *nets, int n, int kind, const void *input, size_t bytes_per_image, float thresh, float nms,
                      y2_detection *dets, int *counts, int max_det)
{
    if (n <= 0 || !nets) error("network_*_multi: no replicas");
    multi_job *jobs = (multi_job *)calloc(n, sizeof(multi_job));
    pthread_t *threads = (pthread_t *)calloc(n, sizeof(pthread_t));
    size_t first = 0;
    for (int i = 0; i < n; ++i) {
        jobs[i].net = nets[i];
        jobs[i].kind = kind;
        jobs[i].input = input ? (const char *)input + first * bytes_per_image : 0;
        jobs[i].thresh = thresh;
        jobs[i].nms = nms;
        jobs[i].dets = dets ? dets + first * (size_t)max_det : 0;
        jobs[i].counts = counts ? counts + first : 0;
        jobs[i].max_det = max_det;
        first += (size_t)nets[i].batch;
        if (n == 1) multi_worker(&jobs[i]);
        else if (pthread_create(&threads[i], 0, multi_worker, &jobs[i])) error("network_*_multi: thread creation failed");
    }
    if (n > 1)
        for (int i = 0; i < n; ++i) pthread_join(threads[i], 0);
    free(threads);
    free(jobs);
}

/* images: fp32 planar [total][c][h][w], total = network_multi_batch(nets, n);  dets [total][max_det], counts [total] */
void network_detect_batch_multi(network *nets, int n, const float *images, float thresh, float nms, y2_detection *dets,
                                int *counts, int max_det)
{
    run_multi(nets, n, JOB_DETECT, images, (size_t)nets[0].inputs * sizeof(float), thresh, nms, dets, counts, max_det);
}

/* images: uint8 interleaved RGB [total][h][w][3] at the network's resolution */
void network_detect_batch_u8_multi(network *nets, int n, const unsigned char *images_hwc, float thresh, float nms,
                                   y2_detection *dets, int *counts, int max_det)
{
    run_multi(nets, n, JOB_DETECT_U8, images_hwc, (size_t)nets[0].h * nets[0].w * 3, thresh, nms, dets, counts, max_det);
}

/* The two-deep pipeline of network_detect_submit / network_detect_wait over all replicas: submit stages (one thread
 * per GPU copies its slice into that GPU's pinned buffer) and enqueues a global batch, wait hands out the oldest
 * one.  Up to two global batches may be in flight. */
void network_detect_submit_multi(network *nets, int n, const float *images, float thresh, float nms, int max_det)
{
    run_multi(nets, n, JOB_SUBMIT, images, (size_t)nets[0].inputs * sizeof(float), thresh, nms, 0, 0, max_det);
}

void network_detect_submit_u8_multi(network *nets, int n, const unsigned char *images_hwc, float thresh, float nms,
                                    int max_det)
{
    run_multi(nets, n, JOB_SUBMIT_U8, images_hwc, (size_t)nets[0].h * nets[0].w * 3, thresh, nms, 0, 0, max_det);
}

void network_detect_wait_multi(network *nets, int n, y2_detection *dets, int *counts, int max_det)
{
    run_multi(nets, n, JOB_WAIT, 0, 0, 0.f, 0.f, dets, counts, max_det);
}
