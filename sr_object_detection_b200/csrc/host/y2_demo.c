/*
 * The video pipeline of demo.c over the GPU forward pass, fed by callbacks instead of an OpenCV capture / window.
 *
 * Reference interface replaced (behavioural spec only): demo.c:57-69 fetch_in_thread, :71-107 detect_in_thread and
 * the loop of demo() (:118-230): frame n+1 is fetched (and converted to a planar float image) on one thread while
 * frame n is detected on another; a detection = network_predict, the mean of the last FRAMES = 3 network outputs
 * (mean_arrays over a ring that starts at zero), get_region_boxes on the mean, do_nms at .4.  The reference draws the
 * boxes into the frame (OpenCV); here they are handed to the caller's sink, frame by frame, in stream order.
 */
#include "y2_host.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define DEMO_FRAMES 3 /* FRAMES of demo.c:18 */

typedef struct {
    network net;
    float thresh;
    y2_frame_source source;
    void *source_ctx;
    y2_detection_sink sink;
    void *sink_ctx;
    unsigned char *rgb;              /* the frame being fetched, uint8 interleaved RGB at the network's size */
    image in_s, det_s;               /* fetched / being detected (planar floats) */
    int stream_open;
    float *predictions[DEMO_FRAMES]; /* ring of network outputs */
    float *avg;
    int demo_index;
    int frame_index;
    box *boxes;
    float **probs;
} demo_state;

/* demo.c:57-69: next frame -> planar float image, value = byte / 255. (image.c ipl_to_image) */
static void *fetch_in_thread(void *ptr)
{
    demo_state *d = (demo_state *)ptr;
    const int w = d->net.w, h = d->net.h;
    d->stream_open = d->source(d->source_ctx, d->rgb, w, h);
    if (!d->stream_open) return 0;
    d->in_s = make_image(w, h, 3);
    for (int k = 0; k < 3; ++k)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x)
                d->in_s.data[((size_t)k * h + y) * w + x] = d->rgb[((size_t)y * w + x) * 3 + k] / 255.;
    return 0;
}

/* demo.c:71-107 */
static void *detect_in_thread(void *ptr)
{
    demo_state *d = (demo_state *)ptr;
    const float nms = .4;
    y2_net_rt *rt = y2_rt(d->net);
    if (rt) Y2_CHECK(y2_set_device(rt->device)); /* a fresh thread has no device selected */
    layer l = d->net.layers[d->net.n - 1];
    float *prediction = network_predict(d->net, d->det_s.data);
    memcpy(d->predictions[d->demo_index], prediction, (size_t)l.outputs * sizeof(float));
    mean_arrays(d->predictions, DEMO_FRAMES, l.outputs, d->avg);
    l.output = d->avg;
    free_image(d->det_s);
    if (l.type != REGION) error("Last layer must produce detections\n");
    get_region_boxes(l, 1, 1, d->thresh, d->probs, d->boxes, 0, 0);
    if (nms > 0) do_nms(d->boxes, d->probs, l.w * l.h * l.n, l.classes, nms);
    d->demo_index = (d->demo_index + 1) % DEMO_FRAMES;
    if (d->sink) d->sink(d->sink_ctx, d->frame_index, d->boxes, d->probs, l.w * l.h * l.n, l.classes);
    ++d->frame_index;
    return 0;
}

/* Runs until the source reports the end of the stream; returns the number of frames detected. */
int demo_frames(char *cfgfile, char *weightfile, float thresh, y2_frame_source source, void *source_ctx,
                y2_detection_sink sink, void *sink_ctx)
{
    demo_state d;
    memset(&d, 0, sizeof(d));
    d.net = parse_network_cfg(cfgfile);
    if (weightfile) load_weights(&d.net, weightfile);
    set_batch_network(&d.net, 1);
    d.thresh = thresh;
    d.source = source;
    d.source_ctx = source_ctx;
    d.sink = sink;
    d.sink_ctx = sink_ctx;
    layer l = d.net.layers[d.net.n - 1];
    const int total = l.w * l.h * l.n;
    d.rgb = (unsigned char *)calloc((size_t)d.net.w * d.net.h * 3, 1);
    d.avg = (float *)calloc(l.outputs, sizeof(float));
    for (int j = 0; j < DEMO_FRAMES; ++j) d.predictions[j] = (float *)calloc(l.outputs, sizeof(float));
    d.boxes = (box *)calloc(total, sizeof(box));
    d.probs = (float **)calloc(total, sizeof(float *));
    for (int j = 0; j < total; ++j) d.probs[j] = (float *)calloc(l.classes, sizeof(float));

    /* demo.c:160-178: one frame ahead, then fetch(n + 1) and detect(n) side by side */
    fetch_in_thread(&d);
    while (d.stream_open) {
        d.det_s = d.in_s;
        pthread_t fetch_thread, detect_thread;
        if (pthread_create(&fetch_thread, 0, fetch_in_thread, &d)) error("Thread creation failed");
        if (pthread_create(&detect_thread, 0, detect_in_thread, &d)) error("Thread creation failed");
        pthread_join(fetch_thread, 0);
        pthread_join(detect_thread, 0);
    }
    const int frames = d.frame_index;
    for (int j = 0; j < total; ++j) free(d.probs[j]);
    free(d.probs);
    free(d.boxes);
    for (int j = 0; j < DEMO_FRAMES; ++j) free(d.predictions[j]);
    free(d.avg);
    free(d.rgb);
    free_network(d.net);
    return frames;
}
