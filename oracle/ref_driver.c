/*
 * TEST INFRASTRUCTURE — not part of the product.
 *
 * Driver that runs the UNMODIFIED reference CPU implementation (compiled from
 * /root/reference/src_yolo2 by oracle/Makefile into oracle/_ref/darknet_ref) and dumps
 * everything the parity tests compare: per-layer activations, the region-layer output,
 * boxes/probs after get_region_boxes and after do_nms_sort.  It is also the CPU baseline
 * of bench.py (`--impl reference` and the `cpu_baseline` leg).
 *
 * Call sequence mirrors detector.c:454-512 (test_detector) with gpu_index = -1.
 *
 *   darknet_ref forward <cfg> <weights> <input.f32> <outdir> <thresh> <nms> <dump_layers>
 *   darknet_ref region  <cfg> <region_in.f32> <outdir> <thresh> <nms>
 *   darknet_ref time    <cfg> <weights> <input.f32> <thresh> <nms> <warmup> <iters>
 *   darknet_ref resize  <in.f32> <c> <h> <w> <out_h> <out_w> <out.f32>
 *   darknet_ref layers  <cfg>                     (layer table as JSON, parser parity)
 *   darknet_ref donms   <boxes.f32> <probs.f32> <total> <classes> <thresh> <out.f32>   (do_nms, box.c:279-297)
 *
 * `forward` and `region` also write dets.f32: the final pick of Detector::detect
 * (yolo_v2_class.cpp:221-227: max_index over the post-NMS probabilities, keep prob > thresh) as rows of
 * 8 floats [image, box_index, obj_id, prob, x, y, w, h], in image / box order.
 *
 * Environment: Y2_DUMP_MAX_MB=N skips the per-layer dump of layers larger than N MB (big batches);
 * Y2_USE_MAP=1 passes the region layer's `map` to get_region_boxes (the 200-class
 * branch of region_layer.c:352-356, as validate_detector does, detector.c:344-347).
 *
 * <cfg> must carry batch=B subdivisions=1 (set_batch_network does not reallocate,
 * network.c:308-320); <input.f32> is raw float32 [B][C][H][W].
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "network.h"
#include "parser.h"
#include "region_layer.h"
#include "box.h"
#include "utils.h"
#include "image.h"

extern int gpu_index;
detectBoxes *GlobleObjBoxes; /* normally defined in darknet.c:358-359 */
int GlobleObjBoxesNum = 0;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static float *read_f32(const char *path, size_t n)
{
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    float *p = calloc(n, sizeof(float));
    size_t got = fread(p, sizeof(float), n, f);
    fclose(f);
    if (got != n) { fprintf(stderr, "%s: expected %zu floats, got %zu\n", path, n, got); exit(2); }
    return p;
}

static void write_f32(const char *dir, const char *name, const float *p, size_t n)
{
    char path[4096];
    snprintf(path, sizeof(path), "%s/%s", dir, name);
    FILE *f = fopen(path, "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", path); exit(2); }
    fwrite(p, sizeof(float), n, f);
    fclose(f);
}

/* get_region_boxes reads batch element 0 only (region_layer.c:331): apply it per image */
static void decode_and_nms(network net, const char *outdir, float thresh, float nms, int write)
{
    layer l = net.layers[net.n - 1];
    if (l.type != REGION) return;
    int total = l.w * l.h * l.n, b, j;
    int *map = (getenv("Y2_USE_MAP") && atoi(getenv("Y2_USE_MAP"))) ? l.map : 0;
    box *boxes = calloc(total, sizeof(box));
    float **probs = calloc(total, sizeof(float *));
    for (j = 0; j < total; ++j) probs[j] = calloc(l.classes, sizeof(float));
    float *all_boxes = calloc((size_t)l.batch * total * 4, sizeof(float));
    float *pre = calloc((size_t)l.batch * total * l.classes, sizeof(float));
    float *post = calloc((size_t)l.batch * total * l.classes, sizeof(float));
    float *dets = calloc((size_t)l.batch * total * 8, sizeof(float));
    size_t n_dets = 0;
    int out_classes = map ? 200 : l.classes;
    for (b = 0; b < l.batch; ++b) {
        layer lb = l;
        lb.output = l.output + (size_t)b * l.outputs;
        get_region_boxes(lb, 1, 1, thresh, probs, boxes, 0, map);
        memcpy(all_boxes + (size_t)b * total * 4, boxes, total * sizeof(box));
        for (j = 0; j < total; ++j)
            memcpy(pre + ((size_t)b * total + j) * l.classes, probs[j], l.classes * sizeof(float));
        if (nms > 0) do_nms_sort(boxes, probs, total, l.classes, nms);
        for (j = 0; j < total; ++j)
            memcpy(post + ((size_t)b * total + j) * l.classes, probs[j], l.classes * sizeof(float));
        for (j = 0; j < total; ++j) { /* yolo_v2_class.cpp:221-227 */
            int obj_id = max_index(probs[j], out_classes);
            float prob = probs[j][obj_id];
            if (prob > thresh) {
                float *d = dets + 8 * n_dets++;
                d[0] = (float)b; d[1] = (float)j; d[2] = (float)obj_id; d[3] = prob;
                d[4] = boxes[j].x; d[5] = boxes[j].y; d[6] = boxes[j].w; d[7] = boxes[j].h;
            }
        }
    }
    if (write) {
        write_f32(outdir, "boxes.f32", all_boxes, (size_t)l.batch * total * 4);
        write_f32(outdir, "probs_pre.f32", pre, (size_t)l.batch * total * l.classes);
        write_f32(outdir, "probs_post.f32", post, (size_t)l.batch * total * l.classes);
        /* region output after get_region_boxes (mutated in the tree case) */
        write_f32(outdir, "region_after_boxes.f32", l.output, (size_t)l.batch * l.outputs);
        write_f32(outdir, "dets.f32", dets, n_dets * 8);
    }
    free(dets);
    for (j = 0; j < total; ++j) free(probs[j]);
    free(probs); free(boxes); free(all_boxes); free(pre); free(post);
}

static int cmd_forward(int argc, char **argv)
{
    if (argc < 9) return 1;
    char *cfg = argv[2], *weights = argv[3], *input = argv[4], *outdir = argv[5];
    float thresh = atof(argv[6]), nms = atof(argv[7]);
    int dump = atoi(argv[8]);
    network net = parse_network_cfg(cfg);
    if (strcmp(weights, "-") != 0) load_weights(&net, weights);
    size_t n_in = (size_t)net.batch * net.inputs;
    float *X = read_f32(input, n_in);
    double t0 = now_s();
    float *out = network_predict(net, X);
    double t1 = now_s();
    int i;
    layer last = net.layers[net.n - 1];
    size_t n_out = (size_t)net.batch * get_network_output_size(net);
    write_f32(outdir, "output.f32", out, n_out);
    if (dump) {
        for (i = 0; i < net.n; ++i) {
            layer l = net.layers[i];
            if (!l.output || l.type == COST) continue;
            if (getenv("Y2_DUMP_MAX_MB") &&
                (double)l.batch * l.outputs * 4 > 1048576.0 * atof(getenv("Y2_DUMP_MAX_MB"))) continue;
            char name[64];
            snprintf(name, sizeof(name), "layer_%03d.f32", i);
            write_f32(outdir, name, l.output, (size_t)l.batch * l.outputs);
        }
    }
    decode_and_nms(net, outdir, thresh, nms, 1);
    printf("{\"batch\": %d, \"n_layers\": %d, \"outputs\": %d, \"predict_s\": %.6f, \"last_type\": %d}\n",
           net.batch, net.n, get_network_output_size(net), t1 - t0, (int)last.type);
    return 0;
}

static int cmd_region(int argc, char **argv)
{
    if (argc < 7) return 1;
    char *cfg = argv[2], *input = argv[3], *outdir = argv[4];
    float thresh = atof(argv[5]), nms = atof(argv[6]);
    network net = parse_network_cfg(cfg);
    layer l = net.layers[net.n - 1];
    if (l.type != REGION) { fprintf(stderr, "last layer is not a region layer\n"); return 2; }
    float *X = read_f32(input, (size_t)l.batch * l.inputs);
    network_state state = {0};
    state.net = net;
    state.input = X; /* NCHW conv output, as the previous layer would hand it over */
    state.train = 0;
    l.forward(l, state);
    write_f32(outdir, "region_out.f32", l.output, (size_t)l.batch * l.outputs);
    decode_and_nms(net, outdir, thresh, nms, 1);
    printf("{\"batch\": %d, \"boxes\": %d, \"classes\": %d}\n", l.batch, l.w * l.h * l.n, l.classes);
    return 0;
}

static int cmd_time(int argc, char **argv)
{
    if (argc < 9) return 1;
    char *cfg = argv[2], *weights = argv[3], *input = argv[4];
    float thresh = atof(argv[5]), nms = atof(argv[6]);
    int warmup = atoi(argv[7]), iters = atoi(argv[8]), i;
    network net = parse_network_cfg(cfg);
    if (strcmp(weights, "-") != 0) load_weights(&net, weights);
    float *X = read_f32(input, (size_t)net.batch * net.inputs);
    for (i = 0; i < warmup; ++i) { network_predict(net, X); decode_and_nms(net, 0, thresh, nms, 0); }
    double t0 = now_s();
    for (i = 0; i < iters; ++i) { network_predict(net, X); decode_and_nms(net, 0, thresh, nms, 0); }
    double t1 = now_s();
    printf("{\"batch\": %d, \"iters\": %d, \"seconds\": %.6f, \"images_per_s\": %.6f}\n", net.batch, iters,
           t1 - t0, (double)net.batch * iters / (t1 - t0));
    return 0;
}

static int cmd_resize(int argc, char **argv)
{
    if (argc < 9) return 1;
    int c = atoi(argv[3]), h = atoi(argv[4]), w = atoi(argv[5]), oh = atoi(argv[6]), ow = atoi(argv[7]);
    image im;
    im.c = c; im.h = h; im.w = w;
    im.data = read_f32(argv[2], (size_t)c * h * w);
    image r = resize_image(im, ow, oh);
    FILE *f = fopen(argv[8], "wb");
    fwrite(r.data, sizeof(float), (size_t)c * oh * ow, f);
    fclose(f);
    return 0;
}

/* letterbox <in.f32> <c> <h> <w> <out_h> <out_w> <out.f32> : the reference's letterbox_image */
static int cmd_letterbox(int argc, char **argv)
{
    if (argc < 9) return 1;
    int c = atoi(argv[3]), h = atoi(argv[4]), w = atoi(argv[5]), oh = atoi(argv[6]), ow = atoi(argv[7]);
    image im;
    im.c = c; im.h = h; im.w = w;
    im.data = read_f32(argv[2], (size_t)c * h * w);
    image r = letterbox_image(im, ow, oh);
    FILE *f = fopen(argv[8], "wb");
    fwrite(r.data, sizeof(float), (size_t)c * oh * ow, f);
    fclose(f);
    return 0;
}

/* topk <in.f32> <n> <k> : the reference's top_k, indices printed as a JSON list */
static int cmd_topk(int argc, char **argv)
{
    if (argc < 5) return 1;
    int n = atoi(argv[3]), k = atoi(argv[4]), j;
    float *a = read_f32(argv[2], (size_t)n);
    int *index = (int *)calloc(k, sizeof(int));
    top_k(a, n, k, index);
    printf("[");
    for (j = 0; j < k; ++j) printf("%s%d", j ? ", " : "", index[j]);
    printf("]\n");
    return 0;
}

/* the reference's unsorted do_nms (box.c:279-297) on given boxes [total][4] / probs [total][classes] */
static int cmd_donms(int argc, char **argv)
{
    if (argc < 8) return 1;
    int total = atoi(argv[4]), classes = atoi(argv[5]), j;
    float thresh = atof(argv[6]);
    box *boxes = (box *)read_f32(argv[2], (size_t)total * 4);
    float *flat = read_f32(argv[3], (size_t)total * classes);
    float **probs = calloc(total, sizeof(float *));
    for (j = 0; j < total; ++j) probs[j] = flat + (size_t)j * classes;
    do_nms(boxes, probs, total, classes, thresh);
    FILE *f = fopen(argv[7], "wb");
    fwrite(flat, sizeof(float), (size_t)total * classes, f);
    fclose(f);
    return 0;
}

static int cmd_layers(int argc, char **argv)
{
    if (argc < 3) return 1;
    network net = parse_network_cfg(argv[2]);
    int i;
    printf("{\"batch\": %d, \"w\": %d, \"h\": %d, \"c\": %d, \"layers\": [", net.batch, net.w, net.h, net.c);
    for (i = 0; i < net.n; ++i) {
        layer l = net.layers[i];
        printf("%s{\"type\": %d, \"w\": %d, \"h\": %d, \"c\": %d, \"out_w\": %d, \"out_h\": %d, \"out_c\": %d, "
               "\"outputs\": %d, \"n\": %d, \"size\": %d, \"stride\": %d, \"pad\": %d}",
               i ? ", " : "", (int)l.type, l.w, l.h, l.c, l.out_w, l.out_h, l.out_c, l.outputs, l.n, l.size,
               l.stride, l.pad);
    }
    printf("]}\n");
    return 0;
}

int main(int argc, char **argv)
{
    gpu_index = -1;
    if (argc < 2) { fprintf(stderr, "usage: darknet_ref forward|region|time|resize ...\n"); return 1; }
    if (!strcmp(argv[1], "forward")) return cmd_forward(argc, argv);
    if (!strcmp(argv[1], "region")) return cmd_region(argc, argv);
    if (!strcmp(argv[1], "time")) return cmd_time(argc, argv);
    if (!strcmp(argv[1], "resize")) return cmd_resize(argc, argv);
    if (!strcmp(argv[1], "layers")) return cmd_layers(argc, argv);
    if (!strcmp(argv[1], "donms")) return cmd_donms(argc, argv);
    if (!strcmp(argv[1], "letterbox")) return cmd_letterbox(argc, argv);
    if (!strcmp(argv[1], "topk")) return cmd_topk(argc, argv);
    fprintf(stderr, "unknown command %s\n", argv[1]);
    return 1;
}
