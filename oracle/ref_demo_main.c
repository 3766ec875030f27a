/* TEST INFRASTRUCTURE: the per-frame chain of demo.c:71-107 (detect_in_thread) over the reference's own CPU functions
 * - network_predict, mean_arrays over a ring of FRAMES = 3 outputs, get_region_boxes, do_nms at .4 - on a given
 * sequence of frames.  demo.c itself only compiles with OpenCV (capture + drawing); this driver runs the arithmetic part
 * of its loop, statement for statement, and dumps boxes and probabilities per frame.
 *   ref_demo <cfg> <weights> <frames.u8> <n_frames> <thresh> <outdir>       frames: uint8 RGB [n][h][w][3] at net size */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "network.h"
#include "parser.h"
#include "region_layer.h"
#include "box.h"
#include "utils.h"
#include "image.h"
#define FRAMES 3
extern int gpu_index;
detectBoxes *GlobleObjBoxes;
int GlobleObjBoxesNum;
int main(int argc, char **argv)
{
    if (argc < 7) return 1;
    gpu_index = -1;
    int n = atoi(argv[4]), f, j, k, y, x;
    float thresh = atof(argv[5]);
    network net = parse_network_cfg(argv[1]);
    load_weights(&net, argv[2]);
    set_batch_network(&net, 1);
    layer l = net.layers[net.n - 1];
    int total = l.w * l.h * l.n;
    float *avg = calloc(l.outputs, sizeof(float));
    float *predictions[FRAMES];
    for (j = 0; j < FRAMES; ++j) predictions[j] = calloc(l.outputs, sizeof(float));
    box *boxes = calloc(total, sizeof(box));
    float **probs = calloc(total, sizeof(float *));
    for (j = 0; j < total; ++j) probs[j] = calloc(l.classes, sizeof(float));
    size_t fb = (size_t)net.w * net.h * 3;
    unsigned char *rgb = malloc(fb);
    FILE *fp = fopen(argv[3], "rb");
    int demo_index = 0;
    for (f = 0; f < n; ++f) {
        if (fread(rgb, 1, fb, fp) != fb) return 2;
        image in_s = make_image(net.w, net.h, 3); /* ipl_to_image: data / 255. per channel plane */
        for (k = 0; k < 3; ++k)
            for (y = 0; y < net.h; ++y)
                for (x = 0; x < net.w; ++x) in_s.data[(k * net.h + y) * net.w + x] = rgb[(y * net.w + x) * 3 + k] / 255.;
        /* detect_in_thread, demo.c:73-92 */
        float nms = .4;
        float *prediction = network_predict(net, in_s.data);
        memcpy(predictions[demo_index], prediction, l.outputs * sizeof(float));
        mean_arrays(predictions, FRAMES, l.outputs, avg);
        l.output = avg;
        free_image(in_s);
        get_region_boxes(l, 1, 1, thresh, probs, boxes, 0, 0);
        if (nms > 0) do_nms(boxes, probs, total, l.classes, nms);
        demo_index = (demo_index + 1) % FRAMES;
        char path[4096];
        snprintf(path, sizeof(path), "%s/frame_%03d.f32", argv[6], f);
        FILE *out = fopen(path, "wb");
        fwrite(boxes, sizeof(box), total, out);
        for (j = 0; j < total; ++j) fwrite(probs[j], sizeof(float), l.classes, out);
        fclose(out);
    }
    return 0;
}
