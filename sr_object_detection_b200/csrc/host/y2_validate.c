/*
 * validate_detector: run a detector over an image list and write the VOC / COCO / ImageNet-detection result files
 * the evaluation scripts read (scripts/voc_eval.py and friends).
 *
 * Reference interface replaced (behavioural spec only): detector.c:244-369 validate_detector with its writers
 * print_detector_detections (:202-221), print_cocos (:175-200), print_imagenet_detections (:223-242),
 * get_coco_image_id (:169-173); data.c:12-23 get_paths; image.c:2069-2095 load_image / load_image_color.
 *
 * Same data-cfg keys (valid, names, results, eval, map), thresholds (.005 / .45), file names and line formats.
 * Differences inside the contract: the images of a chunk go through ONE batched forward pass on the GPU
 * (the reference predicts image by image at batch 1 with four loader threads); every image is then decoded with
 * the caller-facing get_region_boxes / do_nms_sort exactly as the reference's loop does, so the files are
 * byte-identical whenever the network outputs are.  Images are read as binary PPM / PGM (the reference decodes
 * with the third-party stb_image, which is not part of this repository).
 */
#include "y2_host.h"

#include <stdlib.h>
#include <string.h>

/* detector.c:22 */
static const int coco_ids[] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 27,
                               28, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 46, 47, 48, 49, 50, 51, 52, 53,
                               54, 55, 56, 57, 58, 59, 60, 61, 62, 63, 64, 65, 67, 70, 72, 73, 74, 75, 76, 77, 78, 79, 80,
                               81, 82, 84, 85, 86, 87, 88, 89, 90};

/* data.c:12-23: one path per line */
list *get_paths(char *filename)
{
    FILE *file = fopen(filename, "r");
    if (!file) file_error(filename);
    list *lines = make_list();
    char *path;
    while ((path = fgetl(file))) list_insert(lines, path);
    fclose(file);
    return lines;
}

/* ---- binary PPM (P6) / PGM (P5), maxval 255 -> planar floats byte / 255. (image.c load_image_stb) ------- */
static int pnm_token(FILE *f, int *v)
{
    int ch = fgetc(f);
    for (;;) {
        while (ch == ' ' || ch == '\n' || ch == '\r' || ch == '\t') ch = fgetc(f);
        if (ch != '#') break;
        while (ch != '\n' && ch != EOF) ch = fgetc(f);
    }
    if (ch < '0' || ch > '9') return 0;
    *v = 0;
    while (ch >= '0' && ch <= '9') {
        *v = *v * 10 + (ch - '0');
        ch = fgetc(f);
    }
    return 1;
}

/* image.c:2069-2090: `c` planes (grey files are replicated, colour files keep their first c planes), then the
 * optional resize to w x h.  A file that cannot be read ends the program like the reference's loader does
 * (image.c: "Cannot load image", exit(0)). */
image load_image(char *filename, int w, int h, int c)
{
    FILE *f = fopen(filename, "rb");
    char magic[2] = {0, 0};
    int iw = 0, ih = 0, maxval = 0;
    if (!f || fread(magic, 1, 2, f) != 2 || magic[0] != 'P' || (magic[1] != '5' && magic[1] != '6') ||
        !pnm_token(f, &iw) || !pnm_token(f, &ih) || !pnm_token(f, &maxval) || iw <= 0 || ih <= 0 || maxval != 255) {
        fprintf(stderr, "Cannot load image \"%s\"\n", filename);
        exit(0);
    }
    const int src_c = magic[1] == '6' ? 3 : 1;
    if (c <= 0) c = src_c;
    size_t n = (size_t)iw * ih * src_c;
    unsigned char *bytes = (unsigned char *)malloc(n);
    if (fread(bytes, 1, n, f) != n) {
        fprintf(stderr, "Cannot load image \"%s\"\n", filename);
        exit(0);
    }
    fclose(f);
    image im = make_image(iw, ih, c);
    for (int k = 0; k < c; ++k)
        for (int y = 0; y < ih; ++y)
            for (int x = 0; x < iw; ++x)
                im.data[((size_t)k * ih + y) * iw + x] =
                    (float)bytes[((size_t)y * iw + x) * src_c + (src_c == 3 && k < 3 ? k : 0)] / 255.;
    free(bytes);
    if ((h && w) && (h != im.h || w != im.w)) {
        image resized = resize_image(im, w, h);
        free_image(im);
        im = resized;
    }
    return im;
}

image load_image_color(char *filename, int w, int h)
{
    return load_image(filename, w, h, 3);
}

/* ---- writers ---------------------------------------------------------------------------------------------- */
typedef struct { float xmin, ymin, xmax, ymax; } corners;

/* the corner arithmetic shared by the three writers (detector.c:180-188): w/2. is a double expression */
static corners clip_box(box b, int w, int h)
{
    corners c;
    c.xmin = b.x - b.w / 2.;
    c.xmax = b.x + b.w / 2.;
    c.ymin = b.y - b.h / 2.;
    c.ymax = b.y + b.h / 2.;
    if (c.xmin < 0) c.xmin = 0;
    if (c.ymin < 0) c.ymin = 0;
    if (c.xmax > w) c.xmax = w;
    if (c.ymax > h) c.ymax = h;
    return c;
}

static int get_coco_image_id(char *filename)
{
    char *p = strrchr(filename, '_');
    return p ? atoi(p + 1) : 0;
}

static void print_cocos(FILE *fp, char *image_path, box *boxes, float **probs, int num_boxes, int classes, int w, int h)
{
    const int image_id = get_coco_image_id(image_path);
    for (int i = 0; i < num_boxes; ++i) {
        const corners c = clip_box(boxes[i], w, h);
        const float bx = c.xmin, by = c.ymin, bw = c.xmax - c.xmin, bh = c.ymax - c.ymin;
        for (int j = 0; j < classes; ++j)
            if (probs[i][j])
                fprintf(fp, "{\"image_id\":%d, \"category_id\":%d, \"bbox\":[%f, %f, %f, %f], \"score\":%f},\n", image_id,
                        coco_ids[j], bx, by, bw, bh, probs[i][j]);
    }
}

void print_detector_detections(FILE **fps, char *id, box *boxes, float **probs, int total, int classes, int w, int h)
{
    for (int i = 0; i < total; ++i) {
        const corners c = clip_box(boxes[i], w, h);
        for (int j = 0; j < classes; ++j)
            if (probs[i][j]) fprintf(fps[j], "%s %f %f %f %f %f\n", id, probs[i][j], c.xmin, c.ymin, c.xmax, c.ymax);
    }
}

void print_imagenet_detections(FILE *fp, int id, box *boxes, float **probs, int total, int classes, int w, int h)
{
    for (int i = 0; i < total; ++i) {
        const corners c = clip_box(boxes[i], w, h);
        for (int j = 0; j < classes; ++j)
            if (probs[i][j]) fprintf(fp, "%d %d %f %f %f %f %f\n", id, j + 1, probs[i][j], c.xmin, c.ymin, c.xmax, c.ymax);
    }
}

void validate_detector(char *datacfg, char *cfgfile, char *weightfile)
{
    list *options = read_data_cfg(datacfg);
    char *valid_images = option_find_str(options, "valid", "data/train.list");
    char *name_list = option_find_str(options, "names", "data/names.list");
    char *prefix = option_find_str(options, "results", "results");
    char **names = get_labels(name_list);
    char *mapf = option_find_str(options, "map", 0);
    int *map = 0;
    if (mapf) map = read_map(mapf);

    network net = parse_network_cfg(cfgfile);
    if (weightfile) load_weights(&net, weightfile);
    /* the reference predicts at batch 1; here a chunk of images shares one forward pass */
    int chunk = net.batch > 1 ? net.batch : 16;
    if (getenv("Y2_VALID_BATCH")) chunk = atoi(getenv("Y2_VALID_BATCH"));
    if (chunk < 1) chunk = 1;
    set_batch_network(&net, chunk);
    fprintf(stderr, "Learning Rate: %g, Momentum: %g, Decay: %g\n", net.learning_rate, net.momentum, net.decay);

    char *base = "comp4_det_test_";
    list *plist = get_paths(valid_images);
    char **paths = (char **)list_to_array(plist);

    layer l = net.layers[net.n - 1];
    if (l.type != REGION) error("validate_detector: the last layer is not a region layer");
    int classes = l.classes;

    char buff[1024];
    char *type = option_find_str(options, "eval", "voc");
    FILE *fp = 0;
    FILE **fps = 0;
    int coco = 0, imagenet = 0;
    if (0 == strcmp(type, "coco")) {
        snprintf(buff, 1024, "%s/coco_results.json", prefix);
        fp = fopen(buff, "w");
        if (!fp) file_error(buff);
        fprintf(fp, "[\n");
        coco = 1;
    } else if (0 == strcmp(type, "imagenet")) {
        snprintf(buff, 1024, "%s/imagenet-detection.txt", prefix);
        fp = fopen(buff, "w");
        if (!fp) file_error(buff);
        imagenet = 1;
        classes = 200;
    } else {
        fps = (FILE **)calloc(classes, sizeof(FILE *));
        for (int j = 0; j < classes; ++j) {
            snprintf(buff, 1024, "%s/%s%s.txt", prefix, base, names[j]);
            fps[j] = fopen(buff, "w");
            if (!fps[j]) file_error(buff);
        }
    }

    const int total = l.w * l.h * l.n;
    box *boxes = (box *)calloc(total, sizeof(box));
    float **probs = (float **)calloc(total, sizeof(float *));
    /* rows as wide as the widest write of get_region_boxes (the flat branch ignores `map`) */
    const int row = classes > l.classes ? classes : l.classes;
    for (int j = 0; j < total; ++j) probs[j] = (float *)calloc(row, sizeof(float));

    const int m = plist->size;
    const float thresh = .005;
    const float nms = .45;
    const size_t per_image = (size_t)net.w * net.h * net.c;
    float *X = (float *)calloc((size_t)chunk * per_image, sizeof(float));
    int *ws = (int *)calloc(chunk, sizeof(int)), *hs = (int *)calloc(chunk, sizeof(int));

    for (int i = 0; i < m; i += chunk) {
        const int n = m - i < chunk ? m - i : chunk;
        fprintf(stderr, "%d\n", i + n);
        memset(X, 0, (size_t)chunk * per_image * sizeof(float));
        for (int t = 0; t < n; ++t) {
            image im = load_image_color(paths[i + t], 0, 0);
            image sized = resize_image(im, net.w, net.h);
            ws[t] = im.w;
            hs[t] = im.h;
            memcpy(X + (size_t)t * per_image, sized.data, per_image * sizeof(float));
            free_image(im);
            free_image(sized);
        }
        network_predict(net, X);
        l = net.layers[net.n - 1];
        for (int t = 0; t < n; ++t) {
            char *path = paths[i + t];
            char *id = basecfg(path);
            layer lt = l;
            lt.output = l.output + (size_t)t * l.outputs;
            const int w = ws[t], h = hs[t];
            get_region_boxes(lt, w, h, thresh, probs, boxes, 0, map);
            if (nms) do_nms_sort(boxes, probs, total, classes, nms);
            if (coco) print_cocos(fp, path, boxes, probs, total, classes, w, h);
            else if (imagenet) print_imagenet_detections(fp, i + t + 1, boxes, probs, total, classes, w, h);
            else print_detector_detections(fps, id, boxes, probs, total, classes, w, h);
            free(id);
        }
    }
    for (int j = 0; j < classes; ++j)
        if (fps) fclose(fps[j]);
    if (coco) {
        fseek(fp, -2, SEEK_CUR);
        fprintf(fp, "\n]\n");
    }
    if (fp) fclose(fp);
    free(fps);
    for (int j = 0; j < total; ++j) free(probs[j]);
    free(probs);
    free(boxes);
    free(X);
    free(ws);
    free(hs);
    free(paths);
    free_network(net);
}

/* ---- validate_detector_recall (detector.c:371-450): proposals, mean best IoU and recall against label files ------ */

/* utils.c:158-172: the FIRST occurrence of `orig` in `str` replaced by `rep`; `output` may be `str` itself */
void find_replace(char *str, char *orig, char *rep, char *output)
{
    char copy[4096];
    snprintf(copy, sizeof(copy), "%s", str);
    char *hit = strstr(copy, orig);
    if (!hit) {
        snprintf(output, 4096, "%s", copy);
        return;
    }
    *hit = 0;
    snprintf(output, 4096, "%s%s%s", copy, rep, hit + strlen(orig));
}

/* data.c:135-159: "id x y w h" records until the first line that does not parse */
box_label *read_boxes(char *filename, int *n)
{
    FILE *file = fopen(filename, "r");
    if (!file) file_error(filename);
    int count = 0, cap = 0, id;
    float x, y, w, h;
    box_label *out = (box_label *)calloc(1, sizeof(box_label));
    while (fscanf(file, "%d %f %f %f %f", &id, &x, &y, &w, &h) == 5) {
        if (count == cap) {
            cap = cap ? 2 * cap : 8;
            out = (box_label *)realloc(out, (size_t)cap * sizeof(box_label));
        }
        box_label b;
        b.id = id;
        b.x = x; b.y = y; b.w = w; b.h = h;
        b.left = x - w / 2;
        b.right = x + w / 2;
        b.top = y - h / 2;
        b.bottom = y + h / 2;
        out[count++] = b;
    }
    fclose(file);
    *n = count;
    return out;
}

/* the label file that belongs to an image (detector.c:413-418) */
static void label_path_of(char *image_path, char *out)
{
    find_replace(image_path, "images", "labels", out);
    find_replace(out, "JPEGImages", "labels", out);
    find_replace(out, ".jpg", ".txt", out);
    find_replace(out, ".JPEG", ".txt", out);
    find_replace(out, ".png", ".txt", out);
}

/* Objectness proposals of every image (get_region_boxes with only_objectness, do_nms over that one column at .4)
 * against the ground-truth boxes of its label file: running proposal count, sum of the best IoU per truth box and
 * the number of truth boxes matched above .5, one stderr line per image in the reference's format.  As in
 * validate_detector a chunk of images shares one forward pass; the decode runs per image through the caller-facing
 * calls, so the lines are identical whenever the network outputs are. */
void validate_detector_recall(char *datacfg, char *cfgfile, char *weightfile)
{
    network net = parse_network_cfg(cfgfile);
    if (weightfile) load_weights(&net, weightfile);
    int chunk = 16;
    if (getenv("Y2_VALID_BATCH")) chunk = atoi(getenv("Y2_VALID_BATCH"));
    if (chunk < 1) chunk = 1;
    set_batch_network(&net, chunk);
    fprintf(stderr, "Learning Rate: %g, Momentum: %g, Decay: %g\n", net.learning_rate, net.momentum, net.decay);

    list *options = read_data_cfg(datacfg);
    list *plist = get_paths(option_find_str(options, "valid", "data/train.txt"));
    char **paths = (char **)list_to_array(plist);
    layer l = net.layers[net.n - 1];
    if (l.type != REGION) error("validate_detector_recall: the last layer is not a region layer");
    const int nboxes = l.w * l.h * l.n;
    box *boxes = (box *)calloc(nboxes, sizeof(box));
    float **probs = (float **)calloc(nboxes, sizeof(float *));
    for (int j = 0; j < nboxes; ++j) probs[j] = (float *)calloc(l.classes, sizeof(float));

    const float thresh = .2, iou_thresh = .5, nms = .4;
    int total = 0, correct = 0, proposals = 0;
    float avg_iou = 0;
    const int m = plist->size;
    const size_t per_image = (size_t)net.w * net.h * net.c;
    float *X = (float *)calloc((size_t)chunk * per_image, sizeof(float));
    for (int i0 = 0; i0 < m; i0 += chunk) {
        const int n = m - i0 < chunk ? m - i0 : chunk;
        memset(X, 0, (size_t)chunk * per_image * sizeof(float));
        for (int t = 0; t < n; ++t) {
            image orig = load_image_color(paths[i0 + t], 0, 0);
            image sized = resize_image(orig, net.w, net.h);
            memcpy(X + (size_t)t * per_image, sized.data, per_image * sizeof(float));
            free_image(orig);
            free_image(sized);
        }
        network_predict(net, X);
        l = net.layers[net.n - 1];
        for (int t = 0; t < n; ++t) {
            const int i = i0 + t;
            layer lt = l;
            lt.output = l.output + (size_t)t * l.outputs;
            get_region_boxes(lt, 1, 1, thresh, probs, boxes, 1, 0);
            if (nms) do_nms(boxes, probs, nboxes, 1, nms);

            char labelpath[4096];
            label_path_of(paths[i], labelpath);
            int num_labels = 0;
            box_label *truth = read_boxes(labelpath, &num_labels);
            for (int k = 0; k < nboxes; ++k)
                if (probs[k][0] > thresh) ++proposals;
            for (int j = 0; j < num_labels; ++j) {
                ++total;
                box tb = {truth[j].x, truth[j].y, truth[j].w, truth[j].h};
                float best_iou = 0;
                for (int k = 0; k < nboxes; ++k) {
                    const float iou = box_iou(boxes[k], tb);
                    if (probs[k][0] > thresh && iou > best_iou) best_iou = iou;
                }
                avg_iou += best_iou;
                if (best_iou > iou_thresh) ++correct;
            }
            fprintf(stderr, "%5d %5d %5d\tRPs/Img: %.2f\tIOU: %.2f%%\tRecall:%.2f%%\n", i, correct, total,
                    (float)proposals / (i + 1), avg_iou * 100 / total, 100. * correct / total);
            free(truth);
        }
    }
    free(X);
    for (int j = 0; j < nboxes; ++j) free(probs[j]);
    free(probs);
    free(boxes);
    free(paths);
    free_list_contents(plist);
    free_list(plist);
    free_network(net);
}

