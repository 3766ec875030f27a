/*
 * TEST INFRASTRUCTURE — not part of the product.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may build or execute this file.
 *
 * y2_oracle: a plain-C CPU restatement of the reference's YOLOv2 detection forward path
 * (Darknet fork under /root/reference/src_yolo2).  Every function names the reference
 * file:line whose arithmetic it restates; expression types (float vs double promotion,
 * separate roundings, accumulation order) are kept so that results are bit-identical to the
 * reference's CPU build (`gcc -O2 -ffp-contract=off`).
 *
 * Parity pinning: tests/golden/ holds outputs of the UNMODIFIED reference (oracle/_ref/
 * darknet_ref, built by oracle/Makefile from the reference sources) on seeded inputs, made by
 * tests/golden/make_golden.py; tests/test_oracle_golden.py requires this program to reproduce
 * them bit for bit.  The command line mirrors oracle/ref_driver.c so either binary can serve
 * a test:
 *
 *   y2_oracle forward <cfg> <weights|-> <input.f32> <outdir> <thresh> <nms> <dump_layers>
 *   y2_oracle region  <cfg> <region_in.f32> <outdir> <thresh> <nms>
 *   y2_oracle time    <cfg> <weights|-> <input.f32> <thresh> <nms> <warmup> <iters>
 *   y2_oracle resize  <in.f32> <c> <h> <w> <out_h> <out_w> <out.f32>
 *   y2_oracle layers  <cfg>                      (layer table as JSON, parser parity)
 * Environment: Y2_USE_MAP=1 passes the region layer's `map` to get_region_boxes.
 *
 * The only sort used (do_nms_sort) is an explicit stable merge sort: glibc 2.39's qsort is a
 * stable merge sort whenever it can allocate its scratch, which is what the reference ran on.
 */
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef enum { T_CONV, T_MAXPOOL, T_REORG, T_ROUTE, T_REGION, T_SHORTCUT, T_AVGPOOL, T_SOFTMAX, T_COST } ltype;
typedef enum { A_LOGISTIC, A_LINEAR, A_LEAKY, A_RELU } atype;

typedef struct {
    int n, groups;
    int *parent, *group_size, *group_offset;
} otree;

typedef struct {
    ltype type;
    atype act;
    int batch, h, w, c, out_h, out_w, out_c, inputs, outputs;
    int n, size, stride, pad, bn, flipped;
    int classes, coords, softmax, classfix;
    int index;                /* shortcut source */
    int n_in, *in_layers, *in_sizes;
    int groups;
    float temperature;
    float *weights, *biases, *scales, *mean, *var;
    float *output;
    otree *tree;
    int *map;
} olayer;

typedef struct {
    int n, batch, h, w, c, inputs;
    olayer *l;
    float *workspace;
} onet;

static void die(const char *msg)
{
    fprintf(stderr, "y2_oracle: %s\n", msg);
    exit(2);
}

/* ---------------------------------------------------------------------------------------------
 * cfg reader.  Semantics of read_cfg (parser.c:702-735), read_option (option_list.c:35-51) and
 * strip (utils.c:230-241): ALL blanks are removed from a line, '[' starts a section, '#' ';'
 * and empty lines are skipped, the first '=' splits key and value.
 * ------------------------------------------------------------------------------------------- */
typedef struct { char *key, *val; } kv;
typedef struct { char *type; kv *opt; int n; } section;

static void strip_all(char *s)
{
    size_t i, off = 0, len = strlen(s);
    for (i = 0; i < len; ++i) {
        char ch = s[i];
        if (ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r') ++off;
        else s[i - off] = ch;
    }
    s[len - off] = 0;
}

static section *read_sections(const char *path, int *count)
{
    FILE *f = fopen(path, "r");
    if (!f) { fprintf(stderr, "Couldn't open file: %s\n", path); exit(0); }
    section *s = 0;
    int ns = 0;
    char line[8192];
    while (fgets(line, sizeof(line), f)) {
        strip_all(line);
        if (!line[0] || line[0] == '#' || line[0] == ';') continue;
        if (line[0] == '[') {
            s = realloc(s, (ns + 1) * sizeof(section));
            s[ns].type = strdup(line);
            s[ns].opt = 0;
            s[ns].n = 0;
            ++ns;
            continue;
        }
        char *eq = strchr(line, '=');
        if (!eq || !ns) continue; /* parser.c:726-729: complain and go on */
        *eq = 0;
        section *c = &s[ns - 1];
        c->opt = realloc(c->opt, (c->n + 1) * sizeof(kv));
        c->opt[c->n].key = strdup(line);
        c->opt[c->n].val = strdup(eq + 1);
        ++c->n;
    }
    fclose(f);
    *count = ns;
    return s;
}

/* option_find (option_list.c:74-86) returns the first match */
static const char *opt_str(const section *s, const char *key, const char *def)
{
    for (int i = 0; i < s->n; ++i)
        if (!strcmp(s->opt[i].key, key)) return s->opt[i].val;
    return def;
}
static int opt_int(const section *s, const char *key, int def)
{
    const char *v = opt_str(s, key, 0);
    return v ? atoi(v) : def;
}
static float opt_float(const section *s, const char *key, float def)
{
    const char *v = opt_str(s, key, 0);
    return v ? (float)atof(v) : def;
}

/* tree.c:53-103 read_tree: a group is a maximal run of consecutive nodes with one parent */
static otree *load_tree(const char *path)
{
    FILE *f = fopen(path, "r");
    if (!f) { fprintf(stderr, "Couldn't open file: %s\n", path); exit(0); }
    otree *t = calloc(1, sizeof(otree));
    char line[4096], id[1024];
    int last_parent = -1, group_size = 0, groups = 0, n = 0;
    while (fgets(line, sizeof(line), f)) {
        int parent = -1;
        if (sscanf(line, "%1023s %d", id, &parent) < 1) continue;
        t->parent = realloc(t->parent, (n + 1) * sizeof(int));
        t->parent[n] = parent;
        if (parent != last_parent) {
            ++groups;
            t->group_offset = realloc(t->group_offset, groups * sizeof(int));
            t->group_size = realloc(t->group_size, groups * sizeof(int));
            t->group_offset[groups - 1] = n - group_size;
            t->group_size[groups - 1] = group_size;
            group_size = 0;
            last_parent = parent;
        }
        ++n;
        ++group_size;
    }
    ++groups;
    t->group_offset = realloc(t->group_offset, groups * sizeof(int));
    t->group_size = realloc(t->group_size, groups * sizeof(int));
    t->group_offset[groups - 1] = n - group_size;
    t->group_size[groups - 1] = group_size;
    t->n = n;
    t->groups = groups;
    fclose(f);
    return t;
}

/* utils.c:17-33 read_map: one integer per line */
static int *load_map(const char *path)
{
    FILE *f = fopen(path, "r");
    if (!f) { fprintf(stderr, "Couldn't open file: %s\n", path); exit(0); }
    int *m = 0, n = 0;
    char line[256];
    while (fgets(line, sizeof(line), f)) {
        m = realloc(m, (n + 1) * sizeof(int));
        m[n++] = atoi(line);
    }
    fclose(f);
    return m;
}

static atype parse_act(const char *s)
{
    if (!strcmp(s, "logistic")) return A_LOGISTIC;
    if (!strcmp(s, "linear")) return A_LINEAR;
    if (!strcmp(s, "leaky")) return A_LEAKY;
    if (!strcmp(s, "relu")) return A_RELU;
    die("activation outside the hot path");
    return A_LINEAR;
}

/* parse_network_cfg (parser.c:585-700) for the layer types of the north-star cfgs */
static onet parse_cfg(const char *path)
{
    int ns = 0;
    section *sec = read_sections(path, &ns);
    if (!ns) die("Config file has no sections");
    if (strcmp(sec[0].type, "[net]") && strcmp(sec[0].type, "[network]")) die("First section must be [net] or [network]");
    onet net;
    memset(&net, 0, sizeof(net));
    /* parse_net_options, parser.c:504-577 */
    net.batch = opt_int(&sec[0], "batch", 1);
    int subdivs = opt_int(&sec[0], "subdivisions", 1);
    int time_steps = opt_int(&sec[0], "time_steps", 1);
    net.batch /= subdivs;
    net.batch *= time_steps;
    net.h = opt_int(&sec[0], "height", 0);
    net.w = opt_int(&sec[0], "width", 0);
    net.c = opt_int(&sec[0], "channels", 0);
    net.inputs = opt_int(&sec[0], "inputs", net.h * net.w * net.c);
    net.n = ns - 1;
    net.l = calloc(net.n, sizeof(olayer));
    int h = net.h, w = net.w, c = net.c, inputs = net.inputs;
    size_t workspace = 0;
    for (int i = 0; i < net.n; ++i) {
        const section *s = &sec[i + 1];
        olayer *l = &net.l[i];
        l->batch = net.batch;
        if (!strcmp(s->type, "[convolutional]") || !strcmp(s->type, "[conv]")) {
            /* parse_convolutional parser.c:139-171, make_convolutional_layer conv_layer.c:182-319 */
            l->type = T_CONV;
            l->n = opt_int(s, "filters", 1);
            l->size = opt_int(s, "size", 1);
            l->stride = opt_int(s, "stride", 1);
            int pad = opt_int(s, "pad", 0);
            l->pad = opt_int(s, "padding", 0);
            if (pad) l->pad = l->size / 2;
            l->act = parse_act(opt_str(s, "activation", "logistic"));
            l->bn = opt_int(s, "batch_normalize", 0);
            l->flipped = opt_int(s, "flipped", 0);
            if (!(h && w && c)) die("Layer before convolutional layer must output image.");
            l->h = h; l->w = w; l->c = c;
            l->out_h = (h + 2 * l->pad - l->size) / l->stride + 1; /* conv_layer.c:75-83 */
            l->out_w = (w + 2 * l->pad - l->size) / l->stride + 1;
            l->out_c = l->n;
            l->outputs = l->out_h * l->out_w * l->out_c;
            l->inputs = h * w * c;
            size_t nw = (size_t)c * l->n * l->size * l->size;
            l->weights = calloc(nw, sizeof(float));
            l->biases = calloc(l->n, sizeof(float));
            if (l->bn) {
                l->scales = calloc(l->n, sizeof(float));
                for (int k = 0; k < l->n; ++k) l->scales[k] = 1;
                l->mean = calloc(l->n, sizeof(float));
                l->var = calloc(l->n, sizeof(float));
            }
            size_t ws = (size_t)l->out_h * l->out_w * l->size * l->size * c * sizeof(float);
            if (ws > workspace) workspace = ws;
        } else if (!strcmp(s->type, "[maxpool]") || !strcmp(s->type, "[max]")) {
            /* parse_maxpool parser.c:359-374, make_maxpool_layer maxpool_layer.c:20-51 */
            l->type = T_MAXPOOL;
            l->stride = opt_int(s, "stride", 1);
            l->size = opt_int(s, "size", l->stride);
            l->pad = opt_int(s, "padding", (l->size - 1) / 2);
            l->h = h; l->w = w; l->c = c;
            l->out_w = (w + 2 * l->pad) / l->stride;
            l->out_h = (h + 2 * l->pad) / l->stride;
            l->out_c = c;
            l->outputs = l->out_h * l->out_w * l->out_c;
            l->inputs = h * w * c;
        } else if (!strcmp(s->type, "[reorg]")) {
            /* parse_reorg parser.c:343-357, make_reorg_layer reorg_layer.c:7-45 */
            l->type = T_REORG;
            l->stride = opt_int(s, "stride", 1);
            if (opt_int(s, "reverse", 0)) die("reverse reorg is outside the hot path");
            l->h = h; l->w = w; l->c = c;
            l->out_w = w / l->stride;
            l->out_h = h / l->stride;
            l->out_c = c * l->stride * l->stride;
            l->outputs = l->out_h * l->out_w * l->out_c;
            l->inputs = h * w * c;
        } else if (!strcmp(s->type, "[route]")) {
            /* parse_route parser.c:450-489 */
            l->type = T_ROUTE;
            const char *v = opt_str(s, "layers", 0);
            if (!v) die("Route Layer must specify input layers");
            int n = 1;
            for (const char *p = v; *p; ++p) if (*p == ',') ++n;
            l->n_in = n;
            l->in_layers = calloc(n, sizeof(int));
            l->in_sizes = calloc(n, sizeof(int));
            const char *p = v;
            for (int k = 0; k < n; ++k) {
                int idx = atoi(p);
                const char *q = strchr(p, ',');
                p = q ? q + 1 : p;
                if (idx < 0) idx = i + idx;
                l->in_layers[k] = idx;
                l->in_sizes[k] = net.l[idx].outputs;
                l->outputs += net.l[idx].outputs;
            }
            l->inputs = l->outputs;
            olayer *first = &net.l[l->in_layers[0]];
            l->out_w = first->out_w; l->out_h = first->out_h; l->out_c = first->out_c;
            for (int k = 1; k < n; ++k) {
                olayer *nx = &net.l[l->in_layers[k]];
                if (nx->out_w == first->out_w && nx->out_h == first->out_h) l->out_c += nx->out_c;
                else l->out_h = l->out_w = l->out_c = 0;
            }
        } else if (!strcmp(s->type, "[region]")) {
            /* parse_region parser.c:236-284, make_region_layer region_layer.c:14-51 */
            l->type = T_REGION;
            l->coords = opt_int(s, "coords", 4);
            l->classes = opt_int(s, "classes", 20);
            l->n = opt_int(s, "num", 1);
            l->softmax = opt_int(s, "softmax", 0);
            l->classfix = opt_int(s, "classfix", 0);
            l->h = h; l->w = w; /* out_h/out_w stay 0, as in make_region_layer */
            l->outputs = h * w * l->n * (l->classes + l->coords + 1);
            l->inputs = l->outputs;
            if (l->outputs != inputs) die("region layer: outputs != inputs (parser.c:243 assert)");
            l->biases = calloc((size_t)l->n * 2, sizeof(float));
            for (int k = 0; k < l->n * 2; ++k) l->biases[k] = .5;
            const char *tf = opt_str(s, "tree", 0);
            if (tf) l->tree = load_tree(tf);
            const char *mf = opt_str(s, "map", 0);
            if (mf) l->map = load_map(mf);
            const char *a = opt_str(s, "anchors", 0);
            if (a) {
                int n = 1;
                for (const char *p = a; *p; ++p) if (*p == ',') ++n;
                for (int k = 0; k < n; ++k) {
                    l->biases[k] = (float)atof(a);
                    const char *q = strchr(a, ',');
                    a = q ? q + 1 : a;
                }
            }
        } else if (!strcmp(s->type, "[shortcut]")) {
            /* parse_shortcut parser.c:415-430, make_shortcut_layer shortcut_layer.c:7-34 */
            l->type = T_SHORTCUT;
            const char *v = opt_str(s, "from", 0);
            if (!v) die("shortcut without from=");
            int idx = atoi(v);
            if (idx < 0) idx = i + idx;
            l->index = idx;
            olayer *from = &net.l[idx];
            l->w = from->out_w; l->h = from->out_h; l->c = from->out_c;
            l->out_w = w; l->out_h = h; l->out_c = c;
            l->outputs = w * h * c;
            l->inputs = l->outputs;
            l->act = parse_act(opt_str(s, "activation", "linear"));
        } else if (!strcmp(s->type, "[avgpool]") || !strcmp(s->type, "[avg]")) {
            l->type = T_AVGPOOL; /* avgpool_layer.c:6-25 */
            l->h = h; l->w = w; l->c = c;
            l->out_w = 1; l->out_h = 1; l->out_c = c;
            l->outputs = c;
            l->inputs = h * w * c;
        } else if (!strcmp(s->type, "[softmax]") || !strcmp(s->type, "[soft]")) {
            l->type = T_SOFTMAX; /* parse_softmax parser.c:226-234 */
            l->groups = opt_int(s, "groups", 1);
            l->temperature = opt_float(s, "temperature", 1);
            l->inputs = inputs;
            l->outputs = inputs;
            const char *tf = opt_str(s, "tree", 0);
            if (tf) l->tree = load_tree(tf);
        } else if (!strcmp(s->type, "[cost]")) {
            l->type = T_COST; /* no-op at inference, cost_layer.c:75 */
            l->inputs = inputs;
            l->outputs = inputs;
        } else {
            fprintf(stderr, "y2_oracle: layer type %s is outside the hot path\n", s->type);
            exit(2);
        }
        if (l->type != T_COST) l->output = calloc((size_t)l->batch * l->outputs, sizeof(float));
        /* parser.c:677-682 */
        h = l->out_h; w = l->out_w; c = l->out_c; inputs = l->outputs;
    }
    net.workspace = calloc(1, workspace ? workspace : 4);
    return net;
}

/* load_weights_upto / load_convolutional_weights, parser.c:1009-1082, 963-1006 */
static void load_weights(onet *net, const char *path)
{
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "Couldn't open file: %s\n", path); exit(0); }
    int major, minor, revision;
    if (fread(&major, 4, 1, f) != 1 || fread(&minor, 4, 1, f) != 1 || fread(&revision, 4, 1, f) != 1) die("short weights header");
    if (major * 10 + minor >= 2) { unsigned long long seen; if (fread(&seen, 8, 1, f) != 1) die("short header"); }
    else { int seen; if (fread(&seen, 4, 1, f) != 1) die("short header"); }
    for (int i = 0; i < net->n; ++i) {
        olayer *l = &net->l[i];
        if (l->type != T_CONV) continue;
        size_t num = (size_t)l->n * l->c * l->size * l->size;
        size_t got = fread(l->biases, 4, l->n, f);
        if (l->bn) {
            got += fread(l->scales, 4, l->n, f);
            got += fread(l->mean, 4, l->n, f);
            got += fread(l->var, 4, l->n, f);
        }
        got += fread(l->weights, 4, num, f);
        (void)got;
        if (l->flipped) { /* transpose_matrix(l.weights, l.c*l.size*l.size, l.n), parser.c:997-999 */
            int rows = l->c * l->size * l->size, cols = l->n;
            float *t = calloc(num, sizeof(float));
            for (int x = 0; x < rows; ++x)
                for (int y = 0; y < cols; ++y) t[(size_t)y * rows + x] = l->weights[(size_t)x * cols + y];
            memcpy(l->weights, t, num * sizeof(float));
            free(t);
        }
    }
    fclose(f);
}

/* ---------------------------------------------------------------------------------------------
 * layer arithmetic
 * ------------------------------------------------------------------------------------------- */

/* activations.h:35,39-41,34 */
static inline float act_logistic(float x) { return 1. / (1. + exp(-x)); }
static inline float act_leaky(float x) { return (x > 0) ? x : .1 * x; }
static inline float act_relu(float x) { return x * (x > 0); }

static void activate(float *x, size_t n, atype a)
{
    size_t i;
    switch (a) {
    case A_LINEAR: break;
    case A_LEAKY: for (i = 0; i < n; ++i) x[i] = act_leaky(x[i]); break;
    case A_LOGISTIC: for (i = 0; i < n; ++i) x[i] = act_logistic(x[i]); break;
    case A_RELU: for (i = 0; i < n; ++i) x[i] = act_relu(x[i]); break;
    }
}

/* im2col_cpu, im2col.c:16-39: row index = c_im*k*k + kh*k + kw, zero outside the image */
static void unfold(const float *im, int channels, int height, int width, int k, int stride, int pad, float *col)
{
    const int oh = (height + 2 * pad - k) / stride + 1;
    const int ow = (width + 2 * pad - k) / stride + 1;
    const int rows = channels * k * k;
#pragma omp parallel for
    for (int r = 0; r < rows; ++r) {
        const int kw = r % k, kh = (r / k) % k, ci = r / k / k;
        float *dst = col + (size_t)r * oh * ow;
        for (int y = 0; y < oh; ++y) {
            const int iy = kh + y * stride - pad;
            for (int x = 0; x < ow; ++x) {
                const int ix = kw + x * stride - pad;
                float v = 0;
                if (iy >= 0 && ix >= 0 && iy < height && ix < width) v = im[ix + width * (iy + height * ci)];
                dst[(size_t)y * ow + x] = v;
            }
        }
    }
}

/* gemm_cpu / gemm_nn with ALPHA = BETA = 1, gemm.c:141-167, 74-88: for every output row,
 * k ascending, C[i][j] += A[i][k] * B[k][j] in float (no FMA).  Rows are independent, so the
 * result does not depend on the thread count. */
static void matmul_acc(int M, int N, int K, const float *A, const float *B, float *C)
{
#pragma omp parallel for
    for (int i = 0; i < M; ++i) {
        float *c = C + (size_t)i * N;
        for (int k = 0; k < K; ++k) {
            const float a = 1.f * A[(size_t)i * K + k];
            const float *b = B + (size_t)k * N;
            for (int j = 0; j < N; ++j) c[j] += a * b[j];
        }
    }
}

/* forward_convolutional_layer, convolutional_layer.c:435-474, with the inference branch of
 * forward_batchnorm_layer (batchnorm_layer.c:141-144): normalize_cpu (blas.c:115-126),
 * scale_bias (conv_layer.c:413-423), add_bias (401-411), activate_array (activations.c:95-101) */
static void fwd_conv(olayer *l, const float *in, float *workspace)
{
    const int m = l->n, k = l->size * l->size * l->c, n = l->out_h * l->out_w;
    memset(l->output, 0, (size_t)l->batch * l->outputs * sizeof(float));
    for (int b = 0; b < l->batch; ++b) {
        unfold(in + (size_t)b * l->c * l->h * l->w, l->c, l->h, l->w, l->size, l->stride, l->pad, workspace);
        matmul_acc(m, n, k, l->weights, workspace, l->output + (size_t)b * m * n);
    }
#pragma omp parallel for
    for (int bf = 0; bf < l->batch * m; ++bf) {
        const int f = bf % m;
        float *x = l->output + (size_t)bf * n;
        if (l->bn) {
            for (int i = 0; i < n; ++i) x[i] = (x[i] - l->mean[f]) / (sqrt(l->var[f]) + .000001f);
            for (int i = 0; i < n; ++i) x[i] *= l->scales[f];
        }
        for (int i = 0; i < n; ++i) x[i] += l->biases[f];
    }
    activate(l->output, (size_t)l->batch * l->outputs, l->act);
}

/* forward_maxpool_layer, maxpool_layer.c:79-114: invalid taps read as -FLT_MAX, strict > */
static void fwd_maxpool(olayer *l, const float *in)
{
    const int oh = l->out_h, ow = l->out_w, c = l->c;
#pragma omp parallel for
    for (int bk = 0; bk < l->batch * c; ++bk) {
        const float *src = in + (size_t)bk * l->h * l->w;
        float *dst = l->output + (size_t)bk * oh * ow;
        for (int i = 0; i < oh; ++i)
            for (int j = 0; j < ow; ++j) {
                float max = -FLT_MAX;
                for (int n = 0; n < l->size; ++n)
                    for (int m = 0; m < l->size; ++m) {
                        const int cur_h = -l->pad + i * l->stride + n;
                        const int cur_w = -l->pad + j * l->stride + m;
                        const int valid = cur_h >= 0 && cur_h < l->h && cur_w >= 0 && cur_w < l->w;
                        const float val = valid ? src[cur_w + l->w * cur_h] : -FLT_MAX;
                        max = (val > max) ? val : max;
                    }
                dst[j + ow * i] = max;
            }
    }
}

/* forward_reorg_layer (reorg_layer.c:78-85) -> reorg_cpu(x, w, h, c, batch, stride, forward=0, out)
 * (blas.c:8-29): out[in_index] = x[out_index] over the INPUT dims. */
static void fwd_reorg(olayer *l, const float *x)
{
    const int w = l->w, h = l->h, c = l->c, stride = l->stride;
    const int out_c = c / (stride * stride);
    for (int b = 0; b < l->batch; ++b)
        for (int k = 0; k < c; ++k)
            for (int j = 0; j < h; ++j)
                for (int i = 0; i < w; ++i) {
                    const int in_index = i + w * (j + h * (k + c * b));
                    const int c2 = k % out_c;
                    const int offset = k / out_c;
                    const int w2 = i * stride + offset % stride;
                    const int h2 = j * stride + offset / stride;
                    const int out_index = w2 + w * stride * (h2 + h * stride * (c2 + out_c * b));
                    l->output[in_index] = x[out_index];
                }
}

/* forward_route_layer, route_layer.c:73-86 */
static void fwd_route(onet *net, olayer *l)
{
    int offset = 0;
    for (int i = 0; i < l->n_in; ++i) {
        const float *in = net->l[l->in_layers[i]].output;
        const int sz = l->in_sizes[i];
        for (int j = 0; j < l->batch; ++j)
            memcpy(l->output + offset + (size_t)j * l->outputs, in + (size_t)j * sz, (size_t)sz * sizeof(float));
        offset += sz;
    }
}

/* forward_shortcut_layer (shortcut_layer.c:39-44) -> shortcut_cpu (blas.c:57-81) */
static void fwd_shortcut(onet *net, olayer *l, const float *in)
{
    memcpy(l->output, in, (size_t)l->batch * l->outputs * sizeof(float));
    const float *add = net->l[l->index].output;
    const int w1 = l->w, h1 = l->h, c1 = l->c, w2 = l->out_w, h2 = l->out_h, c2 = l->out_c;
    int stride = w1 / w2, sample = w2 / w1;
    if (stride < 1) stride = 1;
    if (sample < 1) sample = 1;
    const int minw = w1 < w2 ? w1 : w2, minh = h1 < h2 ? h1 : h2, minc = c1 < c2 ? c1 : c2;
    for (int b = 0; b < l->batch; ++b)
        for (int k = 0; k < minc; ++k)
            for (int j = 0; j < minh; ++j)
                for (int i = 0; i < minw; ++i) {
                    const int out_index = i * sample + w2 * (j * sample + h2 * (k + c2 * b));
                    const int add_index = i * stride + w1 * (j * stride + h1 * (k + c1 * b));
                    l->output[out_index] += add[add_index];
                }
    activate(l->output, (size_t)l->batch * l->outputs, l->act);
}

/* forward_avgpool_layer, avgpool_layer.c:40-55: sequential float sum, then /= h*w */
static void fwd_avgpool(olayer *l, const float *in)
{
    const int hw = l->h * l->w;
    for (int b = 0; b < l->batch; ++b)
        for (int k = 0; k < l->c; ++k) {
            float s = 0;
            for (int i = 0; i < hw; ++i) s += in[i + hw * (k + b * l->c)];
            s /= hw;
            l->output[k + b * l->c] = s;
        }
}

/* softmax, blas.c:205-221: float argument, double exp rounded to float, float running sum */
static void softmax_row(const float *input, int n, float temp, float *output)
{
    float sum = 0;
    float largest = -FLT_MAX;
    for (int i = 0; i < n; ++i) if (input[i] > largest) largest = input[i];
    for (int i = 0; i < n; ++i) {
        float e = exp(input[i] / temp - largest / temp);
        sum += e;
        output[i] = e;
    }
    for (int i = 0; i < n; ++i) output[i] /= sum;
}

/* softmax_tree, softmax_layer.c:35-47 (batch 1): one softmax per group, groups laid end to end */
static void softmax_groups(const float *input, float temp, const otree *t, float *output)
{
    int count = 0;
    for (int i = 0; i < t->groups; ++i) {
        softmax_row(input + count, t->group_size[i], temp, output + count);
        count += t->group_size[i];
    }
}

/* forward_softmax_layer, softmax_layer.c:49-61 */
static void fwd_softmax(olayer *l, const float *in)
{
    const int inputs = l->inputs / l->groups, batch = l->batch * l->groups;
    for (int b = 0; b < batch; ++b) {
        if (l->tree) softmax_groups(in + (size_t)b * inputs, l->temperature, l->tree, l->output + (size_t)b * inputs);
        else softmax_row(in + (size_t)b * inputs, inputs, l->temperature, l->output + (size_t)b * inputs);
    }
}

/* forward_region_layer inference path in the CPU build, region_layer.c:144-177:
 * copy, flatten(forward=1) (blas.c:31-47), logistic on objectness, softmax / softmax_tree */
static void fwd_region(olayer *l, const float *in)
{
    const int size = l->coords + l->classes + 1;
    const int hw = l->w * l->h, layers = size * l->n;
    for (int b = 0; b < l->batch; ++b)
        for (int c = 0; c < layers; ++c)
            for (int i = 0; i < hw; ++i)
                l->output[(size_t)b * layers * hw + (size_t)i * layers + c] = in[(size_t)b * layers * hw + (size_t)c * hw + i];
    for (int b = 0; b < l->batch; ++b)
        for (int i = 0; i < hw * l->n; ++i) {
            const size_t index = (size_t)size * i + (size_t)b * l->outputs;
            l->output[index + 4] = act_logistic(l->output[index + 4]);
        }
    if (l->tree) {
        for (int b = 0; b < l->batch; ++b)
            for (int i = 0; i < hw * l->n; ++i) {
                float *p = l->output + (size_t)size * i + (size_t)b * l->outputs + 5;
                softmax_groups(p, 1, l->tree, p);
            }
    } else if (l->softmax) {
        for (int b = 0; b < l->batch; ++b)
            for (int i = 0; i < hw * l->n; ++i) {
                float *p = l->output + (size_t)size * i + (size_t)b * l->outputs + 5;
                softmax_row(p, l->classes, 1, p);
            }
    }
}

/* forward_network, network.c:145-158 */
static float *forward(onet *net, const float *input)
{
    const float *in = input;
    float *out = 0;
    for (int i = 0; i < net->n; ++i) {
        olayer *l = &net->l[i];
        switch (l->type) {
        case T_CONV: fwd_conv(l, in, net->workspace); break;
        case T_MAXPOOL: fwd_maxpool(l, in); break;
        case T_REORG: fwd_reorg(l, in); break;
        case T_ROUTE: fwd_route(net, l); break;
        case T_REGION: fwd_region(l, in); break;
        case T_SHORTCUT: fwd_shortcut(net, l, in); break;
        case T_AVGPOOL: fwd_avgpool(l, in); break;
        case T_SOFTMAX: fwd_softmax(l, in); break;
        case T_COST: continue; /* get_network_output skips COST, network.c:173-181 */
        }
        in = l->output;
        out = l->output;
    }
    return out;
}

static int output_layer(const onet *net)
{
    int i;
    for (i = net->n - 1; i > 0; --i) if (net->l[i].type != T_COST) break;
    return i;
}

/* ---------------------------------------------------------------------------------------------
 * decode + NMS
 * ------------------------------------------------------------------------------------------- */
typedef struct { float x, y, w, h; } obox;

/* get_region_box, region_layer.c:73-85 (DOABS = 1) */
static obox region_box(const float *x, const float *biases, int n, int index, int i, int j, int w, int h)
{
    obox b;
    b.x = (i + act_logistic(x[index + 0])) / w;
    b.y = (j + act_logistic(x[index + 1])) / h;
    b.w = exp(x[index + 2]) * biases[2 * n] / w;
    b.h = exp(x[index + 3]) * biases[2 * n + 1] / h;
    return b;
}

/* hierarchy_predictions, tree.c:37-51 (only_leaves = 0) */
static void tree_products(float *p, int n, const otree *t)
{
    for (int j = 0; j < n; ++j) {
        const int parent = t->parent[j];
        if (parent >= 0) p[j] *= p[parent];
    }
}

/* get_region_boxes, region_layer.c:328-379, on one image's predictions */
static void region_boxes(const olayer *l, float *predictions, int w, int h, float thresh, float **probs,
                         obox *boxes, int only_objectness, const int *map)
{
    for (int i = 0; i < l->w * l->h; ++i) {
        const int row = i / l->w, col = i % l->w;
        for (int n = 0; n < l->n; ++n) {
            const int index = i * l->n + n;
            const int p_index = index * (l->classes + 5) + 4;
            float scale = predictions[p_index];
            if (l->classfix == -1 && scale < .5) scale = 0;
            const int box_index = index * (l->classes + 5);
            boxes[index] = region_box(predictions, l->biases, n, box_index, col, row, l->w, l->h);
            boxes[index].x *= w;
            boxes[index].y *= h;
            boxes[index].w *= w;
            boxes[index].h *= h;
            const int class_index = index * (l->classes + 5) + 5;
            if (l->tree) {
                tree_products(predictions + class_index, l->classes, l->tree);
                int found = 0;
                if (map) {
                    for (int j = 0; j < 200; ++j) {
                        float prob = scale * predictions[class_index + map[j]];
                        probs[index][j] = (prob > thresh) ? prob : 0;
                    }
                } else {
                    for (int j = l->classes - 1; j >= 0; --j) {
                        if (!found && predictions[class_index + j] > .5) found = 1;
                        else predictions[class_index + j] = 0;
                        float prob = predictions[class_index + j];
                        probs[index][j] = (scale > thresh) ? prob : 0;
                    }
                }
            } else {
                for (int j = 0; j < l->classes; ++j) {
                    float prob = scale * predictions[class_index + j];
                    probs[index][j] = (prob > thresh) ? prob : 0;
                }
            }
            if (only_objectness) probs[index][0] = scale;
        }
    }
}

/* box.c:67-97 */
static float overlap1d(float x1, float w1, float x2, float w2)
{
    float l1 = x1 - w1 / 2;
    float l2 = x2 - w2 / 2;
    float left = l1 > l2 ? l1 : l2;
    float r1 = x1 + w1 / 2;
    float r2 = x2 + w2 / 2;
    float right = r1 < r2 ? r1 : r2;
    return right - left;
}
static float intersection(obox a, obox b)
{
    float w = overlap1d(a.x, a.w, b.x, b.w);
    float h = overlap1d(a.y, a.h, b.y, b.h);
    if (w < 0 || h < 0) return 0;
    float area = w * h;
    return area;
}
static float iou(obox a, obox b)
{
    float i = intersection(a, b);
    float u = a.w * a.h + b.w * b.h - i;
    return i / u;
}

/* stable merge sort of box indices by probs[.][k] descending: the order nms_comparator
 * (box.c:239-247) induces under a stable sort.  cmp > 0 <=> a sorts after b. */
static int nms_cmp(float **probs, int k, int a, int b)
{
    float diff = probs[a][k] - probs[b][k];
    if (diff < 0) return 1;
    else if (diff > 0) return -1;
    return 0;
}
static void merge_sort(int *s, int *tmp, int n, float **probs, int k)
{
    if (n < 2) return;
    const int h = n / 2;
    merge_sort(s, tmp, h, probs, k);
    merge_sort(s + h, tmp, n - h, probs, k);
    int a = 0, b = h, o = 0;
    while (a < h && b < n) {
        if (nms_cmp(probs, k, s[a], s[b]) <= 0) tmp[o++] = s[a++];
        else tmp[o++] = s[b++];
    }
    while (a < h) tmp[o++] = s[a++];
    while (b < n) tmp[o++] = s[b++];
    memcpy(s, tmp, (size_t)n * sizeof(int));
}

/* do_nms_sort, box.c:249-277: the index array is built once and carried across classes */
static void nms_sort(obox *boxes, float **probs, int total, int classes, float thresh)
{
    int *s = malloc((size_t)total * sizeof(int)), *tmp = malloc((size_t)total * sizeof(int));
    for (int i = 0; i < total; ++i) s[i] = i;
    for (int k = 0; k < classes; ++k) {
        merge_sort(s, tmp, total, probs, k);
        for (int i = 0; i < total; ++i) {
            if (probs[s[i]][k] == 0) continue;
            obox a = boxes[s[i]];
            for (int j = i + 1; j < total; ++j) {
                obox b = boxes[s[j]];
                if (iou(a, b) > thresh) probs[s[j]][k] = 0;
            }
        }
    }
    free(s);
    free(tmp);
}

/* do_nms, box.c:279-297 (the unsorted variant): a row is skipped when none of its probabilities is positive at
 * the moment it is visited; for an overlapping later row the smaller entry of each class is zeroed */
static void nms_unsorted(obox *boxes, float **probs, int total, int classes, float thresh)
{
    for (int i = 0; i < total; ++i) {
        int any = 0;
        for (int k = 0; k < classes; ++k) any = any || (probs[i][k] > 0);
        if (!any) continue;
        for (int j = i + 1; j < total; ++j) {
            if (iou(boxes[i], boxes[j]) > thresh) {
                for (int k = 0; k < classes; ++k) {
                    if (probs[i][k] < probs[j][k]) probs[i][k] = 0;
                    else probs[j][k] = 0;
                }
            }
        }
    }
}

/* resize_image, image.c:1950-1993 (planar CHW, two-pass bilinear) */
static void resize_planar(const float *im, int c, int ih, int iw, int h, int w, float *out)
{
    float *part = calloc((size_t)c * ih * w, sizeof(float));
    float w_scale = (float)(iw - 1) / (w - 1);
    float h_scale = (float)(ih - 1) / (h - 1);
    for (int k = 0; k < c; ++k)
        for (int r = 0; r < ih; ++r)
            for (int x = 0; x < w; ++x) {
                float val = 0;
                if (x == w - 1 || iw == 1) val = im[(size_t)k * ih * iw + (size_t)r * iw + iw - 1];
                else {
                    float sx = x * w_scale;
                    int ix = (int)sx;
                    float dx = sx - ix;
                    val = (1 - dx) * im[(size_t)k * ih * iw + (size_t)r * iw + ix] + dx * im[(size_t)k * ih * iw + (size_t)r * iw + ix + 1];
                }
                part[(size_t)k * ih * w + (size_t)r * w + x] = val;
            }
    for (int k = 0; k < c; ++k)
        for (int r = 0; r < h; ++r) {
            float sy = r * h_scale;
            int iy = (int)sy;
            float dy = sy - iy;
            for (int x = 0; x < w; ++x) out[(size_t)k * h * w + (size_t)r * w + x] = (1 - dy) * part[(size_t)k * ih * w + (size_t)iy * w + x];
            if (r == h - 1 || ih == 1) continue;
            for (int x = 0; x < w; ++x) out[(size_t)k * h * w + (size_t)r * w + x] += dy * part[(size_t)k * ih * w + (size_t)(iy + 1) * w + x];
        }
    free(part);
}

/* ---------------------------------------------------------------------------------------------
 * command line (same dumps as oracle/ref_driver.c)
 * ------------------------------------------------------------------------------------------- */
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
    float *p = calloc(n ? n : 1, sizeof(float));
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

/* per image: get_region_boxes(l, 1, 1, thresh, probs, boxes, 0, 0) then do_nms_sort, the call
 * sequence of detector.c:494-495 */
static void decode_and_nms(onet *net, const char *outdir, float thresh, float nms, int write)
{
    olayer *l = &net->l[net->n - 1];
    if (l->type != T_REGION) return;
    const int total = l->w * l->h * l->n;
    const int *map = (getenv("Y2_USE_MAP") && atoi(getenv("Y2_USE_MAP"))) ? l->map : 0;
    obox *boxes = calloc(total, sizeof(obox));
    float **probs = calloc(total, sizeof(float *));
    for (int j = 0; j < total; ++j) probs[j] = calloc(l->classes, sizeof(float));
    float *all_boxes = calloc((size_t)l->batch * total * 4, sizeof(float));
    float *pre = calloc((size_t)l->batch * total * l->classes, sizeof(float));
    float *post = calloc((size_t)l->batch * total * l->classes, sizeof(float));
    float *dets = calloc((size_t)l->batch * total * 8, sizeof(float));
    size_t n_dets = 0;
    const int out_classes = map ? 200 : l->classes;
    for (int b = 0; b < l->batch; ++b) {
        region_boxes(l, l->output + (size_t)b * l->outputs, 1, 1, thresh, probs, boxes, 0, map);
        memcpy(all_boxes + (size_t)b * total * 4, boxes, (size_t)total * sizeof(obox));
        for (int j = 0; j < total; ++j)
            memcpy(pre + ((size_t)b * total + j) * l->classes, probs[j], (size_t)l->classes * sizeof(float));
        if (nms > 0) nms_sort(boxes, probs, total, l->classes, nms);
        for (int j = 0; j < total; ++j)
            memcpy(post + ((size_t)b * total + j) * l->classes, probs[j], (size_t)l->classes * sizeof(float));
        /* final pick, yolo_v2_class.cpp:221-227 with max_index of utils.c:533-545 (strict >, first maximum) */
        for (int j = 0; j < total; ++j) {
            int obj_id = 0;
            float prob = probs[j][0];
            for (int k = 1; k < out_classes; ++k)
                if (probs[j][k] > prob) { prob = probs[j][k]; obj_id = k; }
            if (prob > thresh) {
                float *d = dets + 8 * n_dets++;
                d[0] = (float)b; d[1] = (float)j; d[2] = (float)obj_id; d[3] = prob;
                d[4] = boxes[j].x; d[5] = boxes[j].y; d[6] = boxes[j].w; d[7] = boxes[j].h;
            }
        }
    }
    if (write) {
        write_f32(outdir, "boxes.f32", all_boxes, (size_t)l->batch * total * 4);
        write_f32(outdir, "probs_pre.f32", pre, (size_t)l->batch * total * l->classes);
        write_f32(outdir, "probs_post.f32", post, (size_t)l->batch * total * l->classes);
        write_f32(outdir, "region_after_boxes.f32", l->output, (size_t)l->batch * l->outputs);
        write_f32(outdir, "dets.f32", dets, n_dets * 8);
    }
    free(dets);
    for (int j = 0; j < total; ++j) free(probs[j]);
    free(probs); free(boxes); free(all_boxes); free(pre); free(post);
}

static int cmd_forward(int argc, char **argv)
{
    if (argc < 9) return 1;
    const char *cfg = argv[2], *weights = argv[3], *input = argv[4], *outdir = argv[5];
    const float thresh = atof(argv[6]), nms = atof(argv[7]);
    const int dump = atoi(argv[8]);
    onet net = parse_cfg(cfg);
    if (strcmp(weights, "-") != 0) load_weights(&net, weights);
    float *X = read_f32(input, (size_t)net.batch * net.inputs);
    double t0 = now_s();
    float *out = forward(&net, X);
    double t1 = now_s();
    const int oi = output_layer(&net);
    write_f32(outdir, "output.f32", out, (size_t)net.batch * net.l[oi].outputs);
    if (dump)
        for (int i = 0; i < net.n; ++i) {
            olayer *l = &net.l[i];
            if (!l->output || l->type == T_COST) continue;
            if (getenv("Y2_DUMP_MAX_MB") &&
                (double)l->batch * l->outputs * 4 > 1048576.0 * atof(getenv("Y2_DUMP_MAX_MB"))) continue;
            char name[64];
            snprintf(name, sizeof(name), "layer_%03d.f32", i);
            write_f32(outdir, name, l->output, (size_t)l->batch * l->outputs);
        }
    decode_and_nms(&net, outdir, thresh, nms, 1);
    printf("{\"batch\": %d, \"n_layers\": %d, \"outputs\": %d, \"predict_s\": %.6f}\n", net.batch, net.n,
           net.l[oi].outputs, t1 - t0);
    return 0;
}

static int cmd_region(int argc, char **argv)
{
    if (argc < 7) return 1;
    const char *cfg = argv[2], *input = argv[3], *outdir = argv[4];
    const float thresh = atof(argv[5]), nms = atof(argv[6]);
    onet net = parse_cfg(cfg);
    olayer *l = &net.l[net.n - 1];
    if (l->type != T_REGION) die("last layer is not a region layer");
    float *X = read_f32(input, (size_t)l->batch * l->inputs);
    fwd_region(l, X);
    write_f32(outdir, "region_out.f32", l->output, (size_t)l->batch * l->outputs);
    decode_and_nms(&net, outdir, thresh, nms, 1);
    printf("{\"batch\": %d, \"boxes\": %d, \"classes\": %d}\n", l->batch, l->w * l->h * l->n, l->classes);
    return 0;
}

static int cmd_time(int argc, char **argv)
{
    if (argc < 9) return 1;
    const char *cfg = argv[2], *weights = argv[3], *input = argv[4];
    const float thresh = atof(argv[5]), nms = atof(argv[6]);
    const int warmup = atoi(argv[7]), iters = atoi(argv[8]);
    onet net = parse_cfg(cfg);
    if (strcmp(weights, "-") != 0) load_weights(&net, weights);
    float *X = read_f32(input, (size_t)net.batch * net.inputs);
    for (int i = 0; i < warmup; ++i) { forward(&net, X); decode_and_nms(&net, 0, thresh, nms, 0); }
    double t0 = now_s();
    for (int i = 0; i < iters; ++i) { forward(&net, X); decode_and_nms(&net, 0, thresh, nms, 0); }
    double t1 = now_s();
    printf("{\"batch\": %d, \"iters\": %d, \"seconds\": %.6f, \"images_per_s\": %.6f}\n", net.batch, iters, t1 - t0,
           (double)net.batch * iters / (t1 - t0));
    return 0;
}

static int cmd_resize(int argc, char **argv)
{
    if (argc < 9) return 1;
    const int c = atoi(argv[3]), h = atoi(argv[4]), w = atoi(argv[5]), oh = atoi(argv[6]), ow = atoi(argv[7]);
    float *im = read_f32(argv[2], (size_t)c * h * w);
    float *out = calloc((size_t)c * oh * ow, sizeof(float));
    resize_planar(im, c, h, w, oh, ow, out);
    FILE *f = fopen(argv[8], "wb");
    if (!f) return 2;
    fwrite(out, sizeof(float), (size_t)c * oh * ow, f);
    fclose(f);
    return 0;
}

static int cmd_donms(int argc, char **argv)
{
    if (argc < 8) return 1;
    const int total = atoi(argv[4]), classes = atoi(argv[5]);
    const float thresh = atof(argv[6]);
    obox *boxes = (obox *)read_f32(argv[2], (size_t)total * 4);
    float *flat = read_f32(argv[3], (size_t)total * classes);
    float **probs = calloc(total, sizeof(float *));
    for (int j = 0; j < total; ++j) probs[j] = flat + (size_t)j * classes;
    nms_unsorted(boxes, probs, total, classes, thresh);
    FILE *f = fopen(argv[7], "wb");
    if (!f) return 2;
    fwrite(flat, sizeof(float), (size_t)total * classes, f);
    fclose(f);
    return 0;
}

static int cmd_layers(int argc, char **argv)
{
    if (argc < 3) return 1;
    /* LAYER_TYPE values of layer.h:13-38 for the types this program knows */
    static const int ref_type[] = {0 /*CONVOLUTIONAL*/, 3 /*MAXPOOL*/, 22 /*REORG*/, 8 /*ROUTE*/, 21 /*REGION*/,
                                   13 /*SHORTCUT*/, 11 /*AVGPOOL*/, 4 /*SOFTMAX*/, 9 /*COST*/};
    onet net = parse_cfg(argv[2]);
    printf("{\"batch\": %d, \"w\": %d, \"h\": %d, \"c\": %d, \"layers\": [", net.batch, net.w, net.h, net.c);
    for (int i = 0; i < net.n; ++i) {
        olayer *l = &net.l[i];
        printf("%s{\"type\": %d, \"w\": %d, \"h\": %d, \"c\": %d, \"out_w\": %d, \"out_h\": %d, \"out_c\": %d, "
               "\"outputs\": %d, \"n\": %d, \"size\": %d, \"stride\": %d, \"pad\": %d}",
               i ? ", " : "", ref_type[l->type], l->w, l->h, l->c, l->out_w, l->out_h, l->out_c, l->outputs,
               l->type == T_ROUTE ? l->n_in : l->n, l->size, l->stride, l->pad);
    }
    printf("]}\n");
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: y2_oracle forward|region|time|resize|layers ...\n"); return 1; }
    if (!strcmp(argv[1], "forward")) return cmd_forward(argc, argv);
    if (!strcmp(argv[1], "region")) return cmd_region(argc, argv);
    if (!strcmp(argv[1], "time")) return cmd_time(argc, argv);
    if (!strcmp(argv[1], "resize")) return cmd_resize(argc, argv);
    if (!strcmp(argv[1], "layers")) return cmd_layers(argc, argv);
    if (!strcmp(argv[1], "donms")) return cmd_donms(argc, argv);
    fprintf(stderr, "unknown command %s\n", argv[1]);
    return 1;
}
