/*
 * .cfg / .weights front end of the drop-in API.
 *
 * Behavioural spec: reference parser.c (read_cfg 702-735, parse_net_options 504-577,
 * parse_convolutional 139-171, parse_maxpool 359-374, parse_reorg 343-357, parse_route 450-489,
 * parse_region 236-284, parse_shortcut 415-430, parse_avgpool 376-387, parse_softmax 226-234,
 * parse_cost 309-317, parse_network_cfg 585-700, load_weights_upto 1009-1082,
 * load_convolutional_weights 963-1006, save_weights_upto 822-878).  Semantics preserved:
 * batch /= subdivisions; conv `pad=1` means padding = size/2; conv default activation is
 * logistic; maxpool size defaults to stride and padding to (size-1)/2; route indices < 0 are
 * relative; region anchors are parsed with atof at each comma; '#' and ';' start comments;
 * unknown section types print a message and leave a zeroed layer.
 */
#include "y2_host.h"

#include <assert.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    char *type;
    list *options;
} section;

typedef struct {
    int batch, inputs, h, w, c, index, time_steps;
    network net;
} size_params;

static void free_section(section *s)
{
    free(s->type);
    for (node *n = s->options->front; n; n = n->next) {
        /* kvp: key points at the line buffer, val inside it */
        char **kv = (char **)n->val;
        free(kv[0]);
        free(n->val);
    }
    free_list(s->options);
    free(s);
}

static list *read_cfg(char *filename)
{
    FILE *file = fopen(filename, "r");
    if (!file) file_error(filename);
    list *sections = make_list();
    section *current = 0;
    char *line;
    int nu = 0;
    while ((line = fgetl(file)) != 0) {
        ++nu;
        strip(line);
        switch (line[0]) {
        case '[':
            current = (section *)malloc(sizeof(section));
            list_insert(sections, current);
            current->options = make_list();
            current->type = line;
            break;
        case '\0':
        case '#':
        case ';':
            free(line);
            break;
        default:
            if (!current || !read_option(line, current->options)) {
                fprintf(stderr, "Config file error line %d, could parse: %s\n", nu, line);
                free(line);
            }
            break;
        }
    }
    fclose(file);
    return sections;
}

static LAYER_TYPE string_to_layer_type(const char *t)
{
    static const struct { const char *s; LAYER_TYPE t; } tab[] = {
        {"[shortcut]", SHORTCUT}, {"[crop]", CROP}, {"[cost]", COST}, {"[detection]", DETECTION},
        {"[region]", REGION}, {"[local]", LOCAL}, {"[conv]", CONVOLUTIONAL},
        {"[convolutional]", CONVOLUTIONAL}, {"[activation]", ACTIVE}, {"[net]", NETWORK},
        {"[network]", NETWORK}, {"[crnn]", CRNN}, {"[gru]", GRU}, {"[rnn]", RNN}, {"[conn]", CONNECTED},
        {"[connected]", CONNECTED}, {"[max]", MAXPOOL}, {"[maxpool]", MAXPOOL}, {"[reorg]", REORG},
        {"[avg]", AVGPOOL}, {"[avgpool]", AVGPOOL}, {"[dropout]", DROPOUT}, {"[lrn]", NORMALIZATION},
        {"[normalization]", NORMALIZATION}, {"[batchnorm]", BATCHNORM}, {"[soft]", SOFTMAX},
        {"[softmax]", SOFTMAX}, {"[route]", ROUTE},
    };
    for (size_t i = 0; i < sizeof(tab) / sizeof(tab[0]); ++i)
        if (strcmp(t, tab[i].s) == 0) return tab[i].t;
    return BLANK;
}

static int count_commas_plus_one(const char *s)
{
    int n = 1;
    for (; *s; ++s)
        if (*s == ',') ++n;
    return n;
}

static learning_rate_policy get_policy(char *s)
{
    static const struct { const char *s; learning_rate_policy p; } tab[] = {
        {"random", RANDOM}, {"poly", POLY}, {"constant", CONSTANT}, {"step", STEP},
        {"exp", EXP}, {"sigmoid", SIG}, {"steps", STEPS}};
    for (size_t i = 0; i < sizeof(tab) / sizeof(tab[0]); ++i)
        if (strcmp(s, tab[i].s) == 0) return tab[i].p;
    fprintf(stderr, "Couldn't find policy %s, going with constant\n", s);
    return CONSTANT;
}

static void parse_net_options(list *options, network *net)
{
    net->batch = option_find_int(options, "batch", 1);
    net->learning_rate = option_find_float(options, "learning_rate", .001);
    net->momentum = option_find_float(options, "momentum", .9);
    net->decay = option_find_float(options, "decay", .0001);
    int subdivs = option_find_int(options, "subdivisions", 1);
    net->time_steps = option_find_int_quiet(options, "time_steps", 1);
    net->batch /= subdivs;
    net->batch *= net->time_steps;
    net->subdivisions = subdivs;
    net->adam = option_find_int_quiet(options, "adam", 0);
    if (net->adam) {
        net->B1 = option_find_float(options, "B1", .9);
        net->B2 = option_find_float(options, "B2", .999);
        net->eps = option_find_float(options, "eps", .000001);
    }
    net->h = option_find_int_quiet(options, "height", 0);
    net->w = option_find_int_quiet(options, "width", 0);
    net->c = option_find_int_quiet(options, "channels", 0);
    net->inputs = option_find_int_quiet(options, "inputs", net->h * net->w * net->c);
    net->max_crop = option_find_int_quiet(options, "max_crop", net->w * 2);
    net->min_crop = option_find_int_quiet(options, "min_crop", net->w);
    net->angle = option_find_float_quiet(options, "angle", 0);
    net->aspect = option_find_float_quiet(options, "aspect", 1);
    net->saturation = option_find_float_quiet(options, "saturation", 1);
    net->exposure = option_find_float_quiet(options, "exposure", 1);
    net->hue = option_find_float_quiet(options, "hue", 0);
    if (!net->inputs && !(net->h && net->w && net->c)) error("No input parameters supplied");

    /* the schedule keys only matter for training; they are consumed so that option_unused()
     * reports the same fields as the reference */
    char *policy_s = option_find_str(options, "policy", "constant");
    net->policy = get_policy(policy_s);
    net->burn_in = option_find_int_quiet(options, "burn_in", 0);
    if (net->policy == STEP) {
        net->step = option_find_int(options, "step", 1);
        net->scale = option_find_float(options, "scale", 1);
    } else if (net->policy == STEPS) {
        char *l = option_find(options, "steps");
        char *p = option_find(options, "scales");
        if (!l || !p) error("STEPS policy must have steps and scales in cfg file");
        int n = count_commas_plus_one(l);
        net->steps = (int *)calloc(n, sizeof(int));
        net->scales = (float *)calloc(n, sizeof(float));
        for (int i = 0; i < n; ++i) {
            net->steps[i] = atoi(l);
            net->scales[i] = (float)atof(p);
            char *ln = strchr(l, ','), *pn = strchr(p, ',');
            l = ln ? ln + 1 : l + strlen(l);
            p = pn ? pn + 1 : p + strlen(p);
        }
        net->num_steps = n;
    } else if (net->policy == EXP) {
        net->gamma = option_find_float(options, "gamma", 1);
    } else if (net->policy == SIG) {
        net->gamma = option_find_float(options, "gamma", 1);
        net->step = option_find_int(options, "step", 1);
    } else if (net->policy == POLY || net->policy == RANDOM) {
        net->power = option_find_float(options, "power", 1);
    }
    net->max_batches = option_find_int(options, "max_batches", 0);
}

static layer parse_convolutional(list *options, size_params params)
{
    int n = option_find_int(options, "filters", 1);
    int size = option_find_int(options, "size", 1);
    int stride = option_find_int(options, "stride", 1);
    int pad = option_find_int_quiet(options, "pad", 0);
    int padding = option_find_int_quiet(options, "padding", 0);
    if (pad) padding = size / 2;
    char *activation_s = option_find_str(options, "activation", "logistic");
    ACTIVATION activation = get_activation(activation_s);
    if (!(params.h && params.w && params.c)) error("Layer before convolutional layer must output image.");
    int batch_normalize = option_find_int_quiet(options, "batch_normalize", 0);
    int binary = option_find_int_quiet(options, "binary", 0);
    int xnor = option_find_int_quiet(options, "xnor", 0);
    layer l = make_convolutional_layer(params.batch, params.h, params.w, params.c, n, size, stride, padding,
                                       activation, batch_normalize, binary, xnor, params.net.adam);
    l.flipped = option_find_int_quiet(options, "flipped", 0);
    l.dot = option_find_float_quiet(options, "dot", 0);
    return l;
}

static layer parse_maxpool(list *options, size_params params)
{
    int stride = option_find_int(options, "stride", 1);
    int size = option_find_int(options, "size", stride);
    int padding = option_find_int_quiet(options, "padding", (size - 1) / 2);
    if (!(params.h && params.w && params.c)) error("Layer before maxpool layer must output image.");
    return make_maxpool_layer(params.batch, params.h, params.w, params.c, size, stride, padding);
}

static layer parse_reorg(list *options, size_params params)
{
    int stride = option_find_int(options, "stride", 1);
    int reverse = option_find_int_quiet(options, "reverse", 0);
    if (!(params.h && params.w && params.c)) error("Layer before reorg layer must output image.");
    return make_reorg_layer(params.batch, params.w, params.h, params.c, stride, reverse);
}

static layer parse_route(list *options, size_params params, network net)
{
    char *l = option_find(options, "layers");
    if (!l) error("Route Layer must specify input layers");
    int n = count_commas_plus_one(l);
    int *layers = (int *)calloc(n, sizeof(int));
    int *sizes = (int *)calloc(n, sizeof(int));
    for (int i = 0; i < n; ++i) {
        int index = atoi(l);
        char *next = strchr(l, ',');
        l = next ? next + 1 : l + strlen(l);
        if (index < 0) index = params.index + index;
        if (index < 0 || index >= params.index) error("Route layer index out of range");
        layers[i] = index;
        sizes[i] = net.layers[index].outputs;
    }
    layer r = make_route_layer(params.batch, n, layers, sizes);
    layer first = net.layers[layers[0]];
    r.out_w = first.out_w;
    r.out_h = first.out_h;
    r.out_c = first.out_c;
    for (int i = 1; i < n; ++i) {
        layer next = net.layers[layers[i]];
        if (next.out_w == first.out_w && next.out_h == first.out_h) r.out_c += next.out_c;
        else r.out_h = r.out_w = r.out_c = 0;
    }
    return r; /* w/h/c stay 0 like the reference's route layer (route_layer.c:6-37) */
}

static layer parse_region(list *options, size_params params)
{
    int coords = option_find_int(options, "coords", 4);
    int classes = option_find_int(options, "classes", 20);
    int num = option_find_int(options, "num", 1);
    layer l = make_region_layer(params.batch, params.w, params.h, num, classes, coords);
    if (l.outputs != params.inputs) error("region layer: outputs != inputs of the previous layer");
    l.log = option_find_int_quiet(options, "log", 0);
    l.sqrt = option_find_int_quiet(options, "sqrt", 0);
    l.softmax = option_find_int(options, "softmax", 0);
    l.max_boxes = option_find_int_quiet(options, "max", 30);
    l.jitter = option_find_float(options, "jitter", .2);
    l.rescore = option_find_int_quiet(options, "rescore", 0);
    l.thresh = option_find_float(options, "thresh", .5);
    l.classfix = option_find_int_quiet(options, "classfix", 0);
    l.absolute = option_find_int_quiet(options, "absolute", 0);
    l.random = option_find_int_quiet(options, "random", 0);
    l.coord_scale = option_find_float(options, "coord_scale", 1);
    l.object_scale = option_find_float(options, "object_scale", 1);
    l.noobject_scale = option_find_float(options, "noobject_scale", 1);
    l.class_scale = option_find_float(options, "class_scale", 1);
    l.bias_match = option_find_int_quiet(options, "bias_match", 0);
    char *tree_file = option_find_str(options, "tree", 0);
    if (tree_file) l.softmax_tree = read_tree(tree_file);
    char *map_file = option_find_str(options, "map", 0);
    if (map_file) l.map = read_map(map_file);
    char *a = option_find_str(options, "anchors", 0);
    if (a) {
        int n = count_commas_plus_one(a);
        for (int i = 0; i < n && i < 2 * num; ++i) {
            l.biases[i] = (float)atof(a);
            char *next = strchr(a, ',');
            a = next ? next + 1 : a + strlen(a);
        }
    }
    return l;
}

static layer parse_shortcut(list *options, size_params params, network net)
{
    char *l = option_find(options, "from");
    if (!l) error("Shortcut layer must specify from");
    int index = atoi(l);
    if (index < 0) index = params.index + index;
    if (index < 0 || index >= params.index) error("Shortcut layer index out of range");
    layer from = net.layers[index];
    layer s = make_shortcut_layer(params.batch, index, params.w, params.h, params.c, from.out_w, from.out_h,
                                  from.out_c);
    char *activation_s = option_find_str(options, "activation", "linear");
    s.activation = get_activation(activation_s);
    return s;
}

static layer parse_softmax(list *options, size_params params)
{
    int groups = option_find_int_quiet(options, "groups", 1);
    layer l = make_softmax_layer(params.batch, params.inputs, groups);
    l.temperature = option_find_float_quiet(options, "temperature", 1);
    char *tree_file = option_find_str(options, "tree", 0);
    if (tree_file) l.softmax_tree = read_tree(tree_file);
    return l;
}

/* parser.c:214-224 */
static layer parse_connected(list *options, size_params params)
{
    int output = option_find_int(options, "output", 1);
    char *activation_s = option_find_str(options, "activation", "logistic");
    ACTIVATION activation = get_activation(activation_s);
    int batch_normalize = option_find_int_quiet(options, "batch_normalize", 0);
    return make_connected_layer(params.batch, params.inputs, output, activation, batch_normalize);
}

/* parser.c:389-399: the layer keeps the extent of its input */
static layer parse_dropout(list *options, size_params params)
{
    float probability = option_find_float(options, "probability", .5);
    layer l = make_dropout_layer(params.batch, params.inputs, probability);
    l.out_w = params.w;
    l.out_h = params.h;
    l.out_c = params.c;
    return l;
}

static layer parse_cost(list *options, size_params params)
{
    char *type_s = option_find_str(options, "type", "sse");
    COST_TYPE type = SSE;
    if (strcmp(type_s, "masked") == 0) type = MASKED;
    else if (strcmp(type_s, "smooth") == 0) type = SMOOTH;
    float scale = option_find_float_quiet(options, "scale", 1);
    layer l = make_cost_layer(params.batch, params.inputs, type, scale);
    l.thresh = option_find_float_quiet(options, "thresh", 0);
    return l;
}

network make_network(int n)
{
    network net;
    memset(&net, 0, sizeof(net));
    net.n = n;
    net.layers = (layer *)calloc(n > 0 ? n : 1, sizeof(layer));
    net.seen = (int *)calloc(2, sizeof(int)); /* room for the uint64 variant of the header */
    net.input_gpu = (float **)calloc(1, sizeof(float *));
    net.truth_gpu = (float **)calloc(1, sizeof(float *));
    return net;
}

network parse_network_cfg(char *filename)
{
    list *sections = read_cfg(filename);
    node *n = sections->front;
    if (!n) error("Config file has no sections");
    network net = make_network(sections->size - 1);
    net.gpu_index = gpu_index;
    size_params params;
    memset(&params, 0, sizeof(params));

    section *s = (section *)n->val;
    list *options = s->options;
    LAYER_TYPE first = string_to_layer_type(s->type);
    if (first != NETWORK) error("First section must be [net] or [network]");
    parse_net_options(options, &net);

    params.h = net.h;
    params.w = net.w;
    params.c = net.c;
    params.inputs = net.inputs;
    params.batch = net.batch;
    params.time_steps = net.time_steps;
    params.net = net;

    n = n->next;
    int count = 0;
    free_section(s);
    fprintf(stderr, "layer     filters    size              input                output\n");
    while (n) {
        params.index = count;
        fprintf(stderr, "%5d ", count);
        s = (section *)n->val;
        options = s->options;
        layer l;
        memset(&l, 0, sizeof(l));
        LAYER_TYPE lt = string_to_layer_type(s->type);
        switch (lt) {
        case CONVOLUTIONAL: l = parse_convolutional(options, params); break;
        case MAXPOOL: l = parse_maxpool(options, params); break;
        case REORG: l = parse_reorg(options, params); break;
        case ROUTE: l = parse_route(options, params, net); break;
        case REGION: l = parse_region(options, params); break;
        case SHORTCUT: l = parse_shortcut(options, params, net); break;
        case AVGPOOL: l = make_avgpool_layer(params.batch, params.w, params.h, params.c); break;
        case SOFTMAX:
            l = parse_softmax(options, params);
            net.hierarchy = l.softmax_tree;
            break;
        case COST: l = parse_cost(options, params); break;
        case CONNECTED: l = parse_connected(options, params); break;
        case DROPOUT: l = parse_dropout(options, params); break;
        case BLANK: fprintf(stderr, "Type not recognized: %s\n", s->type); break;
        default:
            /* layer types outside the detection forward path */
            fprintf(stderr, "Layer type %s is not part of the B200 hot path\n", s->type);
            l.type = lt;
            break;
        }
        l.dontload = option_find_int_quiet(options, "dontload", 0);
        l.dontloadscales = option_find_int_quiet(options, "dontloadscales", 0);
        option_unused(options);
        net.layers[count] = l;
        free_section(s);
        n = n->next;
        ++count;
        if (n) {
            params.h = l.out_h;
            params.w = l.out_w;
            params.c = l.out_c;
            params.inputs = l.outputs;
        }
    }
    free_list(sections);
    net.outputs = get_network_output_size(net);
    if (net.gpu_index >= 0) y2_plan_network(&net);
    net.output = get_network_output(net);
    return net;
}

/* ---- .weights ----------------------------------------------------------------------------- */
static void transpose_matrix(float *a, int rows, int cols)
{
    float *t = (float *)calloc((size_t)rows * cols, sizeof(float));
    for (int x = 0; x < rows; ++x)
        for (int y = 0; y < cols; ++y) t[(size_t)y * rows + x] = a[(size_t)x * cols + y];
    memcpy(a, t, (size_t)rows * cols * sizeof(float));
    free(t);
}

static void read_floats(float *dst, size_t n, FILE *fp)
{
    size_t got = fread(dst, sizeof(float), n, fp);
    (void)got; /* the reference ignores short reads (parser.c:970-991) */
}

static void load_convolutional_weights(layer *l, FILE *fp)
{
    size_t num = (size_t)l->n * l->c * l->size * l->size;
    read_floats(l->biases, l->n, fp);
    if (l->batch_normalize && !l->dontloadscales) {
        read_floats(l->scales, l->n, fp);
        read_floats(l->rolling_mean, l->n, fp);
        read_floats(l->rolling_variance, l->n, fp);
    }
    read_floats(l->weights, num, fp);
    if (l->adam) { /* parser.c:992-995 */
        read_floats(l->m, num, fp);
        read_floats(l->v, num, fp);
    }
    if (l->flipped) transpose_matrix(l->weights, l->c * l->size * l->size, l->n);
    if (gpu_index >= 0 && l->b200) y2_push_convolutional_layer(l);
}

/* parser.c:897-919: biases, weights[outputs][inputs] (transposed in very old files), then the batchnorm vectors */
static void load_connected_weights(layer *l, FILE *fp, int transpose)
{
    read_floats(l->biases, l->outputs, fp);
    read_floats(l->weights, (size_t)l->outputs * l->inputs, fp);
    if (transpose) transpose_matrix(l->weights, l->inputs, l->outputs);
    if (l->batch_normalize && !l->dontloadscales) {
        read_floats(l->scales, l->outputs, fp);
        read_floats(l->rolling_mean, l->outputs, fp);
        read_floats(l->rolling_variance, l->outputs, fp);
    }
    if (gpu_index >= 0 && l->b200) y2_push_convolutional_layer(l);
}

void load_weights_upto(network *net, char *filename, int cutoff)
{
    if (net->gpu_index >= 0) cuda_set_device(net->gpu_index);
    fprintf(stderr, "Loading weight file......");
    fflush(stdout);
    FILE *fp = fopen(filename, "rb");
    if (!fp) file_error(filename);
    int major = 0, minor = 0, revision = 0;
    if (fread(&major, sizeof(int), 1, fp) != 1 || fread(&minor, sizeof(int), 1, fp) != 1 ||
        fread(&revision, sizeof(int), 1, fp) != 1)
        error("weights file: truncated header");
    if (major * 10 + minor >= 2) {
        uint64_t seen64 = 0;
        if (fread(&seen64, sizeof(uint64_t), 1, fp) != 1) error("weights file: truncated header");
        *net->seen = (int)seen64;
    } else {
        int iseen = 0;
        if (fread(&iseen, sizeof(int), 1, fp) != 1) error("weights file: truncated header");
        *net->seen = iseen;
    }
    const int transpose = (major > 1000) || (minor > 1000); /* parser.c:1035 */
    for (int i = 0; i < net->n && i < cutoff; ++i) {
        layer *l = &net->layers[i];
        if (l->dontload) continue;
        if (l->type == CONVOLUTIONAL) load_convolutional_weights(l, fp);
        if (l->type == CONNECTED) load_connected_weights(l, fp, transpose);
    }
    fprintf(stderr, "Done!\n");
    fclose(fp);
}

void load_weights(network *net, char *filename)
{
    load_weights_upto(net, filename, net->n);
}

void save_weights_upto(network net, char *filename, int cutoff)
{
    fprintf(stderr, "Saving weights to %s\n", filename);
    FILE *fp = fopen(filename, "wb");
    if (!fp) file_error(filename);
    int major = 0, minor = 1, revision = 0;
    fwrite(&major, sizeof(int), 1, fp);
    fwrite(&minor, sizeof(int), 1, fp);
    fwrite(&revision, sizeof(int), 1, fp);
    fwrite(net.seen, sizeof(int), 1, fp);
    for (int i = 0; i < net.n && i < cutoff; ++i) {
        layer l = net.layers[i];
        if (l.type == CONNECTED) { /* parser.c:806-820 */
            fwrite(l.biases, sizeof(float), l.outputs, fp);
            fwrite(l.weights, sizeof(float), (size_t)l.outputs * l.inputs, fp);
            if (l.batch_normalize) {
                fwrite(l.scales, sizeof(float), l.outputs, fp);
                fwrite(l.rolling_mean, sizeof(float), l.outputs, fp);
                fwrite(l.rolling_variance, sizeof(float), l.outputs, fp);
            }
        }
        if (l.type != CONVOLUTIONAL) continue;
        size_t num = (size_t)l.n * l.c * l.size * l.size;
        fwrite(l.biases, sizeof(float), l.n, fp);
        if (l.batch_normalize) {
            fwrite(l.scales, sizeof(float), l.n, fp);
            fwrite(l.rolling_mean, sizeof(float), l.n, fp);
            fwrite(l.rolling_variance, sizeof(float), l.n, fp);
        }
        fwrite(l.weights, sizeof(float), num, fp);
        if (l.adam) { /* parser.c:788-791 */
            fwrite(l.m, sizeof(float), num, fp);
            fwrite(l.v, sizeof(float), num, fp);
        }
    }
    fclose(fp);
}

void save_weights(network net, char *filename)
{
    save_weights_upto(net, filename, net.n);
}
