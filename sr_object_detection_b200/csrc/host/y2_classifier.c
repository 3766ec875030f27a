/*
 * predict_classifier: the classifier front end over the GPU forward pass.
 *
 * Reference interface replaced (behavioural spec only): classifier.c:676-730 predict_classifier - load_image_color,
 * letterbox_image to the network size, network_predict, hierarchy_predictions for a WordTree classifier
 * (net.hierarchy), top_k, one "name: probability" line per hit.  Same data-cfg keys (names / labels / top), same
 * output lines; images are read as binary PPM / PGM (the reference decodes with stb_image).  The timing line goes to
 * stderr so that stdout carries exactly the prediction lines.
 */
#include "y2_host.h"

#include <stdlib.h>
#include <string.h>
#include <time.h>

/* the prediction lines of classifier.c:717-722 for one image, written to `out` */
void predict_classifier_image(network net, image im, char **names, int top, FILE *out)
{
    image r = letterbox_image(im, net.w, net.h);
    fprintf(out, "%d %d\n", r.w, r.h);
    float *predictions = network_predict(net, r.data);
    if (net.hierarchy) hierarchy_predictions(predictions, net.outputs, net.hierarchy, 0);
    int *indexes = (int *)calloc(top, sizeof(int));
    top_k(predictions, net.outputs, top, indexes);
    for (int i = 0; i < top; ++i) {
        const int index = indexes[i];
        if (net.hierarchy)
            fprintf(out, "%d, %s: %f, parent: %s \n", index, names[index], predictions[index],
                    (net.hierarchy->parent[index] >= 0) ? names[net.hierarchy->parent[index]] : "Root");
        else fprintf(out, "%s: %f\n", names[index], predictions[index]);
    }
    free(indexes);
    if (r.data != im.data) free_image(r);
}

void predict_classifier(char *datacfg, char *cfgfile, char *weightfile, char *filename, int top)
{
    network net = parse_network_cfg(cfgfile);
    if (weightfile) load_weights(&net, weightfile);
    set_batch_network(&net, 1);
    list *options = read_data_cfg(datacfg);
    char *name_list = option_find_str(options, "names", 0);
    if (!name_list) name_list = option_find_str(options, "labels", "data/labels.list");
    if (top == 0) top = option_find_int(options, "top", 1);
    char **names = get_labels(name_list);
    char buff[256];
    char *input = buff;
    while (1) {
        if (filename) {
            strncpy(input, filename, 255);
            input[255] = 0;
        } else {
            printf("Enter Image Path: ");
            fflush(stdout);
            input = fgets(input, 256, stdin);
            if (!input) break;
            strtok(input, "\n");
        }
        image im = load_image_color(input, 0, 0);
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        predict_classifier_image(net, im, names, top, stdout);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        fprintf(stderr, "%s: Predicted in %f seconds.\n", input,
                (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec));
        free_image(im);
        if (filename) break;
    }
    fflush(stdout);
    free_network(net);
}
