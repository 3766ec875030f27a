/* TEST INFRASTRUCTURE: runs the reference's own predict_classifier (classifier.c:676-730, compiled from the reference
 * sources by oracle/Makefile, target refcls) on its CPU path; tests/golden/make_golden.py stores the lines it prints.
 *   ref_classify <data.cfg> <net.cfg> <net.weights> <image> <top> */
#include <stdio.h>
#include <stdlib.h>
void predict_classifier(char *datacfg, char *cfgfile, char *weightfile, char *filename, int top);
extern int gpu_index;
void *GlobleObjBoxes; /* darknet.c:358-359, not part of the CPU objects */
int GlobleObjBoxesNum;
int main(int argc, char **argv)
{
    if (argc < 6) { fprintf(stderr, "usage: ref_classify data.cfg net.cfg net.weights image top\n"); return 1; }
    gpu_index = -1;
    predict_classifier(argv[1], argv[2], argv[3], argv[4], atoi(argv[5]));
    return 0;
}
