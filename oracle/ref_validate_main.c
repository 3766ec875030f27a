/* TEST INFRASTRUCTURE: runs the reference's own validate_detector (detector.c:244-369, compiled from the reference
 * sources by oracle/Makefile, target refval) on its CPU path.  tests/golden/make_golden.py stores the result files
 * it writes; the GPU tests compare ours byte for byte.   ref_validate [recall] <data.cfg> <net.cfg> <net.weights>
 * (`recall`: validate_detector_recall, whose per-image lines on stderr are the golden) */
#include <stdio.h>
#include <string.h>
void validate_detector(char *datacfg, char *cfgfile, char *weightfile);
void validate_detector_recall(char *datacfg, char *cfgfile, char *weightfile); /* detector.c:371-450 */
extern int gpu_index;
void *GlobleObjBoxes; /* darknet.c:358-359, not part of the CPU objects */
int GlobleObjBoxesNum;
int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: ref_validate data.cfg net.cfg net.weights\n"); return 1; }
    gpu_index = -1;
    if (argc >= 5 && !strcmp(argv[1], "recall")) { /* ref_validate recall <data.cfg> <net.cfg> <net.weights> */
        validate_detector_recall(argv[2], argv[3], argv[4]);
        return 0;
    }
    validate_detector(argv[1], argv[2], argv[3]);
    return 0;
}
