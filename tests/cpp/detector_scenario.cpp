// One caller, two libraries: this file uses nothing but the PUBLIC surface of the reference's
// yolo_v2_class.hpp (class Detector, bbox_t, image_t) and is compiled twice -
//   * against the reference's own header + yolo_v2_class.cpp + CPU objects (oracle/Makefile, target refdet:
//     oracle/_ref/detector_ref), where it writes the golden lines of tests/golden/detector_ref.json;
//   * against include/yolo_v2_class.hpp + libyolo2_b200.so (tests/test_detector_cpp.py), where its output must
//     reproduce those lines character for character.
// The scenario network is exactly representable (synth.exact_detector_*): its head output is bit-identical in
// fp32 on the CPU and in bf16 x bf16 -> fp32 on the tensor cores, so everything behind it - region forward,
// get_region_boxes, do_nms_sort, the final pick, the pixel conversion, the 3-frame mean and tracking() - has to
// agree to the last bit.
//
//   detector_scenario <cfg> <weights> <frames.f32> <n_frames> <w> <h> <thresh> <nms> <story> <image.ppm> <gpu_id>
// n_frames = 0 runs the scripted tracking() part only (pure host logic: works with gpu_id = -1 on a CPU box).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "yolo_v2_class.hpp"

#ifdef Y2_REFERENCE_BUILD
extern "C" {
void *GlobleObjBoxes;  // defined in the reference's darknet.c:358-359, which is not part of the CPU objects
int GlobleObjBoxesNum;
}
#endif

static void dump(const char *tag, int i, const std::vector<bbox_t> &v)
{
    printf("%s %d %zu", tag, i, v.size());
    for (const bbox_t &b : v) printf(" %u %u %u %u %.9g %u %u", b.x, b.y, b.w, b.h, b.prob, b.obj_id, b.track_id);
    printf("\n");
}

static bbox_t mk(unsigned x, unsigned y, unsigned w, unsigned h, unsigned obj)
{
    bbox_t b;
    b.x = x; b.y = y; b.w = w; b.h = h; b.prob = 0.5f; b.obj_id = obj; b.track_id = 0;
    return b;
}

int main(int argc, char **argv)
{
    if (argc < 12) {
        fprintf(stderr, "usage: detector_scenario cfg weights frames.f32 n w h thresh nms story image.ppm gpu_id\n");
        return 1;
    }
    const int n = atoi(argv[4]), w = atoi(argv[5]), h = atoi(argv[6]), story = atoi(argv[9]), gpu = atoi(argv[11]);
    const float thresh = (float)atof(argv[7]), nms = (float)atof(argv[8]);
    std::vector<float> frames((size_t)n * 3 * w * h);
    if (n > 0) {
        FILE *f = fopen(argv[3], "rb");
        if (!f || fread(frames.data(), 4, frames.size(), f) != frames.size()) return 2;
        fclose(f);
    }
  if (n > 0) {
    Detector det(argv[1], argv[2], gpu);
    det.nms = nms;
    printf("size %d %d\n", det.get_net_width(), det.get_net_height());
    image_t im;
    im.w = w; im.h = h; im.c = 3;
    for (int i = 0; i < n; ++i) {  // plain detect, tracked across the sequence
        im.data = frames.data() + (size_t)i * 3 * w * h;
        std::vector<bbox_t> v = det.detect(im, thresh, false);
        dump("detect", i, v);
        dump("track", i, det.tracking(v, story));
    }
    for (int i = 0; i < n; ++i) {  // the 3-frame mean of the network output (yolo_v2_class.cpp:208-213)
        im.data = frames.data() + (size_t)i * 3 * w * h;
        dump("mean", i, det.detect(im, thresh, true));
    }
    dump("file", 0, det.detect(std::string(argv[10]), thresh, false));
    image_t loaded = Detector::load_image(argv[10]);
    printf("loaded %d %d %d %.9g %.9g\n", loaded.w, loaded.h, loaded.c, loaded.data[0],
           loaded.data[(size_t)loaded.w * loaded.h * loaded.c - 1]);
    Detector::free_image(loaded);
    det.nms = 0;  // `if (nms) do_nms_sort(...)`: no suppression at all
    im.data = frames.data();
    dump("nonms", 0, det.detect(im, thresh, false));
    try {
        Detector::load_image("/nonexistent/file.ppm");
        printf("load no-throw\n");
    } catch (const std::runtime_error &e) {
        printf("load %s\n", e.what());
    }
  }

    // tracking() on a scripted sequence with a fresh detector (ids restart at 1 per class)
    Detector det2(argv[1], argv[2], gpu);
    std::vector<std::vector<bbox_t>> seq = {
        {mk(10, 10, 40, 40, 0), mk(200, 200, 50, 50, 0), mk(300, 20, 30, 60, 1)},
        {mk(14, 12, 44, 40, 0), mk(205, 190, 50, 54, 0), mk(500, 400, 30, 30, 1)},
        {},
        {mk(20, 15, 40, 40, 0), mk(290, 30, 30, 60, 1), mk(295, 28, 30, 60, 1)},
        {mk(400, 400, 10, 10, 2)},
        {mk(22, 18, 40, 40, 0), mk(402, 398, 12, 12, 2), mk(60, 60, 40, 40, 0)},
    };
    for (size_t i = 0; i < seq.size(); ++i) dump("script", (int)i, det2.tracking(seq[i], 3));
    return 0;
}
