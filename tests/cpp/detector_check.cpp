// Test driver for the C++ Detector facade (tests/test_detector_cpp.py compiles and runs it).
//   detector_check track                       -> tracking() on a scripted sequence (no GPU needed)
//   detector_check detect cfg weights in.f32 in.u8 w h thresh -> detect / use_mean / detect_rgb8 on cuda:0
#include "yolo_v2_class.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static void dump(const char *tag, const std::vector<bbox_t> &v)
{
    printf("%s %zu", tag, v.size());
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
    if (argc >= 4 && !strcmp(argv[1], "track")) {
        Detector det(argv[2], argv[3], -1);  // host-only description: tracking is pure host logic
        std::vector<std::vector<bbox_t>> frames = {
            {mk(10, 10, 40, 40, 0), mk(200, 200, 50, 50, 0), mk(300, 20, 30, 60, 1)},
            {mk(14, 12, 44, 40, 0), mk(205, 190, 50, 54, 0), mk(500, 400, 30, 30, 1)},
            {},
            {mk(20, 15, 40, 40, 0), mk(290, 30, 30, 60, 1), mk(295, 28, 30, 60, 1)},
            {mk(400, 400, 10, 10, 2)},
        };
        for (auto &f : frames) dump("track", det.tracking(f, 3));
        return 0;
    }
    if (argc >= 9 && !strcmp(argv[1], "detect")) {
        const int w = atoi(argv[6]), h = atoi(argv[7]);
        const float thresh = (float)atof(argv[8]);
        std::vector<float> img((size_t)3 * w * h);
        std::vector<unsigned char> u8((size_t)3 * w * h);
        FILE *f = fopen(argv[4], "rb");
        if (!f || fread(img.data(), 4, img.size(), f) != img.size()) return 2;
        fclose(f);
        f = fopen(argv[5], "rb");
        if (!f || fread(u8.data(), 1, u8.size(), f) != u8.size()) return 2;
        fclose(f);
        Detector det(argv[2], argv[3], 0);
        printf("size %d %d\n", det.get_net_width(), det.get_net_height());
        image_t im;
        im.w = w; im.h = h; im.c = 3; im.data = img.data();
        dump("detect", det.detect(im, thresh, false));
        for (int i = 0; i < 3; ++i) dump("mean", det.detect(im, thresh, true));
        dump("rgb8", det.detect_rgb8(u8.data(), w, h, thresh));  // any frame size: resize runs on the device
        try {
            Detector::load_image("/nonexistent/file.ppm");
            printf("load no-throw\n");
        } catch (const std::runtime_error &e) {
            printf("load %s\n", e.what());
        }
        return 0;
    }
    fprintf(stderr, "usage: detector_check track cfg weights | detect cfg weights in.f32 in.u8 w h thresh\n");
    return 1;
}
