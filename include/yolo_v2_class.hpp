/*
 * C++ detector facade of the drop-in API: the public surface of the reference's
 * yolo_v2_class.hpp (class Detector, bbox_t, image_t; yolo_v2_class.hpp:27-146) over
 * libyolo2_b200.so.  Same names, argument meaning, defaults and exceptions; the implementation
 * (csrc/host/y2_detector.cpp) runs the B200 forward pass, decode and NMS on the device.
 */
#ifndef YOLO_V2_CLASS_B200_HPP
#define YOLO_V2_CLASS_B200_HPP

#include <deque>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef OPENCV
#include <opencv2/opencv.hpp>
#endif

#if defined(_MSC_VER)
#ifdef YOLODLL_EXPORTS
#define YOLODLL_API __declspec(dllexport)
#else
#define YOLODLL_API __declspec(dllimport)
#endif
#else
#define YOLODLL_API __attribute__((visibility("default")))
#endif

/* top-left corner + extent in pixels of the image handed to detect(), as the reference reports them */
struct bbox_t {
    unsigned int x, y, w, h;
    float prob;            /* class probability of the winning class */
    unsigned int obj_id;   /* winning class, [0, classes) */
    unsigned int track_id; /* 0 = untracked; tracking() hands out ids >= 1 per class */
};

/* planar CHW floats in [0,1], RGB */
struct image_t {
    int h, w, c;
    float *data;
};

class Detector {
    std::shared_ptr<void> detector_gpu_ptr; /* opaque state, as in the reference */

public:
    float nms = .4f;

    YOLODLL_API Detector(std::string cfg_filename, std::string weight_filename, int gpu_id = 0);
    YOLODLL_API ~Detector();

    YOLODLL_API std::vector<bbox_t> detect(std::string image_filename, float thresh = 0.2f, bool use_mean = false);
    YOLODLL_API std::vector<bbox_t> detect(image_t img, float thresh = 0.2f, bool use_mean = false);
    static YOLODLL_API image_t load_image(std::string image_filename);
    static YOLODLL_API void free_image(image_t m);
    YOLODLL_API int get_net_width() const;
    YOLODLL_API int get_net_height() const;

    YOLODLL_API std::vector<bbox_t> tracking(std::vector<bbox_t> cur_bbox_vec, int const frames_story = 6);

    /* B200 extension: raw decoded frame, uint8 interleaved RGB of any size; the byte -> float
     * conversion of load_image and resize_image (image.c:1950-1993) happen on the device, with the same
     * result as detect(image_filename) gives for that frame */
    YOLODLL_API std::vector<bbox_t> detect_rgb8(const unsigned char *rgb, int w, int h, float thresh = 0.2f);

#ifdef OPENCV
    std::vector<bbox_t> detect(cv::Mat mat, float thresh = 0.2f, bool use_mean = false)
    {
        if (mat.data == NULL) throw std::runtime_error("Image is empty");
        std::shared_ptr<image_t> resized = mat_to_image_resize(mat);
        return detect_resized(*resized, mat.size(), thresh, use_mean);
    }

    /* boxes of a network-sized image scaled back to the frame it was resized from */
    std::vector<bbox_t> detect_resized(image_t img, cv::Size init_size, float thresh = 0.2f, bool use_mean = false)
    {
        if (img.data == NULL) throw std::runtime_error("Image is empty");
        std::vector<bbox_t> found = detect(img, thresh, use_mean);
        const float kx = (float)init_size.width / img.w, ky = (float)init_size.height / img.h;
        for (bbox_t &b : found) {
            b.x *= kx;
            b.w *= kx;
            b.y *= ky;
            b.h *= ky;
        }
        return found;
    }

    std::shared_ptr<image_t> mat_to_image_resize(cv::Mat mat) const
    {
        if (mat.data == NULL) return std::shared_ptr<image_t>();
        cv::Mat net_sized;
        cv::resize(mat, net_sized, cv::Size(get_net_width(), get_net_height()));
        return mat_to_image(net_sized);
    }

    /* BGR uint8 cv::Mat -> planar RGB float image, value = byte / 255. */
    static std::shared_ptr<image_t> mat_to_image(cv::Mat bgr)
    {
        std::shared_ptr<image_t> out(new image_t, [](image_t *p) {
            free_image(*p);
            delete p;
        });
        const int h = bgr.rows, w = bgr.cols, c = bgr.channels();
        out->h = h;
        out->w = w;
        out->c = c;
        out->data = (float *)calloc((size_t)h * w * c, sizeof(float));
        for (int k = 0; k < c; ++k) {
            const int src_k = (c == 3) ? 2 - k : k; /* BGR -> RGB */
            for (int y = 0; y < h; ++y) {
                const unsigned char *row = bgr.ptr<unsigned char>(y);
                for (int x = 0; x < w; ++x) out->data[((size_t)k * h + y) * w + x] = row[x * c + src_k] / 255.;
            }
        }
        return out;
    }
#endif /* OPENCV */

    std::deque<std::vector<bbox_t>> prev_bbox_vec_deque;
};

#endif
