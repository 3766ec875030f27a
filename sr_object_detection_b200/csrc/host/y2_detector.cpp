/*
 * class Detector over the B200 library: behavioural restatement of yolo_v2_class.cpp:37-304
 * (constructor, detect with the optional 3-frame mean, load_image, tracking).
 *
 * Differences, all inside the contract: detect() without use_mean keeps region decode, NMS and the
 * final pick on the device (network_detect_batch) instead of pulling 845 x 20 probabilities to the
 * host - the kernels are bit-exact against get_region_boxes / do_nms_sort / max_index, so the boxes
 * are the same; load_image reads binary PPM / PGM (the reference uses the third-party stb_image, which
 * is not part of this repository).
 */
#include "yolo_v2_class.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>

extern "C" {
#include "darknet_b200.h"
}

namespace {

constexpr int kMeanFrames = 3; /* FRAMES of yolo_v2_class.cpp:22 */

struct State {
    network net;
    int total = 0, classes = 0, outputs = 0;
    /* use_mean path: the reference's host-side flow */
    box *boxes = nullptr;
    float **probs = nullptr;
    float *avg = nullptr;
    float *frames[kMeanFrames] = {nullptr, nullptr, nullptr};
    int frame_index = 0;
    /* device path */
    std::vector<y2_detection> dets;
    std::vector<unsigned int> next_track_id; /* per class, starts at 1 */
};

State &state_of(const std::shared_ptr<void> &p) { return *static_cast<State *>(p.get()); }

struct DeviceScope { /* run on the detector's GPU, restore the caller's selection afterwards */
    int saved, device;
    explicit DeviceScope(int dev) : saved(gpu_index), device(dev)
    {
        if (device >= 0) cuda_set_device(device);
    }
    ~DeviceScope()
    {
        if (device >= 0 && saved >= 0 && saved != device) cuda_set_device(saved);
        gpu_index = saved;
    }
};

/* yolo_v2_class.cpp:228-236: corner in double, clamped at 0, truncated to unsigned; extent float -> unsigned */
bbox_t to_bbox(float bx, float by, float bw, float bh, float prob, int obj_id, int img_w, int img_h)
{
    bbox_t b;
    b.x = (unsigned int)std::max((double)0, (bx - bw / 2.) * img_w);
    b.y = (unsigned int)std::max((double)0, (by - bh / 2.) * img_h);
    b.w = (unsigned int)(bw * img_w);
    b.h = (unsigned int)(bh * img_h);
    b.obj_id = (unsigned int)obj_id;
    b.prob = prob;
    b.track_id = 0;
    return b;
}

} // namespace

Detector::Detector(std::string cfg_filename, std::string weight_filename, int gpu_id)
{
    State *st = new State();
    detector_gpu_ptr = std::shared_ptr<void>(st, [](void *p) { delete static_cast<State *>(p); });
    const int caller_device = gpu_index;
    if (gpu_id >= 0) cuda_set_device(gpu_id);
    else gpu_index = gpu_id; /* host-only description: detect() is unavailable, tracking() works */
    st->net = parse_network_cfg(const_cast<char *>(cfg_filename.c_str()));
    if (!weight_filename.empty()) load_weights(&st->net, const_cast<char *>(weight_filename.c_str()));
    set_batch_network(&st->net, 1);
    st->net.gpu_index = gpu_id;
    const layer &l = st->net.layers[st->net.n - 1];
    st->total = l.w * l.h * l.n;
    st->classes = l.classes;
    st->outputs = l.outputs;
    st->avg = (float *)calloc(st->outputs, sizeof(float));
    for (float *&f : st->frames) f = (float *)calloc(st->outputs, sizeof(float));
    st->boxes = (box *)calloc(st->total > 0 ? st->total : 1, sizeof(box));
    st->probs = (float **)calloc(st->total > 0 ? st->total : 1, sizeof(float *));
    for (int j = 0; j < st->total; ++j) st->probs[j] = (float *)calloc(st->classes, sizeof(float));
    st->dets.resize(st->total > 0 ? st->total : 1);
    st->next_track_id.assign(st->classes > 0 ? st->classes : 1, 1u);
    /* hand the caller's device selection back (only if a device was touched at all) */
    if (gpu_id >= 0 && caller_device >= 0 && caller_device != gpu_id) cuda_set_device(caller_device);
    gpu_index = caller_device;
}

Detector::~Detector()
{
    State &st = state_of(detector_gpu_ptr);
    for (int j = 0; j < st.total; ++j) free(st.probs[j]);
    free(st.probs);
    free(st.boxes);
    free(st.avg);
    for (float *f : st.frames) free(f);
    DeviceScope scope(st.net.gpu_index);
    free_network(st.net);
}

int Detector::get_net_width() const { return state_of(detector_gpu_ptr).net.w; }
int Detector::get_net_height() const { return state_of(detector_gpu_ptr).net.h; }

std::vector<bbox_t> Detector::detect(std::string image_filename, float thresh, bool use_mean)
{
    image_t img = load_image(image_filename);
    std::shared_ptr<float> guard(img.data, [](float *p) { free(p); });
    return detect(img, thresh, use_mean);
}

/* binary PPM (P6) / PGM (P5), maxval 255 -> planar floats, value = byte / 255. (the conversion of
 * load_image_stb, yolo_v2_class.cpp:129-149); grey images are replicated to three planes */
image_t Detector::load_image(std::string image_filename)
{
    FILE *f = fopen(image_filename.c_str(), "rb");
    if (!f) throw std::runtime_error("file not found");
    auto token = [&](int &v) -> bool {
        int ch = fgetc(f);
        for (;;) {
            while (ch == ' ' || ch == '\n' || ch == '\r' || ch == '\t') ch = fgetc(f);
            if (ch != '#') break;
            while (ch != '\n' && ch != EOF) ch = fgetc(f);
        }
        if (ch < '0' || ch > '9') return false;
        v = 0;
        while (ch >= '0' && ch <= '9') {
            v = v * 10 + (ch - '0');
            ch = fgetc(f);
        }
        return true;
    };
    char magic[2] = {0, 0};
    int w = 0, h = 0, maxval = 0;
    const bool header_ok = fread(magic, 1, 2, f) == 2 && magic[0] == 'P' && (magic[1] == '5' || magic[1] == '6') &&
                           token(w) && token(h) && token(maxval) && w > 0 && h > 0 && maxval == 255;
    if (!header_ok) {
        fclose(f);
        throw std::runtime_error("file not found");
    }
    const int src_c = magic[1] == '6' ? 3 : 1;
    std::vector<unsigned char> bytes((size_t)w * h * src_c);
    const size_t got = fread(bytes.data(), 1, bytes.size(), f);
    fclose(f);
    if (got != bytes.size()) throw std::runtime_error("file not found");
    image_t img;
    img.w = w;
    img.h = h;
    img.c = 3;
    img.data = (float *)calloc((size_t)w * h * 3, sizeof(float));
    for (int k = 0; k < 3; ++k)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x)
                img.data[((size_t)k * h + y) * w + x] =
                    (float)bytes[((size_t)y * w + x) * src_c + (src_c == 3 ? k : 0)] / 255.;
    return img;
}

void Detector::free_image(image_t m)
{
    if (m.data) free(m.data);
}

std::vector<bbox_t> Detector::detect(image_t img, float thresh, bool use_mean)
{
    State &st = state_of(detector_gpu_ptr);
    network &net = st.net;
    if (!img.data) throw std::runtime_error("Image is empty");
    DeviceScope scope(net.gpu_index);
    image in;
    in.w = img.w;
    in.h = img.h;
    in.c = img.c;
    in.data = img.data;
    image sized;
    const bool resized = !(net.w == in.w && net.h == in.h);
    if (resized) sized = resize_image(in, net.w, net.h); /* image.c:1950-1993 */
    else sized = in;
    std::vector<bbox_t> found;
    if (!use_mean) {
        int count = 0;
        network_detect_batch(net, sized.data, thresh, nms, st.dets.data(), &count, st.total);
        if (count > st.total) count = st.total;
        found.reserve(count);
        for (int i = 0; i < count; ++i) {
            const y2_detection &d = st.dets[i];
            found.push_back(to_bbox(d.x, d.y, d.w, d.h, d.prob, d.obj_id, img.w, img.h));
        }
    } else {
        /* yolo_v2_class.cpp:206-216: average the last three network outputs on the host, decode the mean */
        float *prediction = network_predict(net, sized.data);
        layer l = net.layers[net.n - 1];
        memcpy(st.frames[st.frame_index], prediction, (size_t)l.outputs * sizeof(float));
        mean_arrays(st.frames, kMeanFrames, l.outputs, st.avg);
        l.output = st.avg;
        st.frame_index = (st.frame_index + 1) % kMeanFrames;
        get_region_boxes(l, 1, 1, thresh, st.probs, st.boxes, 0, 0);
        if (nms) do_nms_sort(st.boxes, st.probs, st.total, l.classes, nms);
        for (int i = 0; i < st.total; ++i) {
            const int obj_id = max_index(st.probs[i], l.classes);
            const float prob = st.probs[i][obj_id];
            if (prob > thresh) {
                const box &b = st.boxes[i];
                found.push_back(to_bbox(b.x, b.y, b.w, b.h, prob, obj_id, img.w, img.h));
            }
        }
    }
    if (resized) ::free_image(sized);
    return found;
}

std::vector<bbox_t> Detector::detect_rgb8(const unsigned char *rgb, int w, int h, float thresh)
{
    State &st = state_of(detector_gpu_ptr);
    network &net = st.net;
    if (!rgb) throw std::runtime_error("Image is empty");
    if (w <= 0 || h <= 0) throw std::runtime_error("Image is empty");
    DeviceScope scope(net.gpu_index);
    int count = 0;
    /* any frame size: byte/255. and resize_image run on the device, as detect(filename) does on the host */
    network_detect_batch_frames(net, rgb, w, h, thresh, nms, st.dets.data(), &count, st.total);
    if (count > st.total) count = st.total;
    std::vector<bbox_t> found;
    found.reserve(count);
    for (int i = 0; i < count; ++i) {
        const y2_detection &d = st.dets[i];
        found.push_back(to_bbox(d.x, d.y, d.w, d.h, d.prob, d.obj_id, w, h));
    }
    return found;
}

/* yolo_v2_class.cpp:251-304.  Boxes of the current frame inherit the track id of the nearest box of the
 * same class among the last `frames_story` frames (centre distance < 100 px, a closer claim wins, an id
 * is never given twice in a frame) and average their extent with it; the rest get fresh per-class ids. */
std::vector<bbox_t> Detector::tracking(std::vector<bbox_t> cur, int const frames_story)
{
    State &st = state_of(detector_gpu_ptr);
    auto fresh_id = [&](unsigned int obj_id) -> unsigned int {
        if (obj_id >= st.next_track_id.size()) st.next_track_id.resize(obj_id + 1, 1u);
        return st.next_track_id[obj_id]++;
    };
    auto remember = [&]() {
        prev_bbox_vec_deque.push_front(cur);
        if (prev_bbox_vec_deque.size() > (size_t)frames_story) prev_bbox_vec_deque.pop_back();
    };
    bool history = false;
    for (const auto &frame : prev_bbox_vec_deque) history = history || !frame.empty();
    if (!history) {
        for (bbox_t &b : cur) b.track_id = fresh_id(b.obj_id);
        remember();
        return cur;
    }
    std::vector<unsigned int> best(cur.size(), std::numeric_limits<unsigned int>::max());
    for (const auto &frame : prev_bbox_vec_deque) {
        for (const bbox_t &old : frame) {
            int match = -1;
            for (size_t m = 0; m < cur.size(); ++m) {
                const bbox_t &k = cur[m];
                if (old.obj_id != k.obj_id) continue;
                const float dx = (float)(old.x + old.w / 2) - (float)(k.x + k.w / 2);
                const float dy = (float)(old.y + old.h / 2) - (float)(k.y + k.h / 2);
                const unsigned int dist = (unsigned int)std::sqrt(dx * dx + dy * dy);
                if (dist < 100 && (k.track_id == 0 || best[m] > dist)) {
                    best[m] = dist;
                    match = (int)m;
                }
            }
            const bool id_taken = std::any_of(cur.begin(), cur.end(), [&](const bbox_t &b) {
                return b.track_id == old.track_id && b.obj_id == old.obj_id;
            });
            if (match >= 0 && !id_taken) {
                cur[match].track_id = old.track_id;
                cur[match].w = (cur[match].w + old.w) / 2;
                cur[match].h = (cur[match].h + old.h) / 2;
            }
        }
    }
    for (bbox_t &b : cur)
        if (b.track_id == 0) b.track_id = fresh_id(b.obj_id);
    remember();
    return cur;
}
