// gpuimageproc::GpuStereoProcessor re-hosted on libb200stereo.so (include/b200_stereo.h).
//
// Same public method names, argument meaning and error behaviour as the reference class
// (reference: include/gpuimageproc/GPUStereoProcessor.h:63-126, src/GPUStereoProcessor.cpp).  The reference takes
// cv::Mat / sensor_msgs types; neither OpenCV's C++ headers nor ROS exist in the build image, so this header carries a
// minimal `Mat` (rows, cols, OpenCV type code, contiguous bytes).  With -DB200S_WITH_OPENCV the cv::Mat overloads of the
// reference signatures (uploadMat, downloadMat, rectifyImageLeft/Right, computeDisparity, computeDisparityBare,
// filterSpeckles, printStats) are compiled in as thin adapters.  The reference aborts on errors (assert /
// cv::Exception); here every failure throws gpuimageproc::Error carrying the b200s_error code and message.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../b200_stereo.h"
#ifdef B200S_WITH_OPENCV
#include <opencv2/core.hpp>
#endif

namespace gpuimageproc
{

// verbatim id space of the reference (GPUStereoProcessor.h:21-57)
enum GpuMatSource
{
    GPU_MAT_SIDE_L = 1 << 0, GPU_MAT_SIDE_R = 1 << 1, GPU_MAT_SIDE_MASK = 3,
    GPU_MAT_SRC_RAW = 1 << 2, GPU_MAT_SRC_MONO = 1 << 3, GPU_MAT_SRC_COLOR = 1 << 4, GPU_MAT_SRC_RECT_MONO = 1 << 5,
    GPU_MAT_SRC_RECT_COLOR = 1 << 6, GPU_MAT_SRC_DISPARITY = 1 << 7, GPU_MAT_SRC_DISPARITY_32F = 1 << 8,
    GPU_MAT_SRC_DISPARITY_IMG = 1 << 9, GPU_MAT_SRC_POINTS2 = 1 << 10,
    GPU_MAT_SRC_L_RAW = GPU_MAT_SRC_RAW | GPU_MAT_SIDE_L, GPU_MAT_SRC_R_RAW = GPU_MAT_SRC_RAW | GPU_MAT_SIDE_R,
    GPU_MAT_SRC_L_MONO = GPU_MAT_SRC_MONO | GPU_MAT_SIDE_L, GPU_MAT_SRC_R_MONO = GPU_MAT_SRC_MONO | GPU_MAT_SIDE_R,
    GPU_MAT_SRC_L_COLOR = GPU_MAT_SRC_COLOR | GPU_MAT_SIDE_L, GPU_MAT_SRC_R_COLOR = GPU_MAT_SRC_COLOR | GPU_MAT_SIDE_R,
    GPU_MAT_SRC_L_RECT_MONO = GPU_MAT_SRC_RECT_MONO | GPU_MAT_SIDE_L, GPU_MAT_SRC_R_RECT_MONO = GPU_MAT_SRC_RECT_MONO | GPU_MAT_SIDE_R,
    GPU_MAT_SRC_L_RECT_COLOR = GPU_MAT_SRC_RECT_COLOR | GPU_MAT_SIDE_L, GPU_MAT_SRC_R_RECT_COLOR = GPU_MAT_SRC_RECT_COLOR | GPU_MAT_SIDE_R,
    GPU_MAT_SRC_L_DISPARITY = GPU_MAT_SRC_DISPARITY | GPU_MAT_SIDE_L, GPU_MAT_SRC_R_DISPARITY = GPU_MAT_SRC_DISPARITY | GPU_MAT_SIDE_R,
    GPU_MAT_SRC_L_DISPARITY_32F = GPU_MAT_SRC_DISPARITY_32F | GPU_MAT_SIDE_L, GPU_MAT_SRC_R_DISPARITY_32F = GPU_MAT_SRC_DISPARITY_32F | GPU_MAT_SIDE_R,
    GPU_MAT_SRC_L_DISPARITY_IMG = GPU_MAT_SRC_DISPARITY_IMG | GPU_MAT_SIDE_L, GPU_MAT_SRC_R_DISPARITY_IMG = GPU_MAT_SRC_DISPARITY_IMG | GPU_MAT_SIDE_R,
    GPU_MAT_SRC_L_POINTS2 = GPU_MAT_SRC_POINTS2 | GPU_MAT_SIDE_L, GPU_MAT_SRC_R_POINTS2 = GPU_MAT_SRC_POINTS2 | GPU_MAT_SIDE_R
};
inline GpuMatSource operator|(GpuMatSource a, GpuMatSource b) { return static_cast<GpuMatSource>(static_cast<int>(a) | static_cast<int>(b)); }
inline GpuMatSource operator&(GpuMatSource a, GpuMatSource b) { return static_cast<GpuMatSource>(static_cast<int>(a) & static_cast<int>(b)); }

struct Error : std::runtime_error
{
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

// minimal stand-in for cv::Mat: tightly packed rows, OpenCV type codes (B200S_8UC1 ...)
struct Mat
{
    int rows = 0, cols = 0, type = B200S_8UC1;
    std::vector<uint8_t> data;
    static int elemSize(int t) { return t == B200S_8UC1 ? 1 : t == B200S_16SC1 ? 2 : t == B200S_32FC1 ? 4 : t == B200S_8UC3 ? 3 : t == B200S_32FC3 ? 12 : 4; }
    Mat() {}
    Mat(int r, int c, int t) : rows(r), cols(c), type(t), data((size_t)r * c * elemSize(t)) {}
    size_t step() const { return (size_t)cols * elemSize(type); }
    template <typename T> T *ptr(int r = 0) { return reinterpret_cast<T *>(data.data() + (size_t)r * step()); }
    template <typename T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(data.data() + (size_t)r * step()); }
};

struct CameraInfo   // sensor_msgs::CameraInfo subset (K, D, R, P, size)
{
    int width = 0, height = 0;
    double K[9] = {0}, R[9] = {0}, P[12] = {0};
    std::vector<double> D;
};

// Senders (src/GpuSenderIfc.cpp, src/GpuSender{Image,Disparity,Pc2}.cpp): each owns the page-locked host buffer its
// message payload is packed into.  enqueueSend* only enqueues work on the side's stream; once the payload is complete the
// CUDA runtime's callback thread calls the sender's publisher -- the reference's
// cv::cuda::Stream::enqueueHostCallback(GPUSender::callback) (src/GpuSenderIfc.cpp:13-26) -- and wasDataSent() turns true.
struct PinnedBuffer
{
    void *p = nullptr;
    size_t bytes = 0;
    explicit PinnedBuffer(size_t n) : bytes(n)
    {
        if (b200s_host_alloc(&p, n ? n : 1) != B200S_OK) throw Error(B200S_ENOMEM, "cudaHostAlloc failed");
    }
    ~PinnedBuffer() { b200s_host_free(p); }
    PinnedBuffer(const PinnedBuffer &) = delete;
    PinnedBuffer &operator=(const PinnedBuffer &) = delete;
};

class GPUSenderIfc
{
  public:
    virtual ~GPUSenderIfc() {}
    bool wasDataSent() const { return sent_.load(std::memory_order_acquire); }
    const uint8_t *data() const { return static_cast<const uint8_t *>(buf_->p); }
    size_t size() const { return buf_->bytes; }

  protected:
    explicit GPUSenderIfc(size_t bytes) : buf_(new PinnedBuffer(bytes)) {}
    virtual void publish() = 0;
    static void callback(void *self, int /*status*/)
    {
        GPUSenderIfc *s = static_cast<GPUSenderIfc *>(self);
        s->publish();
        s->sent_.store(true, std::memory_order_release);
    }
    std::unique_ptr<PinnedBuffer> buf_;
    std::atomic<bool> sent_{false};
    friend class GpuStereoProcessor;
};

struct ImagePayload { int height = 0, width = 0, step = 0; std::string encoding; const uint8_t *data = nullptr; size_t size = 0; };
struct DisparityPayload { b200s_disparity_meta meta{}; const float *data = nullptr; size_t count = 0; };
struct PointCloud2Payload { b200s_pc2_meta meta{}; const uint8_t *data = nullptr; size_t size = 0; };

template <typename Payload>
class GPUSender : public GPUSenderIfc
{
  public:
    typedef std::function<void(const Payload &)> Publisher;     // stands for ros::Publisher::publish
    GPUSender(size_t bytes, Publisher pub) : GPUSenderIfc(bytes), pub_(pub) {}
    Payload payload;

  protected:
    void publish() override
    {
        if (pub_) pub_(payload);
    }
    Publisher pub_;
};
typedef GPUSender<ImagePayload> GPUSenderImage;
typedef GPUSender<DisparityPayload> GPUSenderDisparity;
typedef GPUSender<PointCloud2Payload> GPUSenderPc2;
typedef std::shared_ptr<GPUSenderIfc> GPUSenderIfcPtr;
typedef std::shared_ptr<GPUSenderImage> GPUSenderImagePtr;
typedef std::shared_ptr<GPUSenderDisparity> GPUSenderDisparityPtr;
typedef std::shared_ptr<GPUSenderPc2> GPUSenderPc2Ptr;

class GpuStereoProcessor
{
  public:
    explicit GpuStereoProcessor(int device = 0)
    {
        int rc = b200s_create(device, &h_);
        if (rc) throw Error(rc, "b200s_create failed: no usable CUDA device (there is no CPU fallback)");
        ck(b200s_get_params(h_, &p_));
    }
    ~GpuStereoProcessor()
    {
        b200s_wait(h_, 0);      // pending sender callbacks reference buffers owned by senders_
        senders_.clear();
        b200s_destroy(h_);
    }
    GpuStereoProcessor(const GpuStereoProcessor &) = delete;
    GpuStereoProcessor &operator=(const GpuStereoProcessor &) = delete;

    void initStereoModel(const CameraInfo &l, const CameraInfo &r)
    {
        b200s_caminfo a = conv(l), b = conv(r);
        ck(b200s_set_calibration(h_, &a, &b));
    }
    void initStereoModel(const std::string &left_cal_file, const std::string &right_cal_file) { ck(b200s_load_calibration_files(h_, left_cal_file.c_str(), right_cal_file.c_str())); }
    bool isStereoModelInitialised() { return b200s_is_model_initialised(h_) != 0; }
    void convertRawToColor(GpuMatSource side) { ck(b200s_convert_raw_to_color(h_, side)); }
    void convertRawToMono(GpuMatSource side) { ck(b200s_convert_raw_to_mono(h_, side)); }
    void convertColor(GpuMatSource mat_source, GpuMatSource mat_dst, const std::string &src_encoding, const std::string &dst_encoding)
    {
        ck(b200s_convert_color(h_, mat_source, mat_dst, src_encoding.c_str(), dst_encoding.c_str()));
    }
    void uploadMat(GpuMatSource mat_source, const Mat &m, std::string encoding = "") { ck(b200s_upload(h_, mat_source, m.data.data(), m.rows, m.cols, m.type, m.step(), encoding.c_str())); }
    void downloadMat(GpuMatSource mat_source, Mat &m)
    {
        int r, c, t;
        ck(b200s_mat_info(h_, mat_source, &r, &c, &t));
        if (m.rows != r || m.cols != c || m.type != t) m = Mat(r, c, t);
        ck(b200s_download(h_, mat_source, m.data.data(), m.step()));
    }
    void rectifyImage(GpuMatSource source, GpuMatSource dest, int interpolation = B200S_INTER_LINEAR) { ck(b200s_rectify(h_, source, dest, interpolation)); }
    void rectifyImageLeft(const Mat &source, Mat &dest, int interpolation = B200S_INTER_LINEAR) { rectifySide(GPU_MAT_SIDE_L, source, dest, interpolation); }
    void rectifyImageRight(const Mat &source, Mat &dest, int interpolation = B200S_INTER_LINEAR) { rectifySide(GPU_MAT_SIDE_R, source, dest, interpolation); }
    void computeDisparity(GpuMatSource left, GpuMatSource right, GpuMatSource disparity)
    {
        sync();
        ck(b200s_compute_disparity(h_, left, right, disparity));
    }
    // cv::cuda::StereoBM compatibility mode: the CV_8UC1 plane the reference's GPU matcher writes (src/GPUStereoProcessor.cpp:283)
    void computeDisparityCudaCompat(GpuMatSource left, GpuMatSource right, GpuMatSource disparity)
    {
        sync();
        ck(b200s_compute_disparity_cuda_compat(h_, left, right, disparity));
    }
    // computeDisparityBare: matcher only on host images -> CV_16SC1 x16 (src/GPUStereoProcessor.cpp:305-310)
    void computeDisparityBare(const Mat &left, const Mat &right, Mat &disparity)
    {
        uploadMat(GPU_MAT_SRC_L_RECT_MONO, left);
        uploadMat(GPU_MAT_SRC_R_RECT_MONO, right);
        computeDisparity(GPU_MAT_SRC_L_RECT_MONO, GPU_MAT_SRC_R_RECT_MONO, GPU_MAT_SRC_L_DISPARITY);
        downloadMat(GPU_MAT_SRC_L_DISPARITY, disparity);
    }
    // computeDisparity(cv::Mat&, cv::Mat&, cv::Mat&): CV_32F = d/16 - (cx_l - cx_r)  (src/GPUStereoProcessor.cpp:312-321)
    void computeDisparity(const Mat &left, const Mat &right, Mat &disparity)
    {
        Mat d16;
        computeDisparityBare(left, right, d16);
        downloadMat(GPU_MAT_SRC_L_DISPARITY_32F, disparity);
    }
    void computeDisparityImage(GpuMatSource disparity_src, GpuMatSource disp_image_dest) { sync(); ck(b200s_compute_disparity_image(h_, disparity_src, disp_image_dest)); }
    void projectDisparityTo3DPoints(GpuMatSource disparity_src, GpuMatSource points_src) { ck(b200s_project_to_3d(h_, disparity_src, points_src)); }
    void waitForStream(GpuMatSource stream_source) { ck(b200s_wait(h_, stream_source & GPU_MAT_SIDE_MASK)); }
    void waitForAllStreams() { ck(b200s_wait(h_, 0)); }
    // src/GPUStereoProcessor.cpp:228-234: drop the senders whose data went out
    void cleanSenders()
    {
        std::vector<GPUSenderIfcPtr> keep;
        for (auto &s : senders_)
            if (!s->wasDataSent()) keep.push_back(s);
        senders_.swap(keep);
    }
    // printStats (src/GPUStereoProcessor.cpp:421-435) of a named buffer, reduced on the GPU
    void printStats(const std::string &name, GpuMatSource mat)
    {
        double mn[4], mx[4], mean[4];
        int ch = 0;
        ck(b200s_mat_stats(h_, mat, mn, mx, mean, &ch));
        for (int i = 0; i < ch; ++i) std::printf("ARRAY STATS:%s; channel:%d; min:%f; max:%f; mean:%f;\n", name.c_str(), i, mn[i], mx[i], mean[i]);
    }
    void printStats(const std::string &name, const Mat &mat)
    {
        uploadMat(GPU_MAT_SRC_R_DISPARITY_IMG, mat);      // a scratch id the chain does not use
        printStats(name, GPU_MAT_SRC_R_DISPARITY_IMG);
    }
    void filterSpeckles(GpuMatSource disparity_src) { sync(); ck(b200s_filter_speckles(h_, disparity_src)); }
    void filterSpeckles(Mat &disparity)   // CV_16SC1 in place, newVal = FILTERED
    {
        ck(b200s_filter_speckles_host(h_, disparity.ptr<int16_t>(), disparity.rows, disparity.cols, disparity.step(),
                                      (p_.min_disparity - 1) * 16, p_.speckle_window_size, p_.speckle_range));
    }
    // senders (src/GPUStereoProcessor.cpp:210-226): asynchronous like the reference -- the call returns once the work is
    // enqueued; `pub` runs on the stream-callback thread when the payload is complete.  waitForStream / waitForAllStreams
    // (or polling wasDataSent) tell when the message may be read.
    GPUSenderImagePtr enqueueSendImage(GpuMatSource source, const std::string &encoding, GPUSenderImage::Publisher pub = nullptr)
    {
        int r, c, t, rows, cols, step;
        ck(b200s_mat_info(h_, source, &r, &c, &t));
        auto s = std::make_shared<GPUSenderImage>((size_t)r * c * Mat::elemSize(t), pub);
        senders_.push_back(s);
        ck(b200s_pack_image_async(h_, source, s->buf_->p, s->buf_->bytes, &rows, &cols, &step, &GPUSenderIfc::callback, s.get()));
        s->payload.height = rows; s->payload.width = cols; s->payload.step = step; s->payload.encoding = encoding;
        s->payload.data = s->data(); s->payload.size = s->size();
        return s;
    }
    GPUSenderDisparityPtr enqueueSendDisparity(GpuMatSource source, GPUSenderDisparity::Publisher pub = nullptr)
    {
        int r, c, t;
        ck(b200s_mat_info(h_, planeOf(source), &r, &c, &t));
        auto s = std::make_shared<GPUSenderDisparity>((size_t)r * c * 4, pub);
        senders_.push_back(s);
        sync();
        s->payload.data = reinterpret_cast<const float *>(s->data()); s->payload.count = (size_t)r * c;
        ck(b200s_pack_disparity_async(h_, source, s->buf_->p, s->buf_->bytes, &s->payload.meta, &GPUSenderIfc::callback, s.get()));
        return s;
    }
    // points_source: the reference passes GPU_MAT_SRC_L_POINTS2 (src/StereoProcessor.cpp:281, test/UTest.cpp:382); reprojection
    // and packing are one kernel here, so that id (or the DISPARITY id) names the side whose fixed-point plane is read
    GPUSenderPc2Ptr enqueueSendPoints(GpuMatSource points_source, GpuMatSource color_source, GPUSenderPc2::Publisher pub = nullptr)
    {
        int r, c, t;
        ck(b200s_mat_info(h_, planeOf(points_source), &r, &c, &t));
        auto s = std::make_shared<GPUSenderPc2>((size_t)r * c * 32, pub);
        senders_.push_back(s);
        s->payload.data = s->data(); s->payload.size = s->size();
        ck(b200s_pack_pointcloud2_async(h_, points_source, color_source, s->buf_->p, s->buf_->bytes, &s->payload.meta,
                                        &GPUSenderIfc::callback, s.get()));
        return s;
    }
    // parameters (src/GPUStereoProcessor.cpp:202-208,389-419 plus the cv::StereoBM ones GPU.cfg lacked)
    void setPreFilterType(int filter_type) { p_.pre_filter_type = filter_type; }
    void setPreFilterSize(int v) { p_.pre_filter_size = v; }
    void setPreFilterCap(int v) { p_.pre_filter_cap = v; }
    void setRefineDisparity(bool ref_disp) { p_.refine_disparity = ref_disp; }
    void setBlockSize(int block_size) { p_.block_size = block_size; }
    void setNumDisparities(int numDisp) { p_.num_disparities = numDisp; }
    void setMinDisparity(int minDisp) { p_.min_disparity = minDisp; }
    void setTextureThreshold(int threshold) { p_.texture_threshold = threshold; }
    void setUniquenessRatio(int v) { p_.uniqueness_ratio = v; }
    void setDisp12MaxDiff(int v) { p_.disp12_max_diff = v; }
    int getMaxSpeckleSize() const { return p_.speckle_window_size; }
    void setMaxSpeckleSize(int maxSpeckleSize) { p_.speckle_window_size = maxSpeckleSize; }
    double getMaxSpeckleDiff() const { return p_.speckle_range / 16.0; }
    void setMaxSpeckleDiff(double maxSpeckleDiff) { p_.speckle_range = (int)(maxSpeckleDiff * 16 + 0.5); }   // integer-disparity units
    void setSpeckleRange(int raw_x16) { p_.speckle_range = raw_x16; }
    b200s_handle *handle() { return h_; }

#ifdef B200S_WITH_OPENCV
    // ---- cv::Mat overloads with the reference's exact signatures (GPUStereoProcessor.h:73-96) ----------------------
    void uploadMat(GpuMatSource mat_source, const cv::Mat &cv_mat, std::string encoding = "")
    {
        ck(b200s_upload(h_, mat_source, cv_mat.data, cv_mat.rows, cv_mat.cols, cv_mat.type(), cv_mat.step, encoding.c_str()));
        ck(b200s_wait(h_, mat_source & GPU_MAT_SIDE_MASK));     // pageable cv::Mat memory may be released by the caller
    }
    void downloadMat(GpuMatSource mat_source, cv::Mat &cv_mat)
    {
        int r, c, t;
        ck(b200s_mat_info(h_, mat_source, &r, &c, &t));
        cv_mat.create(r, c, t);
        ck(b200s_download(h_, mat_source, cv_mat.data, cv_mat.step));
    }
    void rectifyImageLeft(const cv::Mat &source, cv::Mat &dest, int interpolation = B200S_INTER_LINEAR) { rectifySideCv(GPU_MAT_SIDE_L, source, dest, interpolation); }
    void rectifyImageRight(const cv::Mat &source, cv::Mat &dest, int interpolation = B200S_INTER_LINEAR) { rectifySideCv(GPU_MAT_SIDE_R, source, dest, interpolation); }
    void computeDisparityBare(const cv::Mat &left, const cv::Mat &right, cv::Mat &disparity)
    {
        uploadMat(GPU_MAT_SRC_L_RECT_MONO, left);
        uploadMat(GPU_MAT_SRC_R_RECT_MONO, right);
        computeDisparity(GPU_MAT_SRC_L_RECT_MONO, GPU_MAT_SRC_R_RECT_MONO, GPU_MAT_SRC_L_DISPARITY);
        downloadMat(GPU_MAT_SRC_L_DISPARITY, disparity);
    }
    void computeDisparity(cv::Mat &left, cv::Mat &right, cv::Mat &disparity)
    {
        cv::Mat d16;
        computeDisparityBare(left, right, d16);
        downloadMat(GPU_MAT_SRC_L_DISPARITY_32F, disparity);
    }
    void filterSpeckles(cv::Mat &disparity)     // CV_16SC1 in place, newVal = FILTERED (GPUStereoProcessor.cpp:367-385)
    {
        ck(b200s_filter_speckles_host(h_, reinterpret_cast<int16_t *>(disparity.data), disparity.rows, disparity.cols, disparity.step,
                                      (p_.min_disparity - 1) * 16, p_.speckle_window_size, p_.speckle_range));
    }
    void printStats(const std::string &name, cv::Mat &mat)
    {
        uploadMat(GPU_MAT_SRC_R_DISPARITY_IMG, mat);
        printStats(name, GPU_MAT_SRC_R_DISPARITY_IMG);
    }
#endif

  private:
    void ck(int rc)
    {
        if (rc) throw Error(rc, b200s_last_error_string(h_));
    }
    void sync() { ck(b200s_set_params(h_, &p_)); }
    static b200s_caminfo conv(const CameraInfo &c)
    {
        b200s_caminfo o;
        std::memset(&o, 0, sizeof(o));
        o.width = c.width; o.height = c.height;
        std::memcpy(o.K, c.K, sizeof(o.K)); std::memcpy(o.R, c.R, sizeof(o.R)); std::memcpy(o.P, c.P, sizeof(o.P));
        o.n_D = (int)(c.D.size() > 8 ? 8 : c.D.size());
        for (int i = 0; i < o.n_D; ++i) o.D[i] = c.D[i];
        return o;
    }
    void rectifySide(GpuMatSource side, const Mat &src, Mat &dst, int interp)
    {
        GpuMatSource raw = GPU_MAT_SRC_RAW | side, out = (src.type == B200S_8UC1 ? GPU_MAT_SRC_RECT_MONO : GPU_MAT_SRC_RECT_COLOR) | side;
        uploadMat(raw, src);
        rectifyImage(raw, out, interp);
        downloadMat(out, dst);
    }
#ifdef B200S_WITH_OPENCV
    void rectifySideCv(GpuMatSource side, const cv::Mat &src, cv::Mat &dst, int interp)
    {
        GpuMatSource raw = GPU_MAT_SRC_RAW | side, out = (src.channels() == 1 ? GPU_MAT_SRC_RECT_MONO : GPU_MAT_SRC_RECT_COLOR) | side;
        uploadMat(raw, src);
        rectifyImage(raw, out, interp);
        downloadMat(out, dst);
    }
#endif
    static GpuMatSource planeOf(GpuMatSource id)
    {
        return (id & (GPU_MAT_SRC_POINTS2 | GPU_MAT_SRC_DISPARITY_32F)) ? (GPU_MAT_SRC_DISPARITY | (id & GPU_MAT_SIDE_MASK)) : id;
    }
    b200s_handle *h_ = nullptr;
    b200s_params p_{};
    std::vector<GPUSenderIfcPtr> senders_;
};

}  // namespace gpuimageproc
