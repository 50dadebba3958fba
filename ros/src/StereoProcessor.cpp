// Topic plumbing around gpuimageproc::GpuStereoProcessor for the B200 build.
// Behaviour kept from the reference (src/StereoProcessor.cpp): parameters and synchroniser choice of the constructor
// (:20-101), lazy (un)subscription driven by the publishers' subscriber counts (:104-145), the per-frame sequence of
// imageCb (:157-298: upload -> mono/colour conversion -> rectify -> disparity + speckle filter -> disparity_vis ->
// reproject + PointCloud2, every product only when somebody listens) and the dynamic_reconfigure mapping of configCb
// (:307-336).  Deliberate differences: `disparity_min` is wired to setMinDisparity (the reference passes disparity_range,
// SURVEY.md bug B2) and the cv::StereoBM parameters the extended cfg/GPU.cfg adds are forwarded too.
#include "gpuimageproc/StereoProcessor.h"

#include <boost/bind.hpp>
#include <sensor_msgs/image_encodings.h>

#include <chrono>
#include <cstring>

namespace gpuimageproc
{

const std::string StereoProcessor::CAMERA_TOPIC_LEFT = "left";
const std::string StereoProcessor::CAMERA_TOPIC_RIGHT = "right";
const std::string StereoProcessor::CAMERA_TOPIC_IMAGE = "/image_raw";
const std::string StereoProcessor::CAMERA_TOPIC_INFO = "/camera_info";

namespace
{

// The published surface of the reference (src/StereoProcessor.cpp:89-100), as data: bit i of ConnectedTopics <-> entry i.
enum TopicKind { KIND_IMAGE, KIND_DISPARITY, KIND_CLOUD };
struct TopicDesc
{
    const char *name;
    TopicKind kind;
};
const TopicDesc kTopics[] = {
    {"left/image_mono", KIND_IMAGE},  {"right/image_mono", KIND_IMAGE},  {"left/image_color", KIND_IMAGE}, {"right/image_color", KIND_IMAGE},
    {"left/rect_mono", KIND_IMAGE},   {"right/rect_mono", KIND_IMAGE},   {"left/rect_color", KIND_IMAGE},  {"right/rect_color", KIND_IMAGE},
    {"disparity", KIND_DISPARITY},    {"disparity_vis", KIND_IMAGE},     {"pointcloud", KIND_CLOUD}};

CameraInfo toCameraInfo(const sensor_msgs::CameraInfo &m)
{
    CameraInfo c;
    c.width = (int)m.width;
    c.height = (int)m.height;
    for (int i = 0; i < 9; ++i) { c.K[i] = m.K[i]; c.R[i] = m.R[i]; }
    for (int i = 0; i < 12; ++i) c.P[i] = m.P[i];
    c.D.assign(m.D.begin(), m.D.end());
    return c;
}

double msSince(const std::chrono::steady_clock::time_point &t0)
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace

StereoProcessor::StereoProcessor(ros::NodeHandle &nh_in, ros::NodeHandle &private_nh_in)
    : nh(nh_in), private_nh(private_nh_in), camera_info_from_files_(false)
{
    int device = 0;
    private_nh.param("cuda_device", device, 0);
    stereoProcessor_.reset(new GpuStereoProcessor(device));
    it_.reset(new image_transport::ImageTransport(nh));

    int queue_size = 5;
    bool approx = false;
    private_nh.param("queue_size", queue_size, 5);
    private_nh.param("approximate_sync", approx, false);
    private_nh.param<std::string>("camera_info_file_left", camera_info_file_left_, std::string());
    private_nh.param<std::string>("camera_info_file_right", camera_info_file_right_, std::string());
    camera_info_from_files_ = !camera_info_file_left_.empty() || !camera_info_file_right_.empty();
    ROS_INFO("PARAM: camera_info_file_left:%s camera_info_file_right:%s queue_size:%d approximate_sync:%s", camera_info_file_left_.c_str(),
             camera_info_file_right_.c_str(), queue_size, approx ? "true" : "false");

    // With calibration files only the two images are synchronised; otherwise images and camera infos of both sides.
    if (camera_info_from_files_) {
        if (approx) {
            approximate_sync_images_.reset(new message_filters::Synchronizer<ApproxImages>(ApproxImages(queue_size), sub_l_raw_image_, sub_r_raw_image_));
            approximate_sync_images_->registerCallback(boost::bind(&StereoProcessor::imageCb, this, _1, _2));
        } else {
            exact_sync_images_.reset(new message_filters::Synchronizer<ExactImages>(ExactImages(queue_size), sub_l_raw_image_, sub_r_raw_image_));
            exact_sync_images_->registerCallback(boost::bind(&StereoProcessor::imageCb, this, _1, _2));
        }
    } else {
        if (approx) {
            approximate_sync_images_and_info_.reset(new message_filters::Synchronizer<ApproxImagesAndInfo>(
                ApproxImagesAndInfo(queue_size), sub_l_raw_image_, sub_l_info_, sub_r_raw_image_, sub_r_info_));
            approximate_sync_images_and_info_->registerCallback(boost::bind(&StereoProcessor::imageAndInfoCb, this, _1, _2, _3, _4));
        } else {
            exact_sync_images_and_info_.reset(new message_filters::Synchronizer<ExactImagesAndInfo>(
                ExactImagesAndInfo(queue_size), sub_l_raw_image_, sub_l_info_, sub_r_raw_image_, sub_r_info_));
            exact_sync_images_and_info_->registerCallback(boost::bind(&StereoProcessor::imageAndInfoCb, this, _1, _2, _3, _4));
        }
    }

    reconfigure_server_.reset(new ReconfigureServer(config_mutex_, private_nh));
    ReconfigureServer::CallbackType f = boost::bind(&StereoProcessor::configCb, this, _1, _2);
    reconfigure_server_->setCallback(f);

    // Publishers; subscriptions to the cameras happen on demand in connectCb.  The lock keeps connectCb out until every
    // publisher is assigned.
    ros::SubscriberStatusCallback connect_cb = boost::bind(&StereoProcessor::connectCb, this);
    boost::lock_guard<boost::mutex> lock(connect_mutex_);
    int depth = 1;
    private_nh.param("publisher_queue_size", depth, 1);
    for (int i = 0; i < N_TOPICS; ++i) {
        switch (kTopics[i].kind) {
        case KIND_DISPARITY: publishers_[i] = private_nh.advertise<stereo_msgs::DisparityImage>(kTopics[i].name, depth, connect_cb, connect_cb); break;
        case KIND_CLOUD: publishers_[i] = private_nh.advertise<sensor_msgs::PointCloud2>(kTopics[i].name, depth, connect_cb, connect_cb); break;
        default: publishers_[i] = private_nh.advertise<sensor_msgs::Image>(kTopics[i].name, depth, connect_cb, connect_cb); break;
        }
    }
}

void StereoProcessor::connectCb()
{
    boost::lock_guard<boost::mutex> lock(connect_mutex_);
    for (int i = 0; i < N_TOPICS; ++i)
        connected_.set(static_cast<ConnectedTopics::Topic>(1u << i), publishers_[i].getNumSubscribers() > 0);
    if (!connected_.any()) {
        ROS_INFO("Un-subscribing from images and camera infos");
        sub_l_raw_image_.unsubscribe();
        sub_l_info_.unsubscribe();
        sub_r_raw_image_.unsubscribe();
        sub_r_info_.unsubscribe();
    } else if (!sub_l_raw_image_.getSubscriber()) {
        // queue size 1 on the transport side; the synchroniser's queue is the one that matters
        image_transport::TransportHints hints("raw", ros::TransportHints(), private_nh);
        ROS_INFO("Subscribing to raw images");
        sub_l_raw_image_.subscribe(*it_, CAMERA_TOPIC_LEFT + CAMERA_TOPIC_IMAGE, 1, hints);
        sub_r_raw_image_.subscribe(*it_, CAMERA_TOPIC_RIGHT + CAMERA_TOPIC_IMAGE, 1, hints);
        if (!camera_info_from_files_) {
            ROS_INFO("Subscribing to camera infos");
            sub_l_info_.subscribe(nh, CAMERA_TOPIC_LEFT + CAMERA_TOPIC_INFO, 1);
            sub_r_info_.subscribe(nh, CAMERA_TOPIC_RIGHT + CAMERA_TOPIC_INFO, 1);
        }
    }
}

void StereoProcessor::imageAndInfoCb(const sensor_msgs::ImageConstPtr &l_raw_msg, const sensor_msgs::CameraInfoConstPtr &l_info_msg,
                                     const sensor_msgs::ImageConstPtr &r_raw_msg, const sensor_msgs::CameraInfoConstPtr &r_info_msg)
{
    if (!stereoProcessor_->isStereoModelInitialised() && !camera_info_from_files_)
        stereoProcessor_->initStereoModel(toCameraInfo(*l_info_msg), toCameraInfo(*r_info_msg));
    imageCb(l_raw_msg, r_raw_msg);
}

// sensor_msgs/Image -> named raw buffer.  8-bit 1- and 3-channel encodings are on the hot path; the others are rejected
// by convertRawToMono / convertRawToColor with an error naming the encoding (no CPU fallback, SURVEY.md 8 out of scope).
void StereoProcessor::uploadRaw(GpuMatSource id, const sensor_msgs::ImageConstPtr &msg)
{
    const int ch = sensor_msgs::image_encodings::numChannels(msg->encoding);
    const int bits = sensor_msgs::image_encodings::bitDepth(msg->encoding);
    if (bits != 8 || (ch != 1 && ch != 3)) throw Error(B200S_EUNSUPPORTED, "raw encoding '" + msg->encoding + "' is not supported by the B200 path");
    Mat view;                       // the facade's Mat owns its bytes; one host copy like cv_bridge::toCvShare + upload
    view.rows = (int)msg->height;
    view.cols = (int)msg->width;
    view.type = ch == 1 ? B200S_8UC1 : B200S_8UC3;
    view.data.resize((size_t)view.rows * view.cols * ch);
    for (int y = 0; y < view.rows; ++y) std::memcpy(&view.data[(size_t)y * view.cols * ch], &msg->data[(size_t)y * msg->step], (size_t)view.cols * ch);
    stereoProcessor_->uploadMat(id, view, msg->encoding);
    stereoProcessor_->waitForStream(id);      // `view` dies at the end of this scope
}

void StereoProcessor::sendImage(GpuMatSource source, const sensor_msgs::ImageConstPtr &pattern, const std::string &encoding, ros::Publisher *pub)
{
    const std_msgs::Header header = pattern->header;
    stereoProcessor_->enqueueSendImage(source, encoding, [header, pub](const ImagePayload &p) {
        sensor_msgs::ImagePtr msg(new sensor_msgs::Image);
        msg->header = header;
        msg->encoding = p.encoding;
        msg->height = p.height;
        msg->width = p.width;
        msg->step = p.step;               // width * bitdepth * channels / 8 (src/GpuSenderImage.cpp:20)
        msg->data.assign(p.data, p.data + p.size);
        pub->publish(msg);
    });
}

void StereoProcessor::sendDisparity(GpuMatSource source, const sensor_msgs::ImageConstPtr &pattern, ros::Publisher *pub)
{
    const std_msgs::Header header = pattern->header;
    stereoProcessor_->enqueueSendDisparity(source, [header, pub](const DisparityPayload &p) {
        stereo_msgs::DisparityImagePtr msg(new stereo_msgs::DisparityImage);
        msg->header = header;
        msg->image.header = header;
        msg->image.encoding = sensor_msgs::image_encodings::TYPE_32FC1;
        msg->image.height = p.meta.height;
        msg->image.width = p.meta.width;
        msg->image.step = p.meta.step;
        msg->image.data.resize(p.count * sizeof(float));
        std::memcpy(&msg->image.data[0], p.data, p.count * sizeof(float));
        msg->f = p.meta.f;
        msg->T = p.meta.T;
        msg->min_disparity = p.meta.min_disparity;
        msg->max_disparity = p.meta.max_disparity;
        msg->delta_d = p.meta.delta_d;
        msg->valid_window.x_offset = p.meta.valid_x_offset;
        msg->valid_window.y_offset = p.meta.valid_y_offset;
        msg->valid_window.width = p.meta.valid_width;
        msg->valid_window.height = p.meta.valid_height;
        pub->publish(msg);
    });
}

void StereoProcessor::sendPoints(GpuMatSource points, GpuMatSource color, const sensor_msgs::ImageConstPtr &pattern, ros::Publisher *pub)
{
    const std_msgs::Header header = pattern->header;
    stereoProcessor_->enqueueSendPoints(points, color, [header, pub](const PointCloud2Payload &p) {
        sensor_msgs::PointCloud2Ptr msg(new sensor_msgs::PointCloud2);
        msg->header = header;
        msg->height = p.meta.height;
        msg->width = p.meta.width;
        msg->is_bigendian = p.meta.is_bigendian != 0;
        msg->is_dense = p.meta.is_dense != 0;
        msg->point_step = p.meta.point_step;
        msg->row_step = p.meta.row_step;
        const char *names[4] = {"x", "y", "z", "rgb"};
        const int offsets[4] = {p.meta.off_x, p.meta.off_y, p.meta.off_z, p.meta.off_rgb};
        msg->fields.resize(4);
        for (int i = 0; i < 4; ++i) {
            msg->fields[i].name = names[i];
            msg->fields[i].offset = offsets[i];
            msg->fields[i].datatype = sensor_msgs::PointField::FLOAT32;
            msg->fields[i].count = 1;
        }
        msg->data.assign(p.data, p.data + p.size);     // records are packed on the GPU (src/GpuSenderPc2.cpp:15-72 did this on the host)
        pub->publish(msg);
    });
}

void StereoProcessor::imageCb(const sensor_msgs::ImageConstPtr &l_raw_msg, const sensor_msgs::ImageConstPtr &r_raw_msg)
{
    const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    boost::lock_guard<boost::recursive_mutex> config_lock(config_mutex_);
    boost::lock_guard<boost::mutex> connect_lock(connect_mutex_);
    const ConnectedTopics want = connected_;
    try {
        stereoProcessor_->cleanSenders();
        if (!stereoProcessor_->isStereoModelInitialised() && camera_info_from_files_)
            stereoProcessor_->initStereoModel(camera_info_file_left_, camera_info_file_right_);

        uploadRaw(GPU_MAT_SRC_L_RAW, l_raw_msg);
        uploadRaw(GPU_MAT_SRC_R_RAW, r_raw_msg);
        const double t_upload = msSince(t0);

        if (want.needsMonoLeft()) stereoProcessor_->convertRawToMono(GPU_MAT_SIDE_L);
        if (want.has(ConnectedTopics::MONO_LEFT)) sendImage(GPU_MAT_SRC_L_MONO, l_raw_msg, sensor_msgs::image_encodings::MONO8, pub(ConnectedTopics::MONO_LEFT));
        if (want.needsMonoRight()) stereoProcessor_->convertRawToMono(GPU_MAT_SIDE_R);
        if (want.has(ConnectedTopics::MONO_RIGHT)) sendImage(GPU_MAT_SRC_R_MONO, r_raw_msg, sensor_msgs::image_encodings::MONO8, pub(ConnectedTopics::MONO_RIGHT));
        if (want.needsColorLeft()) stereoProcessor_->convertRawToColor(GPU_MAT_SIDE_L);
        if (want.has(ConnectedTopics::COLOR_LEFT)) sendImage(GPU_MAT_SRC_L_COLOR, l_raw_msg, sensor_msgs::image_encodings::BGR8, pub(ConnectedTopics::COLOR_LEFT));
        if (want.needsColorRight()) stereoProcessor_->convertRawToColor(GPU_MAT_SIDE_R);
        if (want.has(ConnectedTopics::COLOR_RIGHT)) sendImage(GPU_MAT_SRC_R_COLOR, r_raw_msg, sensor_msgs::image_encodings::BGR8, pub(ConnectedTopics::COLOR_RIGHT));
        const double t_convert = msSince(t0);

        if (want.needsRectMonoLeft()) stereoProcessor_->rectifyImage(GPU_MAT_SRC_L_MONO, GPU_MAT_SRC_L_RECT_MONO, B200S_INTER_LINEAR);
        if (want.needsRectMonoRight()) stereoProcessor_->rectifyImage(GPU_MAT_SRC_R_MONO, GPU_MAT_SRC_R_RECT_MONO, B200S_INTER_LINEAR);
        if (want.has(ConnectedTopics::RECT_MONO_LEFT)) sendImage(GPU_MAT_SRC_L_RECT_MONO, l_raw_msg, sensor_msgs::image_encodings::MONO8, pub(ConnectedTopics::RECT_MONO_LEFT));
        if (want.has(ConnectedTopics::RECT_MONO_RIGHT)) sendImage(GPU_MAT_SRC_R_RECT_MONO, r_raw_msg, sensor_msgs::image_encodings::MONO8, pub(ConnectedTopics::RECT_MONO_RIGHT));
        if (want.needsRectColorLeft()) stereoProcessor_->rectifyImage(GPU_MAT_SRC_L_COLOR, GPU_MAT_SRC_L_RECT_COLOR, B200S_INTER_LINEAR);
        if (want.needsRectColorRight()) stereoProcessor_->rectifyImage(GPU_MAT_SRC_R_COLOR, GPU_MAT_SRC_R_RECT_COLOR, B200S_INTER_LINEAR);
        if (want.has(ConnectedTopics::RECT_COLOR_LEFT)) sendImage(GPU_MAT_SRC_L_RECT_COLOR, l_raw_msg, sensor_msgs::image_encodings::BGR8, pub(ConnectedTopics::RECT_COLOR_LEFT));
        if (want.has(ConnectedTopics::RECT_COLOR_RIGHT)) sendImage(GPU_MAT_SRC_R_RECT_COLOR, r_raw_msg, sensor_msgs::image_encodings::BGR8, pub(ConnectedTopics::RECT_COLOR_RIGHT));
        const double t_rectify = msSince(t0);

        if (want.needsDisparity()) {
            // cv::StereoBM semantics incl. its own speckle stage; filterSpeckles is kept for call-sequence parity and is
            // idempotent on an already filtered plane
            stereoProcessor_->computeDisparity(GPU_MAT_SRC_L_RECT_MONO, GPU_MAT_SRC_R_RECT_MONO, GPU_MAT_SRC_L_DISPARITY);
            stereoProcessor_->filterSpeckles(GPU_MAT_SRC_L_DISPARITY);
        }
        if (want.has(ConnectedTopics::DISPARITY)) sendDisparity(GPU_MAT_SRC_L_DISPARITY, l_raw_msg, pub(ConnectedTopics::DISPARITY));
        const double t_disparity = msSince(t0);

        if (want.has(ConnectedTopics::DISPARITY_VIS)) {
            stereoProcessor_->computeDisparityImage(GPU_MAT_SRC_L_DISPARITY, GPU_MAT_SRC_L_DISPARITY_IMG);
            sendImage(GPU_MAT_SRC_L_DISPARITY_IMG, l_raw_msg, sensor_msgs::image_encodings::BGRA8, pub(ConnectedTopics::DISPARITY_VIS));
        }
        const double t_vis = msSince(t0);

        if (want.has(ConnectedTopics::POINTCLOUD)) {
            stereoProcessor_->projectDisparityTo3DPoints(GPU_MAT_SRC_L_DISPARITY, GPU_MAT_SRC_L_POINTS2);
            sendPoints(GPU_MAT_SRC_L_POINTS2, GPU_MAT_SRC_L_RECT_COLOR, l_raw_msg, pub(ConnectedTopics::POINTCLOUD));
        }
        stereoProcessor_->waitForAllStreams();      // every sender has published from its stream callback by now
        stereoProcessor_->cleanSenders();
        const double t_total = msSince(t0);
        ROS_DEBUG("TIMING [ms]: upload:%.2f; color convert:%.2f; rectify:%.2f; disparity:%.2f; disparity img:%.2f; pc2:%.2f; Total:%.2f;", t_upload,
                  t_convert - t_upload, t_rectify - t_convert, t_disparity - t_rectify, t_vis - t_disparity, t_total - t_vis, t_total);
    } catch (const Error &e) {
        ROS_ERROR("gpuimageproc: frame dropped, %s (code %d)", e.what(), e.code);
    }
}

void StereoProcessor::configCb(Config &config, uint32_t /*level*/)
{
    // Tweak the settings to valid values first, as the reference does
    config.correlation_window_size |= 0x1;                          // must be odd
    config.disparity_range = (config.disparity_range / 16) * 16;    // must be a multiple of 16
    config.prefilter_size |= 0x1;

    stereoProcessor_->setPreFilterType(config.xsobel ? 1 /* PREFILTER_XSOBEL */ : 0 /* PREFILTER_NORMALIZED_RESPONSE */);
    stereoProcessor_->setPreFilterSize(config.prefilter_size);
    stereoProcessor_->setPreFilterCap(config.prefilter_cap);
    stereoProcessor_->setRefineDisparity(config.refine_disparity);
    stereoProcessor_->setBlockSize(config.correlation_window_size);
    stereoProcessor_->setNumDisparities(config.disparity_range);
    stereoProcessor_->setMinDisparity(config.disparity_min);        // the reference passes disparity_range here (bug B2)
    stereoProcessor_->setTextureThreshold((int)config.texture_threshold);
    stereoProcessor_->setUniquenessRatio((int)config.uniqueness_ratio);
    stereoProcessor_->setDisp12MaxDiff(config.disp12_max_diff);
    stereoProcessor_->setMaxSpeckleDiff(config.max_speckle_diff);
    stereoProcessor_->setMaxSpeckleSize(config.max_speckle_size);
    // bilateral_filter / filter_* are accepted and ignored: every use is commented out in the reference (:324-335)
    ROS_INFO("Reconfigure winsz:%d ndisp:%d mind:%d tex:%3.1f", config.correlation_window_size, config.disparity_range, config.disparity_min,
             config.texture_threshold);
}

}  // namespace gpuimageproc
