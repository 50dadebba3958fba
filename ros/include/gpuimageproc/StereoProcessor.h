// ROS 1 glue of the B200 build: the node / nodelet body that feeds camera topics through
// gpuimageproc::GpuStereoProcessor (include/b200_gpuimageproc/GpuStereoProcessor.hpp over libb200stereo.so) and publishes
// the reference's eleven topics.  Keeps the reference's surface (include/gpuimageproc/StereoProcessor.h:25-110): the
// constructor signature, the topic names, the ~queue_size / ~approximate_sync / ~camera_info_file_left/right /
// ~publisher_queue_size parameters and the GPU.cfg dynamic_reconfigure server.  Builds only where ROS exists
// (ros/CMakeLists.txt); tests/cpp/ros_stubs/ holds just enough of the ROS headers for a syntax check here.
#pragma once
#include <boost/shared_ptr.hpp>
#include <boost/thread/mutex.hpp>
#include <boost/thread/recursive_mutex.hpp>

#include <dynamic_reconfigure/server.h>
#include <image_transport/image_transport.h>
#include <image_transport/subscriber_filter.h>
#include <message_filters/subscriber.h>
#include <message_filters/sync_policies/approximate_time.h>
#include <message_filters/sync_policies/exact_time.h>
#include <message_filters/synchronizer.h>
#include <ros/ros.h>
#include <sensor_msgs/CameraInfo.h>
#include <sensor_msgs/Image.h>
#include <sensor_msgs/PointCloud2.h>
#include <stereo_msgs/DisparityImage.h>

#include "b200_gpuimageproc/GpuStereoProcessor.hpp"
#include "gpuimageproc/ConnectedTopics.h"
#include "gpuimageproc/GPUConfig.h"

namespace gpuimageproc
{

class StereoProcessor
{
  public:
    static const std::string CAMERA_TOPIC_LEFT, CAMERA_TOPIC_RIGHT, CAMERA_TOPIC_IMAGE, CAMERA_TOPIC_INFO;

    StereoProcessor(ros::NodeHandle &nh, ros::NodeHandle &private_nh);

  protected:
    typedef sensor_msgs::Image Image;
    typedef sensor_msgs::CameraInfo CameraInfoMsg;
    typedef message_filters::sync_policies::ExactTime<Image, Image> ExactImages;
    typedef message_filters::sync_policies::ExactTime<Image, CameraInfoMsg, Image, CameraInfoMsg> ExactImagesAndInfo;
    typedef message_filters::sync_policies::ApproximateTime<Image, Image> ApproxImages;
    typedef message_filters::sync_policies::ApproximateTime<Image, CameraInfoMsg, Image, CameraInfoMsg> ApproxImagesAndInfo;
    typedef gpuimageproc::GPUConfig Config;
    typedef dynamic_reconfigure::Server<Config> ReconfigureServer;

    void connectCb();
    void imageAndInfoCb(const sensor_msgs::ImageConstPtr &l_raw_msg, const sensor_msgs::CameraInfoConstPtr &l_info_msg,
                        const sensor_msgs::ImageConstPtr &r_raw_msg, const sensor_msgs::CameraInfoConstPtr &r_info_msg);
    void imageCb(const sensor_msgs::ImageConstPtr &l_raw_msg, const sensor_msgs::ImageConstPtr &r_raw_msg);
    void configCb(Config &config, uint32_t level);

    // message construction on the stream-callback thread (what the reference's GpuSender*::fillInData do)
    void sendImage(GpuMatSource source, const sensor_msgs::ImageConstPtr &pattern, const std::string &encoding, ros::Publisher *pub);
    void sendDisparity(GpuMatSource source, const sensor_msgs::ImageConstPtr &pattern, ros::Publisher *pub);
    void sendPoints(GpuMatSource points, GpuMatSource color, const sensor_msgs::ImageConstPtr &pattern, ros::Publisher *pub);
    void uploadRaw(GpuMatSource id, const sensor_msgs::ImageConstPtr &msg);

    ros::NodeHandle &nh;
    ros::NodeHandle &private_nh;
    boost::shared_ptr<image_transport::ImageTransport> it_;
    image_transport::SubscriberFilter sub_l_raw_image_, sub_r_raw_image_;
    message_filters::Subscriber<CameraInfoMsg> sub_l_info_, sub_r_info_;
    boost::shared_ptr<message_filters::Synchronizer<ExactImages> > exact_sync_images_;
    boost::shared_ptr<message_filters::Synchronizer<ExactImagesAndInfo> > exact_sync_images_and_info_;
    boost::shared_ptr<message_filters::Synchronizer<ApproxImages> > approximate_sync_images_;
    boost::shared_ptr<message_filters::Synchronizer<ApproxImagesAndInfo> > approximate_sync_images_and_info_;

    // the eleven output topics, one publisher per ConnectedTopics bit (table in StereoProcessor.cpp)
    enum { N_TOPICS = 11 };
    ros::Publisher publishers_[N_TOPICS];
    ros::Publisher *pub(ConnectedTopics::Topic t) { return &publishers_[__builtin_ctz((unsigned)t)]; }
    boost::mutex connect_mutex_;
    ConnectedTopics connected_;
    std::string camera_info_file_left_, camera_info_file_right_;
    bool camera_info_from_files_;

    boost::shared_ptr<GpuStereoProcessor> stereoProcessor_;
    boost::recursive_mutex config_mutex_;
    boost::shared_ptr<ReconfigureServer> reconfigure_server_;
};

}  // namespace gpuimageproc
