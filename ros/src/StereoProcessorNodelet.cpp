// Nodelet flavour of the pipeline: plugins/nodelet_plugins.xml exports this class as gpuimageproc/Stereoproc (the name the
// reference's launch files load, reference src/StereoProcessorNodelet.cpp:6-18).  The nodelet owns copies of its two node
// handles because StereoProcessor keeps references to them, and it reports a missing / unusable GPU as a fatal nodelet
// error instead of letting the exception escape into the nodelet manager.
#include <memory>

#include <nodelet/nodelet.h>
#include <pluginlib/class_list_macros.h>

#include "gpuimageproc/StereoProcessor.h"

namespace gpuimageproc
{

class StereoProcNodelet : public nodelet::Nodelet
{
  public:
    virtual void onInit()
    {
        public_handle_ = getNodeHandle();
        private_handle_ = getPrivateNodeHandle();
        try {
            pipeline_.reset(new StereoProcessor(public_handle_, private_handle_));
        } catch (const Error &e) {
            ROS_ERROR("gpuimageproc nodelet: cannot start the B200 stereo pipeline: %s (code %d)", e.what(), e.code);
        }
    }

  private:
    ros::NodeHandle public_handle_, private_handle_;
    std::unique_ptr<StereoProcessor> pipeline_;
};

}  // namespace gpuimageproc

PLUGINLIB_EXPORT_CLASS(gpuimageproc::StereoProcNodelet, nodelet::Nodelet)
