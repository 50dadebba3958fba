// reference: src/StereoProcessorNodelet.cpp:6-18
#include "gpuimageproc/StereoProcessorNodelet.h"

#include <pluginlib/class_list_macros.h>

namespace gpuimageproc
{

void StereoProcNodelet::onInit()
{
    nh_ = getNodeHandle();
    private_nh_ = getPrivateNodeHandle();
    stereoProcessorPtr.reset(new StereoProcessor(nh_, private_nh_));
}

}  // namespace gpuimageproc

PLUGINLIB_EXPORT_CLASS(gpuimageproc::StereoProcNodelet, nodelet::Nodelet)
