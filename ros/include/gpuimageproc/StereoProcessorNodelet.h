// Nodelet wrapper (reference: include/gpuimageproc/StereoProcessorNodelet.h:8-16); exported as gpuimageproc/Stereoproc by
// plugins/nodelet_plugins.xml.
#pragma once
#include <nodelet/nodelet.h>

#include "gpuimageproc/StereoProcessor.h"

namespace gpuimageproc
{

class StereoProcNodelet : public nodelet::Nodelet
{
  public:
    virtual void onInit();

  protected:
    ros::NodeHandle nh_, private_nh_;      // StereoProcessor keeps references to its node handles
    boost::shared_ptr<StereoProcessor> stereoProcessorPtr;
};

}  // namespace gpuimageproc
