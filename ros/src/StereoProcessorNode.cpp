// Stand-alone node (reference: src/StereoProcessorNode.cpp:4-34): node name "gpuimageproc", public and private handles,
// one StereoProcessor, spin.
#include <ros/ros.h>

#include "gpuimageproc/StereoProcessor.h"

int main(int argc, char **argv)
{
    ros::init(argc, argv, "gpuimageproc");
    ros::NodeHandle nh;
    ros::NodeHandle private_nh("~");
    gpuimageproc::StereoProcessor processor(nh, private_nh);
    ros::spin();
    return 0;
}
