// Stand-alone executable `gpuimageproc_node` (reference: src/StereoProcessorNode.cpp:4-34): one StereoProcessor on the
// node's public and private handles, then the ROS event loop.  Exit code 2 when no usable GPU is present.
#include <cstdio>

#include <ros/ros.h>

#include "gpuimageproc/StereoProcessor.h"

int main(int argc, char **argv)
{
    ros::init(argc, argv, "gpuimageproc");
    ros::NodeHandle public_handle;
    ros::NodeHandle private_handle("~");
    try {
        gpuimageproc::StereoProcessor pipeline(public_handle, private_handle);
        ros::spin();
    } catch (const gpuimageproc::Error &e) {
        std::fprintf(stderr, "gpuimageproc_node: %s (code %d)\n", e.what(), e.code);
        return 2;
    }
    return 0;
}
