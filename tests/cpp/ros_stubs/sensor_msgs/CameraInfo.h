#pragma once
#include <memory>
#include <vector>
#include <std_msgs/Header.h>
namespace sensor_msgs {
struct CameraInfo { std_msgs::Header header; uint32_t height = 0, width = 0; std::string distortion_model; std::vector<double> D; double K[9], R[9], P[12]; };
typedef std::shared_ptr<const CameraInfo> CameraInfoConstPtr;
}
