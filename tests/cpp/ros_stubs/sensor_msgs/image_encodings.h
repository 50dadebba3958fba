#pragma once
#include <string>
namespace sensor_msgs { namespace image_encodings {
const std::string MONO8 = "mono8", BGR8 = "bgr8", RGB8 = "rgb8", BGRA8 = "bgra8", TYPE_32FC1 = "32FC1";
int numChannels(const std::string &encoding);
int bitDepth(const std::string &encoding);
} }
