#pragma once
#include <sensor_msgs/Image.h>
namespace sensor_msgs { struct RegionOfInterest { uint32_t x_offset = 0, y_offset = 0, height = 0, width = 0; bool do_rectify = false; }; }
namespace stereo_msgs {
struct DisparityImage { std_msgs::Header header; sensor_msgs::Image image; float f = 0, T = 0; sensor_msgs::RegionOfInterest valid_window; float min_disparity = 0, max_disparity = 0, delta_d = 0; };
typedef std::shared_ptr<DisparityImage> DisparityImagePtr;
}
