#pragma once
#include <ros/ros.h>
namespace image_transport {
struct TransportHints { TransportHints(const std::string &default_transport, const ros::TransportHints &hints, const ros::NodeHandle &nh); };
class ImageTransport { public: explicit ImageTransport(const ros::NodeHandle &nh); };
}
