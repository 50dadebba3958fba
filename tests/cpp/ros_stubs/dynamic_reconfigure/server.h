#pragma once
#include <functional>
#include <mutex>
#include <ros/ros.h>
namespace dynamic_reconfigure {
template <class ConfigType> class Server {
  public:
    typedef std::function<void(ConfigType &, uint32_t)> CallbackType;
    Server(std::recursive_mutex &mutex, const ros::NodeHandle &nh);
    void setCallback(const CallbackType &callback);
};
}
