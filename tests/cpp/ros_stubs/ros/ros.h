#pragma once
#include <functional>
#include <memory>
#include <string>
namespace ros { void stub_log(const char *fmt, ...); }
#define ROS_INFO(...) ros::stub_log(__VA_ARGS__)
#define ROS_DEBUG(...) ros::stub_log(__VA_ARGS__)
#define ROS_ERROR(...) ros::stub_log(__VA_ARGS__)
namespace ros {
struct SingleSubscriberPublisher {};
typedef std::function<void(const SingleSubscriberPublisher &)> SubscriberStatusCallback;
struct TransportHints {};
struct Publisher {
    unsigned getNumSubscribers() const;
    template <class M> void publish(const std::shared_ptr<M> &msg) const;
};
struct Subscriber { explicit operator bool() const; };
class NodeHandle {
  public:
    NodeHandle();
    explicit NodeHandle(const std::string &ns);
    template <class T> bool param(const std::string &name, T &value, const T &def) const;
    template <class T> void setParam(const std::string &name, const T &value) const;
    template <class M> Publisher advertise(const std::string &topic, unsigned queue, const SubscriberStatusCallback &connect, const SubscriberStatusCallback &disconnect);
};
void init(int &argc, char **argv, const std::string &name);
void spin();
}
