#pragma once
#include <image_transport/image_transport.h>
namespace image_transport {
class SubscriberFilter {
  public:
    void subscribe(ImageTransport &it, const std::string &topic, unsigned queue, const TransportHints &hints);
    void unsubscribe();
    const ros::Subscriber &getSubscriber() const;
};
}
