#pragma once
namespace message_filters {
template <class Policy> class Synchronizer {
  public:
    template <class F0, class F1> Synchronizer(const Policy &p, F0 &f0, F1 &f1);
    template <class F0, class F1, class F2, class F3> Synchronizer(const Policy &p, F0 &f0, F1 &f1, F2 &f2, F3 &f3);
    template <class C> void registerCallback(const C &callback);
};
}
