#pragma once
#include <mutex>
namespace boost { using mutex = std::mutex; template <class M> using lock_guard = std::lock_guard<M>; }
