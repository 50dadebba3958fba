#pragma once
#include <mutex>
namespace boost { using recursive_mutex = std::recursive_mutex; }
