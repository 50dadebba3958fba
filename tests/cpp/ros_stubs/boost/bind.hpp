#pragma once
#include <functional>
namespace boost { using std::bind; }
using namespace std::placeholders;
