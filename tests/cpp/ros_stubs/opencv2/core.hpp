// the slice of cv::Mat the B200S_WITH_OPENCV adapters of GpuStereoProcessor.hpp touch (syntax check only)
#pragma once
#include <stddef.h>
namespace cv {
class Mat {
  public:
    int rows = 0, cols = 0;
    unsigned char *data = nullptr;
    size_t step = 0;
    int type() const;
    int channels() const;
    void create(int rows, int cols, int type);
};
}
