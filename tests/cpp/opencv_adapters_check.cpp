// Instantiates the cv::Mat overloads of the facade (compiled with -DB200S_WITH_OPENCV against the stub cv::Mat, syntax only).
#include "b200_gpuimageproc/GpuStereoProcessor.hpp"

void use_adapters(gpuimageproc::GpuStereoProcessor &p, cv::Mat &a, cv::Mat &b, cv::Mat &c)
{
    using namespace gpuimageproc;
    p.uploadMat(GPU_MAT_SRC_L_RAW, a, "mono8");
    p.downloadMat(GPU_MAT_SRC_L_RAW, b);
    p.rectifyImageLeft(a, b);
    p.rectifyImageRight(a, b);
    p.computeDisparityBare(a, b, c);
    p.computeDisparity(a, b, c);
    p.filterSpeckles(c);
    p.printStats("disparity", c);
}
