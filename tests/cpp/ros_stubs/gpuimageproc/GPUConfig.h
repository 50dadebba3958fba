// what dynamic_reconfigure generates from cfg/GPU.cfg (fields only)
#pragma once
namespace gpuimageproc {
struct GPUConfig {
    bool xsobel; int prefilter_size, prefilter_cap; bool refine_disparity;
    int correlation_window_size, disparity_min, disparity_range;
    bool bilateral_filter; int filter_ndisp, filter_radius, filter_iters; double filter_edge_threshold, filter_max_disc_threshold, filter_sigma_range;
    double texture_threshold, uniqueness_ratio; int disp12_max_diff;
    int max_speckle_size; double max_speckle_diff;
};
}
