// Which output topics currently have subscribers, and what each of them needs upstream.
// Same role as the reference's bit-field (include/gpuimageproc/ConnectedTopics.h:5-28); the dependency rules that the
// reference spells out as or-chains inside imageCb (src/StereoProcessor.cpp:187-281) live here as named predicates.
#pragma once
#include <stdint.h>

namespace gpuimageproc
{

struct ConnectedTopics
{
    enum Topic
    {
        MONO_LEFT = 1u << 0, MONO_RIGHT = 1u << 1, COLOR_LEFT = 1u << 2, COLOR_RIGHT = 1u << 3,
        RECT_MONO_LEFT = 1u << 4, RECT_MONO_RIGHT = 1u << 5, RECT_COLOR_LEFT = 1u << 6, RECT_COLOR_RIGHT = 1u << 7,
        DISPARITY = 1u << 8, DISPARITY_VIS = 1u << 9, POINTCLOUD = 1u << 10
    };
    uint32_t bits;
    ConnectedTopics() : bits(0) {}
    void set(Topic t, bool on) { bits = on ? (bits | t) : (bits & ~(uint32_t)t); }
    bool has(Topic t) const { return (bits & t) != 0; }
    bool any() const { return bits != 0; }
    // 0 = nobody listens; otherwise the index of the deepest stage somebody needs (the reference's level())
    int level() const { return bits == 0 ? 0 : 32 - __builtin_clz(bits); }

    bool needsDisparity() const { return (bits & (DISPARITY | DISPARITY_VIS | POINTCLOUD)) != 0; }
    bool needsMonoLeft() const { return has(MONO_LEFT) || has(RECT_MONO_LEFT) || needsDisparity(); }
    bool needsMonoRight() const { return has(MONO_RIGHT) || has(RECT_MONO_RIGHT) || needsDisparity(); }
    bool needsColorLeft() const { return has(COLOR_LEFT) || has(RECT_COLOR_LEFT) || has(POINTCLOUD); }
    bool needsColorRight() const { return has(COLOR_RIGHT) || has(RECT_COLOR_RIGHT); }
    bool needsRectMonoLeft() const { return has(RECT_MONO_LEFT) || needsDisparity(); }
    bool needsRectMonoRight() const { return has(RECT_MONO_RIGHT) || needsDisparity(); }
    bool needsRectColorLeft() const { return has(RECT_COLOR_LEFT) || has(POINTCLOUD); }
    bool needsRectColorRight() const { return has(RECT_COLOR_RIGHT); }
};

}  // namespace gpuimageproc
