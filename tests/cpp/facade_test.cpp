// Exercises include/b200_gpuimageproc/GpuStereoProcessor.hpp the way the reference's gtest drives its class
// (test/UTest.cpp: RectifyMonoGpu :262-288, DisparityGpu :290-331, PointCloud :365-398).
// usage: facade_test <left.yaml> <right.yaml> <in.bin> <out.bin> <W> <H> <nd> <block>
//   in.bin  = left raw (W*H u8) followed by right raw;  out.bin = rectL, rectR (u8), disparity (s16), pointcloud2 (32 B/px)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "b200_gpuimageproc/GpuStereoProcessor.hpp"

using namespace gpuimageproc;

int main(int argc, char **argv)
{
    if (argc < 9) { std::fprintf(stderr, "usage\n"); return 2; }
    const int W = std::atoi(argv[5]), H = std::atoi(argv[6]), nd = std::atoi(argv[7]), block = std::atoi(argv[8]);
    try {
        GpuStereoProcessor proc;
        Mat l(H, W, B200S_8UC1), r(H, W, B200S_8UC1);
        FILE *f = std::fopen(argv[3], "rb");
        if (!f || std::fread(l.data.data(), 1, l.data.size(), f) != l.data.size() || std::fread(r.data.data(), 1, r.data.size(), f) != r.data.size()) return 3;
        std::fclose(f);
        bool threw = false;   // reference: assert(model_.initialized())
        try { proc.uploadMat(GPU_MAT_SRC_L_MONO, l); proc.rectifyImage(GPU_MAT_SRC_L_MONO, GPU_MAT_SRC_L_RECT_MONO); } catch (const Error &e) { threw = e.code == B200S_ENOTINIT; }
        if (!threw) return 4;
        proc.initStereoModel(std::string(argv[1]), std::string(argv[2]));
        if (!proc.isStereoModelInitialised()) return 5;
        proc.setPreFilterType(1); proc.setPreFilterSize(9); proc.setPreFilterCap(31);
        proc.setNumDisparities(nd); proc.setBlockSize(block); proc.setMinDisparity(0);
        proc.setTextureThreshold(10); proc.setUniquenessRatio(15); proc.setDisp12MaxDiff(-1);
        proc.setMaxSpeckleSize(0);
        proc.uploadMat(GPU_MAT_SRC_L_RAW, l, "mono8");
        proc.uploadMat(GPU_MAT_SRC_R_RAW, r, "mono8");
        proc.convertRawToMono(GPU_MAT_SIDE_L);
        proc.convertRawToMono(GPU_MAT_SIDE_R);
        proc.rectifyImage(GPU_MAT_SRC_L_MONO, GPU_MAT_SRC_L_RECT_MONO);
        proc.rectifyImage(GPU_MAT_SRC_R_MONO, GPU_MAT_SRC_R_RECT_MONO);
        proc.computeDisparity(GPU_MAT_SRC_L_RECT_MONO, GPU_MAT_SRC_R_RECT_MONO, GPU_MAT_SRC_L_DISPARITY);
        proc.filterSpeckles(GPU_MAT_SRC_L_DISPARITY);
        // the reference's exact ids (test/UTest.cpp:378-382, src/StereoProcessor.cpp:272-281): DISPARITY_32F into
        // projectDisparityTo3DPoints, POINTS2 into enqueueSendPoints
        proc.projectDisparityTo3DPoints(GPU_MAT_SRC_L_DISPARITY_32F, GPU_MAT_SRC_L_POINTS2);
        int published = 0;
        GPUSenderPc2Ptr snd = proc.enqueueSendPoints(GPU_MAT_SRC_L_POINTS2, GPU_MAT_SRC_L_RECT_MONO,
                                                     [&published](const PointCloud2Payload &m) { published += m.meta.point_step; });
        GPUSenderDisparityPtr dsnd = proc.enqueueSendDisparity(GPU_MAT_SRC_L_DISPARITY);
        GPUSenderImagePtr isnd = proc.enqueueSendImage(GPU_MAT_SRC_L_RECT_MONO, "mono8");
        proc.waitForAllStreams();      // the senders complete from stream callbacks (src/GpuSenderIfc.cpp:13-26)
        if (!snd->wasDataSent() || !dsnd->wasDataSent() || !isnd->wasDataSent() || published != 32) return 7;
        const PointCloud2Payload &pc = snd->payload;
        Mat rl, rr, d;
        proc.downloadMat(GPU_MAT_SRC_L_RECT_MONO, rl);
        proc.downloadMat(GPU_MAT_SRC_R_RECT_MONO, rr);
        proc.downloadMat(GPU_MAT_SRC_L_DISPARITY, d);
        if (pc.meta.point_step != 32 || pc.meta.width != W || d.type != B200S_16SC1) return 6;
        if (std::memcmp(isnd->payload.data, rl.data.data(), rl.data.size()) != 0 || dsnd->payload.meta.width != W) return 8;
        proc.convertColor(GPU_MAT_SRC_L_RECT_MONO, GPU_MAT_SRC_L_RECT_COLOR, "mono8", "bgr8");
        proc.printStats("Disparity", GPU_MAT_SRC_L_DISPARITY);
        FILE *o = std::fopen(argv[4], "wb");
        std::fwrite(rl.data.data(), 1, rl.data.size(), o);
        std::fwrite(rr.data.data(), 1, rr.data.size(), o);
        std::fwrite(d.data.data(), 1, d.data.size(), o);
        std::fwrite(pc.data, 1, pc.size, o);
        std::fclose(o);
        proc.cleanSenders();
        std::printf("facade ok\n");
        return 0;
    } catch (const Error &e) {
        std::fprintf(stderr, "Error %d: %s\n", e.code, e.what());
        return 1;
    }
}
