// Compiles against the C++ host mirror (include/orb_b200/orb_slam3.hpp) like a reference caller would
// (Frame::ExtractORB, src/Frame.cc:418-425) and runs one extraction.  Built and run by tests/test_cpp_shim.py.
#include <cstdio>
#include <vector>
#include "orb_b200/orb_slam3.hpp"

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const int rows = atoi(argv[2]), cols = atoi(argv[3]);
    std::vector<uint8_t> img((size_t)rows * cols);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(img.data(), 1, img.size(), f) != img.size()) return 3;
    fclose(f);
    ORB_SLAM3::ORBextractor extractor(1000, 1.2f, 8, 20, 7, cols, rows);
    ORB_SLAM3::Image im; im.data = img.data(); im.rows = rows; im.cols = cols; im.step = cols;
    std::vector<ORB_SLAM3::KeyPoint> mvKeys;
    ORB_SLAM3::Descriptors mDescriptors;
    std::vector<int> vLapping = {0, 1000};
    const int monoLeft = extractor(im, ORB_SLAM3::Image(), mvKeys, mDescriptors, vLapping);
    unsigned long long sum = 0;
    for (uint8_t b : mDescriptors.data) sum = sum * 1315423911ull + b;
    printf("%d %zu %llu\n", monoLeft, mvKeys.size(), sum);
    ORB_SLAM3::Image empty;
    return extractor(empty, empty, mvKeys, mDescriptors, vLapping) == -1 ? 0 : 4;
}
