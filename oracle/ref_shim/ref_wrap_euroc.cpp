// TEST INFRASTRUCTURE ONLY -- a stand-alone program over the REFERENCE's own EuRoC readers: the bodies of LoadImages / LoadIMU
// (Examples/Monocular-Inertial/mono_inertial_euroc.cc:252-310) are cut out of the reference at build time (extract_ranges.py ->
// oracle/_ref/gen/euroc_loaders.inc) and compiled as they are; cv::Point3f is the stand-in of minicv.hpp.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "minicv.hpp"

using namespace std;

#include "euroc_loaders.inc"

// Stand-alone form (oracle/_ref/ref_euroc <imagePath> <timesFile> <imuFile>): prints what the reference's readers return, one record per line with
// every double / float as its exact hexadecimal value (%a).  tests/test_euroc_io_cpu.py runs it as a subprocess.
int main(int argc, char** argv) {
    if (argc < 4) return 2;
    vector<string> v; vector<double> t;
    LoadImages(argv[1], argv[2], v, t);
    printf("images %zu\n", v.size());
    for (size_t i = 0; i < v.size(); ++i) printf("%s %a\n", v[i].c_str(), t[i]);
    vector<double> ti; vector<cv::Point3f> a, g;
    LoadIMU(argv[3], ti, a, g);
    printf("imu %zu\n", ti.size());
    for (size_t i = 0; i < ti.size(); ++i) printf("%a %a %a %a %a %a %a\n", ti[i], (double)a[i].x, (double)a[i].y, (double)a[i].z, (double)g[i].x, (double)g[i].y, (double)g[i].z);
    return 0;
}
