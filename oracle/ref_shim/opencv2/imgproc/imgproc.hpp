// TEST INFRASTRUCTURE: forwards <opencv2/imgproc/imgproc.hpp> to the minimal type stand-in (see minicv.hpp).
#pragma once
#include "../../minicv.hpp"
