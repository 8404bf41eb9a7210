// TEST INFRASTRUCTURE: forwards <opencv2/features2d/features2d.hpp> to the minimal type stand-in (see minicv.hpp).
#pragma once
#include "../../minicv.hpp"
