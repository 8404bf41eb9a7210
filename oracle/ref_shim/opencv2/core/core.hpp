// TEST INFRASTRUCTURE: forwards <opencv2/core/core.hpp> to the minimal type stand-in (see minicv.hpp).
#pragma once
#include "../../minicv.hpp"
