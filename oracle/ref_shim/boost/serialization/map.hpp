// TEST INFRASTRUCTURE: forwards <boost/serialization/map.hpp> to the stand-in (see _common.hpp).
#pragma once
#include "_common.hpp"
