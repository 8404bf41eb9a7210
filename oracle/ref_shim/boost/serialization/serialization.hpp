// TEST INFRASTRUCTURE: forwards <boost/serialization/serialization.hpp> to the stand-in (see _common.hpp).
#pragma once
#include "_common.hpp"
