// TEST INFRASTRUCTURE: forwards <boost/serialization/assume_abstract.hpp> to the stand-in (see _common.hpp).
#pragma once
#include "_common.hpp"
