// TEST INFRASTRUCTURE ONLY: forwards <tbb/blocked_range2d.h> to the std::thread stand-in (see tbb_shim.h).
#include "tbb_shim.h"
