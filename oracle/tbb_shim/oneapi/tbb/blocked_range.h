// TEST INFRASTRUCTURE ONLY: forwards <oneapi/tbb/blocked_range.h> to the std::thread stand-in.
#include "../../tbb/tbb_shim.h"
