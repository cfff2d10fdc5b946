// TEST INFRASTRUCTURE ONLY: forwards <oneapi/tbb/global_control.h> to the std::thread stand-in.
#include "../../tbb/tbb_shim.h"
